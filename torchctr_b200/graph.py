"""Whole-step CUDA graphs: forward + backward + fused table update + dense optimizer captured once,
replayed per batch.  The reference loop (``torchctr/trainer.py:291-303``) issues ~150 kernel
launches per step from Python; at B200 speeds the step is launch-bound, a graph replay is not.

    step = GraphedTrainStep(model, optimizer, example_batch)
    loss = step(batch)            # batch: host (pinned or not) or device tensors, same shapes

Inputs are staged into static device buffers (asynchronous copies when the batch sits in pinned
host memory), the loss is a static device scalar.  Table hyper-parameters (lr
schedule, Adam step) reach the kernels through a device tensor refreshed before each replay.
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, example_batch, warmup: int = 3, backward_fn=None, zero_grad_fn=None):
        """``backward_fn(loss)`` replaces ``loss.backward()`` (multi-GPU: scale by 1 / world, all-reduce the
        replicated gradients -- NCCL collectives are captured with the rest of the step)."""
        self._backward_fn = backward_fn
        self._zero_grad_fn = zero_grad_fn        # replaces optimizer.zero_grad(set_to_none=True), e.g. flat gradient buffers
        feats, labels = example_batch
        self.model, self.optimizer = model, optimizer
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device")
        for m in model.modules():
            if getattr(m, "index_kind", None) == "vocab":
                raise RuntimeError("growing vocabularies read the table size on the host every step; "
                                   "such models cannot be captured into a CUDA graph")
        self.device = dev
        # one packed static buffer: every tensor of the batch is a view into it
        self._keys = [k for k in feats if torch.is_tensor(feats[k])]
        tensors = [feats[k] for k in self._keys] + [labels]
        self._layout, off = [], 0
        for t in tensors:
            nbytes = t.numel() * t.element_size()
            self._layout.append((off, nbytes, t.dtype, tuple(t.shape)))
            off = (off + nbytes + 255) // 256 * 256
        self._nbytes = off
        self._static = torch.empty(off, dtype=torch.uint8, device=dev)
        views = [self._static[o:o + n].view(dt).view(shape) for o, n, dt, shape in self._layout]
        self._views = views
        self.static_feats = dict(zip(self._keys, views[:-1]))
        self.static_labels = views[-1]
        self._bindings = [g.binding for g in getattr(model, "_groups", []) if g.binding is not None]
        self._bindings += [b for b in getattr(getattr(model, "_sharded", None), "bindings", []) if b is not None]
        for b in self._bindings:
            b.enable_device_hyper(dev)
        self._stage(example_batch)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager_step()
        torch.cuda.synchronize(dev)

    def _eager_step(self):
        if self._zero_grad_fn is not None:
            self._zero_grad_fn()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.model.training_step((self.static_feats, self.static_labels), 0)
        if self._backward_fn is not None:
            self._backward_fn(loss)
        else:
            loss.backward()
        self.optimizer.step()
        return loss

    def _stage(self, batch):
        feats, labels = batch
        tensors = [feats[k] for k in self._keys] + [labels]
        for v, t in zip(self._views, tensors):               # async from pinned host memory, D2D otherwise
            v.copy_(t, non_blocking=True)

    def pack(self, batch, device=None) -> torch.Tensor:
        """ONE buffer holding ``batch`` in the layout of the graph's input block.  On the device (``device=None``): a resident
        data set is packed once, and ``step(packed)`` then needs a single device-to-device copy.  ``device='cpu'``: a pinned host
        buffer -- what a collate function that writes its features into one pinned block produces -- so that the upload of a
        batch is ONE host-to-device copy (``prefetch(packed)``) instead of one per feature."""
        feats, labels = batch
        tensors = [feats[k] for k in self._keys] + [labels]
        if device is None:
            packed = torch.empty_like(self._static)
        else:
            packed = torch.empty(self._nbytes, dtype=torch.uint8, device=device, pin_memory=(str(device) == "cpu"))
        for (o, n, dt, shape), t in zip(self._layout, tensors):
            packed[o:o + n].view(dt).view(shape).copy_(t, non_blocking=True)
        return packed

    def prefetch(self, batch) -> None:
        """Start copying ``batch`` (pinned host tensors, or one pinned buffer made by ``pack(batch, 'cpu')``) to the device on a
        copy stream while the current step runs.  The next ``__call__(batch)`` with the same batch object then only does one
        device-to-device move into the graph's input buffers.  What a DataLoader's prefetching does for the reference loop
        (``torchctr/trainer.py:291``), one level down."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staging = torch.empty_like(self._static)
            self._staging_views = [self._staging[o:o + n].view(dt).view(shape) for o, n, dt, shape in self._layout]
        cs = self._copy_stream
        if self._moved is not None:
            cs.wait_event(self._moved)          # only the previous staging -> static move, NOT the step that is running now
        with torch.cuda.stream(cs):
            if torch.is_tensor(batch):
                self._staging.copy_(batch, non_blocking=True)
            else:
                feats, labels = batch
                tensors = [feats[k] for k in self._keys] + [labels]
                for v, t in zip(self._staging_views, tensors):
                    v.copy_(t, non_blocking=True)
        self._prefetched = batch

    _copy_stream = None
    _prefetched = None
    _moved = None

    def __call__(self, batch) -> torch.Tensor:
        if self._prefetched is batch and batch is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._copy_stream)
            self._static.copy_(self._staging, non_blocking=True)    # device to device, ~5 us for a Criteo batch
            if self._moved is None:
                self._moved = torch.cuda.Event()
            self._moved.record()
            self._prefetched = None
        elif torch.is_tensor(batch):                 # a buffer made by pack(): one copy
            self._static.copy_(batch, non_blocking=True)
        else:
            self._stage(batch)
        for b in self._bindings:
            b.advance()
        self.graph.replay()
        return self.static_loss
