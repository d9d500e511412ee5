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
        # the wire layout: the same block with every i64 tensor (ids) narrowed to i32 -- what pack(..., ids="i32") writes and what
        # crosses PCIe; _segments are the (merged) pieces the device widens / copies into the static block
        self._wire, woff = [], 0
        for (o, n, dt, shape) in self._layout:
            wn = n // 2 if dt == torch.int64 else n
            self._wire.append((woff, wn, torch.int32 if dt == torch.int64 else dt, shape))
            woff = (woff + wn + 255) // 256 * 256
        self._wire_nbytes = woff
        self._segments = []
        for (o, n, dt, _), (wo, wn, wdt, _) in zip(self._layout, self._wire):
            narrow = dt == torch.int64
            last = self._segments[-1] if self._segments else None
            if last is not None and last[0] == narrow and last[1] + last[2] == o and last[3] + last[4] == wo:
                last[2] += n
                last[4] += wn
            else:
                self._segments.append([narrow, o, n, wo, wn])
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

    def pack(self, batch, device=None, ids: str = "i64") -> torch.Tensor:
        """ONE buffer holding ``batch`` in the layout of the graph's input block.  On the device (``device=None``): a resident
        data set is packed once, and ``step(packed)`` then needs a single device-to-device copy.  ``device='cpu'``: a pinned host
        buffer -- what a collate function that writes its features into one pinned block produces -- so that the upload of a
        batch is ONE host-to-device copy (``prefetch(packed)``) instead of one per feature.

        ``ids="i32"``: the wire layout -- every i64 tensor of the batch (the ids; the reference collate emits i64,
        ``torchctr/dataset.py:52-57,69``) is stored as i32, which halves the id bytes that cross PCIe; the device widens them back
        into the i64 input block the kernels read (one conversion kernel per run of id tensors).  Raises ``ValueError`` when an
        id does not fit in 32 bits (padding, -100, does)."""
        feats, labels = batch
        tensors = [feats[k] for k in self._keys] + [labels]
        if ids not in ("i64", "i32"):
            raise ValueError("ids must be 'i64' or 'i32'")
        wire = ids == "i32"
        layout = self._wire if wire else self._layout
        nbytes = self._wire_nbytes if wire else self._nbytes
        if device is None:
            packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        else:
            packed = torch.empty(nbytes, dtype=torch.uint8, device=device, pin_memory=(str(device) == "cpu"))
        for (o, n, dt, shape), t in zip(layout, tensors):
            if wire and t.dtype == torch.int64 and t.numel():
                lo, hi = int(t.min()), int(t.max())
                if lo < -2 ** 31 or hi >= 2 ** 31:
                    raise ValueError(f"ids in [{lo}, {hi}] do not fit the i32 wire layout; pack with ids='i64'")
            packed[o:o + n].view(dt).view(shape).copy_(t, non_blocking=True)
        return packed

    def _is_wire(self, t) -> bool:
        return t.numel() == self._wire_nbytes and self._wire_nbytes != self._nbytes

    def _widen(self, wire_block):
        """device wire block -> the static input block: id runs i32 -> i64, everything else byte for byte"""
        for narrow, o, n, wo, wn in self._segments:
            if narrow:
                self._static[o:o + n].view(torch.int64).copy_(wire_block[wo:wo + wn].view(torch.int32), non_blocking=True)
            else:
                self._static[o:o + n].copy_(wire_block[wo:wo + wn], non_blocking=True)

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
            if torch.is_tensor(batch) and self._is_wire(batch):
                self._staging[:self._wire_nbytes].copy_(batch, non_blocking=True)
            elif torch.is_tensor(batch):
                self._staging.copy_(batch, non_blocking=True)
            else:
                feats, labels = batch
                tensors = [feats[k] for k in self._keys] + [labels]
                for v, t in zip(self._staging_views, tensors):
                    v.copy_(t, non_blocking=True)
        self._prefetched = batch

    _copy_stream = None
    _wire_dev = None
    _prefetched = None
    _moved = None

    def __call__(self, batch) -> torch.Tensor:
        if self._prefetched is batch and batch is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._copy_stream)
            if torch.is_tensor(batch) and self._is_wire(batch):
                self._widen(self._staging)
            else:
                self._static.copy_(self._staging, non_blocking=True)    # device to device, ~5 us for a Criteo batch
            if self._moved is None:
                self._moved = torch.cuda.Event()
            self._moved.record()
            self._prefetched = None
        elif torch.is_tensor(batch) and self._is_wire(batch):      # a wire block made by pack(..., ids="i32")
            if batch.device != self.device:
                if self._wire_dev is None:
                    self._wire_dev = torch.empty(self._wire_nbytes, dtype=torch.uint8, device=self.device)
                self._wire_dev.copy_(batch, non_blocking=True)
                batch = self._wire_dev
            self._widen(batch)
        elif torch.is_tensor(batch):                 # a buffer made by pack(): one copy
            self._static.copy_(batch, non_blocking=True)
        else:
            self._stage(batch)
        for b in self._bindings:
            b.advance()
        self.graph.replay()
        return self.static_loss
