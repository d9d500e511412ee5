"""Embedding tables and the fused pooled lookup (host side of kernels K1 / K2).

``EmbeddingTable`` is an ``nn.Embedding`` (same constructor, same ``weight`` parameter, same
state-dict key) so models built from it load and save reference checkpoints
(``torchctr/models/dnn.py:17-23``).  ``PooledLookupGroup`` turns the per-feature Python loop of
``torchctr/models/dnn.py:53-67`` into ONE kernel launch that writes the tower input, and its
backward into sort/dedup + fused sparse update (or sparse gradients when no optimizer is
bound), so neither a ``[B, L, D]`` activation nor a dense ``[V, D]`` gradient exists.
"""
from __future__ import annotations

import os

import warnings

import torch
import torch.nn as nn

from .. import _lib, ops


class EmbeddingTable(nn.Embedding):
    """``nn.Embedding`` whose lookups run through libctr_b200.

    Extra keyword arguments describe how raw ids reach table rows:
    ``index_kind`` 'direct' (ids are rows, the reference's only mode at model level), 'hash'
    (row = murmur3_32(str(id), seed) % num_embeddings, i.e. ``torchctr.utils.hash_bucket`` fused
    into the lookup) or 'vocab' (row = VocabIndex[id], unknown -> 0); ``pooling`` 'sum' | 'mean'.
    ``forward`` keeps ``nn.Embedding`` semantics (un-pooled gather).
    """

    def __init__(self, num_embeddings, embedding_dim, padding_idx=None, max_norm=None, norm_type=2.0,
                 scale_grad_by_freq=False, sparse=False, _weight=None, *, pooling: str = "sum",
                 index_kind: str = "direct", hash_seed: int = 0, vocab=None, use_id_weight: bool = False, **kw):
        if max_norm is not None or scale_grad_by_freq or padding_idx is not None:
            raise NotImplementedError("max_norm / scale_grad_by_freq / padding_idx are not supported by the fused "
                                      "lookup (the reference models never set them, models/dnn.py:21)")
        super().__init__(num_embeddings, embedding_dim, None, None, norm_type, False, sparse, _weight, **kw)
        if pooling not in ("sum", "mean"):
            raise ValueError(f"unknown pooling {pooling!r}")
        if index_kind not in ("direct", "hash", "vocab"):
            raise ValueError(f"unknown index_kind {index_kind!r}")
        self.pooling = pooling
        self.index_kind = index_kind
        self.hash_seed = int(hash_seed)
        self.vocab = vocab                       # VocabIndex module when index_kind == 'vocab'
        self.use_id_weight = bool(use_id_weight)  # multiply rows by feats['<name>_weight'] (dataset.py:59-67)
        self._opt_kind = None

    # -- optimizer state of the fused row update ---------------------------------------------------------------------
    # Where it lives decides how it is checkpointed by the UNMODIFIED reference Trainer (torchctr/trainer.py:353-381 saves
    # model.state_dict() and optimizer.state_dict()):
    #   * the table's weight is one of the torch optimizer's parameters (the reference way: optim.Adagrad(model.parameters()))
    #     -> the state IS the optimizer's own state for that parameter (Adagrad 'sum', Adam 'exp_avg' / 'exp_avg_sq'), so
    #     optimizer.state_dict() carries it in torch's format and a checkpoint moves freely between this model and the
    #     reference's nn.Embedding + torch.optim;
    #   * otherwise (tables kept out of the dense optimizer, row-wise Adagrad, grown tables) -> non-persistent module buffers,
    #     saved / restored with table_optimizer_state_dict() / load_table_optimizer_state_dict().
    # Either way model.state_dict() has exactly the reference's keys.
    _TORCH_STATE_KEYS = {"adagrad": ("sum", None), "adam": ("exp_avg", "exp_avg_sq")}

    def _ensure_state(self, kind: str, initial_accumulator_value: float = 0.0, optimizer=None):
        w = self.weight
        keys = self._TORCH_STATE_KEYS.get(kind)
        self._state_owner = None
        if optimizer is not None and keys is not None and any(w is p for g in optimizer.param_groups for p in g["params"]):
            st = optimizer.state[w]
            for key in keys:
                if key is not None and (key not in st or st[key].shape != w.shape or st[key].device != w.device):
                    st[key] = torch.full_like(w.data, initial_accumulator_value if key == "sum" else 0.0)
            if "step" not in st:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
            self._state_owner = (optimizer, keys)
            self._opt_kind = kind
            return
        store = getattr(self, "_store", None)
        rows = store.shape[0] if store is not None and store.data_ptr() == w.data_ptr() else w.shape[0]
        if kind in ("adagrad", "adam") and getattr(self, "_opt_state0", None) is None:
            self._opt_state0 = torch.full((rows, w.shape[1]), initial_accumulator_value if kind == "adagrad" else 0.0, device=w.device)
        if kind == "rowwise_adagrad" and getattr(self, "_opt_state0", None) is None:
            self._opt_state0 = torch.full((rows,), initial_accumulator_value, device=w.device)
        if kind == "adam" and getattr(self, "_opt_state1", None) is None:
            self._opt_state1 = torch.zeros((rows, w.shape[1]), device=w.device)
        for name in ("_opt_state0", "_opt_state1"):          # follow the table if it grew / moved: sized like its reservation,
            buf = getattr(self, name, None)                   # so that growth inside it costs nothing here either
            if buf is not None and (buf.shape[0] < w.shape[0] or buf.device != w.device):
                new = torch.zeros((rows,) + tuple(buf.shape[1:]), device=w.device)
                n = min(buf.shape[0], w.shape[0])
                new[:n] = buf[:n].to(w.device)
                setattr(self, name, new)
        self._opt_kind = kind

    _state_owner = None
    _opt_state0 = None
    _opt_state1 = None

    def _state(self, which: int):
        owner = self._state_owner
        if owner is not None:
            key = owner[1][which]
            return None if key is None else owner[0].state[self.weight].get(key)
        return self._opt_state0 if which == 0 else self._opt_state1

    @property
    def opt_state0(self):
        """Adagrad sum [V, D] | row-wise accumulator [V] | Adam exp_avg -- wherever it currently lives."""
        return self._state(0)

    @property
    def opt_state1(self):
        """Adam exp_avg_sq [V, D]."""
        return self._state(1)

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        for name in ("_opt_state0", "_opt_state1"):           # module-held state follows .to() / .cuda()
            buf = getattr(self, name, None)
            if buf is not None:
                setattr(self, name, fn(buf))
        return out

    # -- counter-based initialisation: the same table whatever the sharding ------------------------------------------------
    @staticmethod
    def counter_seed(base_seed: int, table_index: int, width_index: int = 0) -> int:
        return (int(base_seed) * 1000003 + width_index * 4099 + table_index) & 0x7FFFFFFFFFFFFFFF

    def counter_init_(self, seed: int, std: float = 1.0):
        """weight[r, c] := N(0, std) drawn from a counter-based generator keyed by (seed, r, c) (``ctr_normal_fill_rows``)
        instead of torch's sequential stream: any rank can then create just ITS rows of the table
        (``ctr_normal_fill_rows_strided``) and the union equals this table -- row-sharded models need no full copy anywhere."""
        with torch.no_grad():
            ops.normal_fill_rows(self.weight.data, 0, self.num_embeddings, 0.0, std, seed)
        return self

    # -- growth (DynamicEmbedding._expand_embeddings, torchctr/nn/embedding.py:69-78) -----------
    def grow_to(self, new_num_embeddings: int, std: float = 0.01) -> None:
        """Append rows ~ N(0, std) up to ``new_num_embeddings``; existing rows keep their values.
        The table sits in a capacity-managed store, so growth is amortised O(new rows)."""
        old = self.num_embeddings
        if new_num_embeddings <= old:
            return
        w = self.weight
        store = getattr(self, "_store", None)
        if store is None or store.data_ptr() != w.data_ptr() or store.shape[0] < new_num_embeddings:
            # reserve once: the configured row capacity (vocab_max_rows, when the table belongs to a growing vocabulary) or
            # 1.5x -- growth inside the reservation neither reallocates nor copies, it only fills the new rows
            cap = max(new_num_embeddings, old + old // 2 + 1024, int(getattr(self, "vocab_max_rows", 0) or 0))
            store = torch.empty(cap, self.embedding_dim, dtype=w.dtype, device=w.device)
            store[:old] = w.data
            self._store = store
        seed = getattr(self, "_init_seed", None)
        if seed is None:
            seed = self._init_seed = torch.initial_seed()
        ops.normal_fill_rows(store, old, new_num_embeddings - old, 0.0, std, seed)
        self.weight = nn.Parameter(store[:new_num_embeddings], requires_grad=w.requires_grad)
        self.num_embeddings = new_num_embeddings

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        key = prefix + "weight"
        if getattr(self, "_store", None) is not None and key in destination and not keep_vars:
            destination[key] = destination[key].clone()      # a view would drag the whole store into torch.save

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for which, name in enumerate(("opt_state0", "opt_state1")):   # round-1 checkpoints carried the state as buffers
            key = prefix + name
            if key in state_dict:
                setattr(self, "_" + name, state_dict.pop(key).to(self.weight.device).clone())
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        """``nn.Embedding.forward``: ids of any shape (all >= 0) -> [..., D]."""
        flat = input.reshape(-1, 1)
        out = pooled_lookup([(self, flat, None)])
        return out[:, : self.embedding_dim].reshape(*input.shape, self.embedding_dim)


class SparseOptimizerBinding:
    """Reads the hyper-parameters of a ``torch.optim`` optimizer every step, so the fused row
    update follows the user's optimizer (and any lr scheduler) set up for the reference Trainer
    (``torchctr/trainer.py:303``)."""

    def __init__(self, optimizer: torch.optim.Optimizer, tables, kind: str | None = None):
        self.optimizer = optimizer
        self.tables = list(tables)
        self.step = 0
        ids = {id(t.weight) for t in self.tables}
        self.group = None
        for g in optimizer.param_groups:
            if any(id(p) in ids for p in g["params"]):
                self.group = g
                break
        if self.group is None:
            self.group = optimizer.param_groups[0]
        if kind is None:
            name = type(optimizer).__name__.lower()
            if name == "sgd":
                kind = "sgd"
            elif name == "adagrad":
                kind = "adagrad"
            elif name in ("sparseadam", "adam", "adamw"):
                kind = "adam"
                if name != "sparseadam":
                    warnings.warn("torchctr_b200: dense Adam on embedding tables is applied as lazy Adam "
                                  "(torch.optim.SparseAdam semantics: untouched rows do not decay)", stacklevel=3)
            else:
                raise ValueError(f"no fused row update for optimizer {type(optimizer).__name__}; "
                                 "pass kind='sgd'|'adagrad'|'rowwise_adagrad'|'adam'")
        self.kind = kind
        g = self.group
        if g.get("weight_decay", 0) or g.get("momentum", 0):
            raise ValueError("fused table update supports neither weight_decay nor momentum "
                             "(a touched-rows-only update cannot reproduce them)")

    def _opt_for_step(self, step: int, device_hyper=None) -> _lib.Opt:
        g = self.group
        lr = float(g["lr"])
        if self.kind in ("adagrad", "rowwise_adagrad"):
            lr = lr / (1.0 + (step - 1) * float(g.get("lr_decay", 0.0)))
        eps = float(g.get("eps", 1e-10 if self.kind != "adam" else 1e-8))
        betas = g.get("betas", (0.9, 0.999))
        return ops.make_opt(self.kind, lr, eps, (float(betas[0]), float(betas[1])), step, device_hyper)

    def next_opt(self) -> _lib.Opt:
        """Hyper-parameters of the next fused update.  Eager mode: passed by value.  Graph mode
        (``enable_device_hyper``): the kernels read them from a device tensor that ``advance``
        refreshes before every replay, so lr schedules and Adam's step keep working."""
        if self._hyper_dev is not None:
            if torch.cuda.is_current_stream_capturing():
                return self._opt_for_step(max(self.step, 1), self._hyper_dev)
            self.advance()
            return self._opt_for_step(self.step, self._hyper_dev)
        self.step += 1
        return self._opt_for_step(self.step)

    _hyper_dev = None
    _hyper_last = None

    def enable_device_hyper(self, device):
        if self._hyper_dev is None:
            self._hyper_dev = torch.zeros(5, dtype=torch.float32, device=device)
        return self

    def advance(self):
        """One optimizer step further: refresh the device copy of the scalars if they changed."""
        self.step += 1
        vals = ops.opt_hyper(self._opt_for_step(self.step))
        if vals != self._hyper_last:
            self._hyper_dev.copy_(torch.tensor(vals, dtype=torch.float32).pin_memory(), non_blocking=True)
            self._hyper_last = vals

    def initial_accumulator_value(self) -> float:
        return float(self.group.get("initial_accumulator_value", 0.0))


class _Workspace:
    """Grow-only device scratch shared by the groups of one device."""
    _bufs: dict = {}

    @classmethod
    def get(cls, device, nbytes: int) -> torch.Tensor:
        buf = cls._bufs.get(device)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                cls._retired.append(buf)      # a captured CUDA graph may still point into it
            buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
            cls._bufs[device] = buf
        return buf

    _retired: list = []


_side_streams: dict = {}
PLAN_START = os.environ.get("CTR_PLAN_START", "after")      # where the backward plan (the sort) is issued: before | after | head
_late_starts: list = []


def run_late_starts() -> None:
    """Issue the backward plans whose start was left for later (PLAN_START = "head": the fused logit head calls this)."""
    while _late_starts:
        link = _late_starts.pop()
        fn, link.late = link.late, None
        if fn is not None:
            fn()


def _side_stream(device) -> torch.cuda.Stream:
    key = torch.device(device)
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=key)
    return st


# Off by default: measured on cfg2 (B200), the sweep takes 163.8 us with the blocked handover and 164.7 us with the row-major
# one -- its long-scoreboard stalls are dependent-latency chains (gather -> reduce -> row read-modify-write), not DRAM
# efficiency of the gathers, so where the slices come from does not matter.  Kept as a tested option (CTR_BLOCKED_GRAD=1).
BLOCKED_GRAD = os.environ.get("CTR_BLOCKED_GRAD", "0") == "1"


class BlockedGrad:
    """dL/d(pooled output) handed over COLUMN-BLOCKED (``ctr_group_t.grad_blocked``): feature j's gradient slices are the
    contiguous matrix ``buffer[j * B * D:]`` viewed [B, D] instead of D-float pieces strided over [B, stride].  The sweep then
    gathers a table's slices from one dense 4 D B-byte block that sits in L2 while that table's rows are swept, instead of
    random 64-byte DRAM reads over the whole gradient matrix.

    Protocol: the lookup offers it (``take_blocked_offer``) to the ONE consumer of its output, the tower's first block, as
    that block's ``gx_provider``; the block's input-gradient GEMM writes the blocked layout (``ctr_linear_fwd_blocked``), sets
    ``marked`` and returns ``buffer`` viewed [B, stride] as the gradient autograd carries back; the lookup's backward checks
    that what arrives IS that buffer and reads it blocked.  Anything else that touched the gradient on the way would have read
    a scrambled matrix, hence the check -- and the opt-in by the model (``CTRModelBase._run_tower(blocked_ok=True)``)."""

    def __init__(self, B: int, stride: int, cols: int, block: int, device):
        self.B, self.stride, self.cols, self.block, self.device = B, stride, cols, block, device
        self.buffer = None
        self.marked = False

    def __call__(self):
        if self.buffer is None:
            self.buffer = torch.empty(self.B * self.stride, dtype=torch.float32, device=self.device)
        return self


blocked_backwards = 0      # how many lookups read their gradient column-blocked (tests, diagnostics)
_blocked_offer = None      # (data_ptr of the lookup output, BlockedGrad) of the most recent eligible lookup


def take_blocked_offer(x: torch.Tensor):
    """The ``BlockedGrad`` of the lookup that produced ``x`` (only if ``x`` IS that lookup's output), or None."""
    global _blocked_offer
    offer, _blocked_offer = _blocked_offer, None
    if offer is not None and offer[0] == x.data_ptr() and tuple(x.shape) == (offer[1].B, offer[1].stride):
        return offer[1]
    return None


class _PooledLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, call, dense, *weights):
        ctx.call = call
        ctx.has_dense = dense is not None
        call.dense_needs_grad = bool(ctx.needs_input_grad[1])
        return call.run_forward(dense, weights)

    @staticmethod
    def backward(ctx, grad_out):
        call = ctx.call
        wgrads = call.run_backward(grad_out)
        gdense = None
        if ctx.has_dense and ctx.needs_input_grad[1]:
            gdense = grad_out[:, call.dense_col: call.dense_col + call.dense_width]
        return (None, gdense, *wgrads)


class _PooledLookupExtraFn(torch.autograd.Function):
    """The lookup with the fused per-bag scalar (twin tables + FM term): two outputs, (pooled [B, stride], extra [B])."""

    @staticmethod
    def forward(ctx, call, dense, *weights):
        ctx.call = call
        ctx.has_dense = dense is not None
        call.dense_needs_grad = bool(ctx.needs_input_grad[1])
        out = call.run_forward(dense, weights)
        # hand `extra` out WITHOUT keeping it on the call: an output of this node that the node's own ctx also held would be
        # a reference cycle, and the autograd graph (AccumulateGrad nodes bound to the stream of that step included) would
        # outlive the step -- which breaks the next CUDA-graph capture
        extra, call.extra = call.extra, None
        ctx.device = out.device
        return out, extra

    @staticmethod
    def backward(ctx, grad_out, grad_extra):
        call = ctx.call
        B = call.B
        if grad_out is None:
            grad_out = torch.zeros(B, call.stride, dtype=torch.float32, device=ctx.device)
        if grad_extra is None:
            grad_extra = torch.zeros(B, dtype=torch.float32, device=ctx.device)
        wgrads = call.run_backward(grad_out, grad_extra.reshape(-1).contiguous())
        gdense = None
        if ctx.has_dense and ctx.needs_input_grad[1]:
            gdense = grad_out[:, call.dense_col: call.dense_col + call.dense_width]
        return (None, gdense, *wgrads)


class _LookupCall:
    """One forward/backward of a group of tables over one batch."""

    def __init__(self, entries, dense_width, layout, binding, training, plan_link=None, twins=None, fm=False):
        # entries: [(table module, ids [B, L] i64, id_weight or None)]
        self.entries = entries
        self.twins = twins              # one-column twin tables (same ids) or None; fm: add the FM term to `extra`
        self.fm = bool(fm)
        self.extra = None               # f32 [B]: sum of the twins (+ FM term), written by the lookup kernel
        self.fm_sum = None
        self.plan_link = plan_link
        self.layout = layout            # (out_cols, width, stride, dense_col)
        self.out_cols, self.width, self.stride, self.dense_col = layout
        self.dense_width = dense_width
        self.binding = binding
        self.training = training
        self.B = entries[0][1].shape[0] if entries else 0
        self.bag_scales = None
        self.status = None
        self.grad_enabled = torch.is_grad_enabled()   # read outside the autograd Function (inside, grad mode is off)
        self.dense_needs_grad = False
        self.blocked = None             # BlockedGrad offered to the consumer of the output (see take_blocked_offer)

    def _specs(self, tables_data, with_state, twin_data=None):
        specs = []
        for i, (mod, ids, wgt) in enumerate(self.entries):
            vocab = mod.vocab.handle() if mod.index_kind == "vocab" else None
            tw = {}
            if twin_data is not None:
                tm = self.twins[i]
                tw = dict(twin_table=twin_data[i],
                          twin_state0=getattr(tm, "opt_state0", None) if with_state else None,
                          twin_state1=getattr(tm, "opt_state1", None) if with_state else None)
            specs.append(ops.FeatureSpec(**tw, 
                ids=ids, table=tables_data[i], num_rows=mod.num_embeddings, D=mod.embedding_dim,
                out_col=self.out_cols[i], pooling=mod.pooling, index_kind=mod.index_kind, hash_seed=mod.hash_seed,
                id_weight=wgt, vocab=vocab,
                state0=getattr(mod, "opt_state0", None) if with_state else None,
                state1=getattr(mod, "opt_state1", None) if with_state else None,
                bag_scale=self.bag_scales[i]))
        return specs

    def _single_id_one_width(self) -> bool:
        """the group the compile-time-layout sweep takes: single-id bags, one width of 16 / 32 / 64, sum pooling, no weights"""
        D = self.entries[0][0].embedding_dim
        return D in (16, 32, 64) and all(m.embedding_dim == D and m.pooling == "sum" and wgt is None and ids.shape[1] == 1
                                         for m, ids, wgt in self.entries)

    def run_forward(self, dense, weights):
        with torch.cuda.device(weights[0].device):       # the model may live on a device that is not the current one
            return self._run_forward(dense, weights)

    def run_backward(self, grad_out, grad_extra=None):
        with torch.cuda.device(grad_out.device):
            return self._run_backward(grad_out, grad_extra)

    def _run_forward(self, dense, weights):
        dev = weights[0].device
        B = self.B
        out = torch.empty(B, self.stride, dtype=torch.float32, device=dev)
        self.bag_scales = [torch.empty(B, dtype=torch.float32, device=dev) if m.pooling == "mean" else None
                           for m, _, _ in self.entries]
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        zero_from = self.width if self.stride > self.width else -1
        n = len(self.entries)
        twin_data = [w.detach() for w in weights[n:]] if self.twins is not None else None
        if self.twins is not None or self.fm:
            self.extra = torch.empty(B, dtype=torch.float32, device=dev)
            if self.fm:
                self.fm_sum = torch.empty(B, self.entries[0][0].embedding_dim, dtype=torch.float32, device=dev)
        call = ops.make_group(self._specs([w.detach() for w in weights[:n]], False, twin_data), B, out, self.stride,
                              dense=dense, dense_col=self.dense_col, zero_from=zero_from, status=self.status,
                              extra=self.extra, fm_sum=self.fm_sum, fm=self.fm)
        link = self.plan_link
        early = None
        if link is not None and self.training:               # every linked forward runs before any backward
            need = ops.emb_bwd_workspace_bytes(call)
            link.nbytes = max(link.nbytes, need)
            if link.ws is None and self.binding is not None and self.grad_enabled:
                # The sort depends on the ids only: it runs on a side stream, next to the lookup / the tower (in a captured
                # CUDA graph: a parallel branch).  It is a chain of small latency-bound kernels that leaves most of the
                # machine idle -- but its blocks do hold SM slots the tower's GEMM CTAs (~200 KB of shared memory each) wait
                # for, so WHERE it starts matters (PLAN_START): "before" the lookup kernel, "after" it, or at the logit "head".
                link.ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)

                def early(link=link, call=call):
                    side = _side_stream(dev)
                    if not torch.cuda.is_current_stream_capturing():
                        link.ws.record_stream(side)      # a forward without backward must not hand the block back too early
                        for _, ids, _ in self.entries:
                            ids.record_stream(side)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        ops.emb_bwd_plan(call, link.ws, runs=False)
                    link.pending = side
        if early is not None and PLAN_START == "before":
            early()
        ops.emb_pool_fwd(call)
        global _blocked_offer
        _blocked_offer = None
        if (BLOCKED_GRAD and self.training and self.grad_enabled and self.binding is not None and not self.dense_needs_grad
                and B > 128 and self._single_id_one_width()):
            D = self.entries[0][0].embedding_dim
            self.blocked = BlockedGrad(B, self.stride, len(self.entries) * D, D, dev)
            _blocked_offer = (out.data_ptr(), self.blocked)
        if early is not None and PLAN_START == "head":
            link.late = early
            _late_starts.append(link)
        elif early is not None and PLAN_START != "before":
            early()
        return out

    def _run_backward(self, grad_out, grad_extra=None):
        mods = [m for m, _, _ in self.entries]
        needs = [m.weight.requires_grad for m in mods]
        nw = len(mods) + (len(self.twins) if self.twins is not None else 0)
        if not any(needs):
            return [None] * nw
        if grad_out.stride(1) != 1 or grad_out.stride(0) != grad_out.shape[1]:
            grad_out = grad_out.contiguous()
        dev = grad_out.device
        fused = self.binding is not None and self.training
        if (self.twins is not None or self.fm) and not fused:
            raise RuntimeError("the fused twin / FM lookup needs a bound optimizer in training (bind_optimizer)")
        if fused:
            for m in mods + (list(self.twins) if self.twins is not None else []):
                m._ensure_state(self.binding.kind, self.binding.initial_accumulator_value(), self.binding.optimizer)
        tables = [m.weight.data for m in mods]
        twin_data = [t.weight.data for t in self.twins] if self.twins is not None else None
        blocked = self.blocked is not None and self.blocked.marked
        if blocked:
            global blocked_backwards
            blocked_backwards += 1
        if blocked and (not fused or grad_out.data_ptr() != self.blocked.buffer.data_ptr()):
            raise RuntimeError("the tower wrote dL/dx column-blocked (BlockedGrad) but a different tensor came back to the lookup: "
                               "something else consumed the lookup's output; set CTR_BLOCKED_GRAD=0")
        call = ops.make_group(self._specs(tables, fused, twin_data), self.B, grad_out, grad_out.shape[1],
                              extra=grad_extra, fm_sum=self.fm_sum, fm=self.fm, grad_blocked=blocked)
        link = self.plan_link
        if link is None:
            ws = _Workspace.get(dev, ops.emb_bwd_workspace_bytes(call))
            ops.emb_bwd_plan(call, ws, runs=not fused)
        else:
            # the sort / run list depends on the ids only: tables that share them (DeepFM's first-order
            # weights next to its embeddings) share one plan per step
            need = ops.emb_bwd_workspace_bytes(call)
            link.nbytes = max(link.nbytes, need)
            if link.late is not None:                         # nobody started the sort yet (PLAN_START = "head", no fused head)
                run_late_starts()
            if link.pending is not None:                      # the early sort of run_forward
                torch.cuda.current_stream(dev).wait_stream(link.pending)
                link.pending = None
            if link.ws is None or link.ws.numel() < need or (not fused and not link.has_runs):
                link.ws = torch.empty(max(link.nbytes, need) + 256, dtype=torch.uint8, device=dev)
                ops.emb_bwd_plan(call, link.ws, runs=not fused)
                link.has_runs = not fused
            ws = link.ws
        if fused:
            ops.emb_bwd_apply(call, ws, self.binding.next_opt())
            return [None] * nw
        # no optimizer bound: hand autograd sparse gradients (torch.optim SGD / Adagrad / SparseAdam accept them)
        S = sum(ids.numel() for _, ids, _ in self.entries)
        dmax = max(m.embedding_dim for m in mods)
        uf = torch.empty(S, dtype=torch.int32, device=dev)
        ur = torch.empty(S, dtype=torch.int32, device=dev)
        rg = torch.empty(S, dmax, dtype=torch.float32, device=dev)
        nu = torch.zeros(1, dtype=torch.int64, device=dev)
        ops.emb_bwd_apply(call, ws, ops.make_opt("none"), uf, ur, rg, nu)
        U = int(nu.item())
        uf, ur, rg = uf[:U], ur[:U], rg[:U]
        bounds = torch.searchsorted(uf, torch.arange(len(mods) + 1, dtype=torch.int32, device=dev)).tolist()
        grads = []
        for i, m in enumerate(mods):
            if not needs[i]:
                grads.append(None)
                continue
            lo, hi = bounds[i], bounds[i + 1]
            grads.append(torch.sparse_coo_tensor(ur[lo:hi].long().unsqueeze(0), rg[lo:hi, : m.embedding_dim],
                                                 size=tuple(m.weight.shape), is_coalesced=True))
        return grads


def _layout(entries, dense_width, align=4):
    cols, col = [], 0
    for m, _, _ in entries:
        cols.append(col)
        col += m.embedding_dim
    dense_col = col
    width = col + dense_width
    stride = (width + align - 1) // align * align
    return cols, width, stride, dense_col


class PlanLink:
    """Shared by the lookups of ONE step whose ids, index mapping and table sizes are identical: the
    first backward builds the sort / run plan, the others reuse it.  Create one per forward."""

    def __init__(self, nbytes: int = 0):
        self.ws = None
        self.nbytes = nbytes
        self.pending = None      # side stream an early sort is running on
        self.late = None         # the launch of that sort, when it has been left to run_late_starts()
        self.has_runs = False    # the plan in ws lists the runs (needed for unique-row outputs only)


def fused_extra_eligible(tables, twins, feats_L, binding, training: bool) -> bool:
    """Can the twin tables / FM term ride inside the lookup and update kernels (``ctr_group_t.extra``)?  Single-id bags,
    one width of 16 / 32 / 64, sum pooling, no per-id weights, direct / hashed / vocabulary ids alike -- and, when a
    gradient will be asked for, a bound optimizer (the fused update is the only consumer of the twin gradients)."""
    if len(tables) == 0 or len(tables) > _lib.MAX_FEATURES:
        return False
    D = tables[0].embedding_dim
    if D not in (16, 32, 64) or any(L != 1 for L in feats_L):
        return False
    for t in tables:
        if t.embedding_dim != D or t.pooling != "sum" or t.use_id_weight:
            return False
    if twins is not None:
        for t, tw in zip(tables, twins):
            if tw.embedding_dim != 1 or tw.num_embeddings != t.num_embeddings or tw.weight.device != t.weight.device:
                return False
    if training and torch.is_grad_enabled() and binding is None:
        return False
    return True


def pooled_lookup(entries, dense: torch.Tensor | None = None, binding: SparseOptimizerBinding | None = None,
                  training: bool = False, plan_link: PlanLink | None = None, twins=None, fm: bool = False):
    """Pools every (table, ids [B, L], id_weight) entry and concatenates them with ``dense``.

    Returns f32 ``[B, stride]`` with ``stride`` = total width rounded up to 4 floats; columns past
    the width are zero.  This is ``torchctr/models/dnn.py:53-67`` as one launch.

    ``twins`` (one ``EmbeddingTable(V, 1)`` per entry, indexed by the same ids) and / or ``fm=True`` ask for the fused
    per-bag scalar of ``ctr_group_t.extra``: the call then returns ``(pooled, extra [B])`` with
    ``extra[b] = sum_f twin_f[id] + (FM second-order term of the pooled vectors if fm)`` -- DeepFM's logit terms outside
    the tower -- and backward folds their gradients into the one fused update (``fused_extra_eligible`` says when).
    """
    if not entries:
        raise ValueError("pooled_lookup needs at least one table")
    dev = entries[0][0].weight.device
    if dev.type != "cuda":
        raise RuntimeError("torchctr_b200 tables must live on a CUDA device: the lookup kernels have no CPU path")
    prepared = []
    for mod, ids, wgt in entries:
        if ids.dim() == 1:
            ids = ids.unsqueeze(1)
        if ids.dim() != 2:
            raise ValueError(f"ids must be [B] or [B, L], got {tuple(ids.shape)}")
        if ids.dtype != torch.int64:
            ids = ids.long()
        ids = ids.to(dev, non_blocking=True).contiguous()
        if wgt is not None:
            wgt = wgt.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        prepared.append((mod, ids, wgt))
    dense_width = 0
    if dense is not None:
        dense = dense.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        dense_width = dense.shape[1]
    # a launch group carries at most MAX_FEATURES tables and a 32-bit key space
    if len(prepared) <= _lib.MAX_FEATURES and sum(m.num_embeddings for m, _, _ in prepared) < 2 ** 32 - 1:
        if twins is not None or fm:
            twins = list(twins) if twins else None
            if twins is not None and len(twins) != len(prepared):
                raise ValueError("one twin table per entry")
            call = _LookupCall(prepared, dense_width, _layout(prepared, dense_width), binding, training, plan_link,
                               twins=twins, fm=fm)
            weights = [m.weight for m, _, _ in prepared] + [t.weight for t in (twins or [])]
            return _PooledLookupExtraFn.apply(call, dense, *weights)
        call = _LookupCall(prepared, dense_width, _layout(prepared, dense_width), binding, training, plan_link)
        return _PooledLookupFn.apply(call, dense, *[m.weight for m, _, _ in prepared])
    raise NotImplementedError("split the features into several pooled_lookup calls "
                              f"(more than {_lib.MAX_FEATURES} tables or >= 2^32 rows in one group)")


def discover_optimizer(weights):
    """The live ``torch.optim.Optimizer`` that holds any of ``weights`` among its parameters, or None.  The reference loop
    hands the optimizer to the Trainer, never to the model (``torchctr/trainer.py:28-31``), so a model dropped into it is
    not told which optimizer steps it; one scan of the garbage collector's objects at the first training step finds out."""
    import gc
    ids = {id(w) for w in weights}
    for obj in gc.get_objects():
        try:
            if isinstance(obj, torch.optim.Optimizer) and any(id(p) in ids for g in obj.param_groups for p in g["params"]):
                return obj
        except ReferenceError:
            continue
    return None


class PooledLookupGroup:
    """The sparse half of a model: tables in ``feat_configs`` order + the dense block."""

    def __init__(self, names, tables: nn.ModuleDict):
        self.names = list(names)
        self.tables = tables
        self.binding: SparseOptimizerBinding | None = None
        self._autobind_tried = False

    def autobind(self) -> None:
        """Drop-in behaviour under the unmodified reference Trainer: no ``bind_optimizer`` call was made, so look for the
        optimizer that steps these tables and fuse their update into backward when it is one the fused update can follow
        (SGD / Adagrad / Adam without weight decay or momentum).  Otherwise the tables keep handing autograd sparse
        gradients, which torch's SGD / Adagrad / SparseAdam accept."""
        if self.binding is not None or self._autobind_tried:
            return
        self._autobind_tried = True
        opt = discover_optimizer([self.tables[n].weight for n in self.names])
        if opt is None:
            return
        try:
            self.bind_optimizer(opt)
        except ValueError as e:
            warnings.warn(f"torchctr_b200: table update not fused ({e}); backward will emit sparse gradients", stacklevel=3)

    def bind_optimizer(self, optimizer, kind: str | None = None):
        self.binding = SparseOptimizerBinding(optimizer, [self.tables[n] for n in self.names], kind)
        return self.binding

    def __call__(self, feats: dict, dense: torch.Tensor | None, training: bool, plan_link: PlanLink | None = None,
                 twins=None, fm: bool = False):
        if training and self.binding is None and torch.is_grad_enabled():
            self.autobind()
        entries = [(self.tables[n], feats[n], feats.get(n + "_weight") if self.tables[n].use_id_weight else None)
                   for n in self.names]
        return pooled_lookup(entries, dense, self.binding, training, plan_link, twins=twins, fm=fm)

    def fused_extra_eligible(self, feats: dict, twins, training: bool) -> bool:
        Ls = [feats[n].shape[1] if feats[n].dim() == 2 else 1 for n in self.names]
        return fused_extra_eligible([self.tables[n] for n in self.names], twins, Ls, self.binding, training)
