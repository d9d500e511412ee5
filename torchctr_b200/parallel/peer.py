"""Row-sharded embedding tables over peer memory: NVLink P2P loads inside the kernels, no all-to-all.

The reference is replicas-only (``Accelerator().prepare`` -> DDP, ``torchctr/trainer.py:128-130``): every rank
holds every table and dense ``[V, D]`` gradients are all-reduced.  Here one process per GPU holds ``1/P`` of every
table -- row ``r`` of table ``f`` lives on rank ``(r + f) mod P`` (the ``+ f`` spreads the hottest row of every
table, id 0 under a Zipf law, over the ranks) -- the batch stays data-parallel and a step is

  forward   ONE lookup kernel per width that reads every row straight from its owner's shard through the peer
            mapping (``ctr_emb_pool_fwd_sharded``); the slots are also bucketed by owner into this rank's routing
            lists (``ctr_route_p2p_build``), which live in peer-visible memory;
  backward  the gradient w.r.t. the pooled output is copied into a peer-visible buffer; barrier; every OWNER pulls
            the (row, slot) lists addressed to it from all ranks, sorts them (``ctr_emb_bwd_plan_p2p``) and runs the
            fused reduce + optimizer sweep, reading each slot's gradient from the rank that produced it
            (``ctr_emb_bwd_apply_p2p``); barrier.

No host read happens anywhere (pair counts stay on the device), so the whole step can be captured into a CUDA
graph together with the NCCL all-reduce of the replicated tower.  The two barriers are tiny NCCL all-reduces on the
compute stream (``IpcTransport``); the tests run the ranks as threads of one process on one GPU
(``ThreadTransport``), which exercises every kernel of the path without a second device.
"""
from __future__ import annotations

import ctypes as C
import threading

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import _lib, ops
from ..nn.embedding import SparseOptimizerBinding

_TYPESTR = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}


class PeerBuffer:
    """Device memory other processes can map (``ctr_peer_alloc``), viewed as torch tensors."""

    def __init__(self, nbytes: int, device):
        self.device = torch.device(device)
        self.nbytes = max(int(nbytes), 256)
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctr_peer_alloc(self.nbytes, C.byref(ptr)), "ctr_peer_alloc")
        self.ptr = int(ptr.value)

    def handle(self) -> bytes:
        buf = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctr_peer_export(self.ptr, buf), "ctr_peer_export")
        return buf.raw

    def tensor(self, dtype, shape, offset_bytes: int = 0) -> torch.Tensor:
        """A torch view of the buffer (it keeps this object alive)."""
        view = _ArrayView(self, self.ptr + offset_bytes, tuple(int(s) for s in shape), _TYPESTR[dtype])
        return torch.as_tensor(view, device=self.device)

    def release(self):
        """Free the memory.  Only call once no other rank maps it any more (the ranks agree on that themselves, e.g.
        with a barrier); dropping the object does NOT free it, because cudaFree of memory a peer still has open is
        undefined and process exit releases it anyway."""
        if getattr(self, "ptr", 0):
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().ctr_peer_free(self.ptr), "ctr_peer_free")
            self.ptr = 0


class _ArrayView:
    def __init__(self, owner, ptr, shape, typestr):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2,
                                         "strides": None}


class IpcTransport:
    """One process per GPU: buffers are exchanged as CUDA IPC handles through ``torch.distributed``."""

    def __init__(self, pg=None, device=None):
        self.pg = pg
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        self.device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._opened = []

    def share(self, buf: PeerBuffer):
        """Collective: every rank passes its buffer, gets the device pointer of every rank's buffer."""
        handles = [None] * self.world
        dist.all_gather_object(handles, buf.handle(), group=self.pg)
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(buf.ptr)
                continue
            p = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().ctr_peer_open(C.create_string_buffer(h, len(h)), C.byref(p)), "ctr_peer_open")
            self._opened.append(int(p.value))
            ptrs.append(int(p.value))
        return ptrs

    def barrier(self):
        """After this point of the stream every rank's earlier work on its stream is complete (and visible
        through the peer mappings).  A 4-byte NCCL all-reduce: capturable into a CUDA graph."""
        dist.all_reduce(self._token, group=self.pg)

    def all_gather_object(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.pg)
        return out

    def all_reduce(self, t: torch.Tensor) -> None:
        """In-place sum over the ranks (NCCL, on the current stream; capturable)."""
        dist.all_reduce(t, group=self.pg)

    def all_gather_tensor(self, t: torch.Tensor) -> torch.Tensor:
        """[n, ...] on every rank (same shape) -> [world * n, ...] in rank order (NCCL all-gather on the current stream)."""
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.pg)
        return out


class ThreadTransport:
    """Test transport: the ranks are threads of one process on ONE device; 'peer' pointers are plain device pointers."""

    class Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = {}
            self.lock = threading.Lock()

    def __init__(self, shared: "ThreadTransport.Shared", rank: int, device):
        self.shared, self.world, self.rank, self.device = shared, shared.world, rank, torch.device(device)
        self._seq = 0

    def share(self, buf: PeerBuffer):
        key = self._seq
        self._seq += 1
        with self.shared.lock:
            self.shared.slots.setdefault(key, {})[self.rank] = buf.ptr
        self.shared.barrier.wait()
        ptrs = [self.shared.slots[key][r] for r in range(self.world)]
        self.shared.barrier.wait()
        return ptrs

    def barrier(self):
        torch.cuda.synchronize(self.device)
        self.shared.barrier.wait()

    def all_gather_tensor(self, t: torch.Tensor) -> torch.Tensor:
        torch.cuda.synchronize(self.device)
        return torch.cat(self.all_gather_object(t), 0)

    def all_reduce(self, t: torch.Tensor) -> None:
        torch.cuda.synchronize(self.device)
        parts = self.all_gather_object(t)
        total = parts[0].clone()
        for p in parts[1:]:                 # rank order on every rank: all replicas get bit-identical sums
            total += p
        torch.cuda.synchronize(self.device)
        self.shared.barrier.wait()          # everybody has read everybody's input
        t.copy_(total)
        torch.cuda.synchronize(self.device)

    def all_gather_object(self, obj):
        key = ("obj", self._seq)
        self._seq += 1
        with self.shared.lock:
            self.shared.slots.setdefault(key, {})[self.rank] = obj
        self.shared.barrier.wait()
        out = [self.shared.slots[key][r] for r in range(self.world)]
        self.shared.barrier.wait()
        return out


def owned_rows(num_rows: int, table: int, rank: int, world: int):
    """(first row, count) of the rows of table ``table`` that rank ``rank`` owns: rows r with (r + table) % world == rank."""
    first = (rank - table) % world
    count = 0 if first >= num_rows else (num_rows - first + world - 1) // world
    return first, count


def shard_geometry(num_rows_list, world: int):
    """base[o][f] = first row of table f inside owner o's fused shard (every table gets at least one row so that no
    table is empty anywhere), total[o] = rows of that shard, adj i64 [world * F] as the kernels use it:
    virtual row = adj[o * F + f] + (r + f) // world."""
    F = len(num_rows_list)
    base = [[0] * F for _ in range(world)]
    total = [0] * world
    adj = torch.zeros(world * F, dtype=torch.int64)
    for o in range(world):
        acc = 0
        for f, v in enumerate(num_rows_list):
            first, count = owned_rows(v, f, o, world)
            base[o][f] = acc
            adj[o * F + f] = acc - (first + f) // world
            acc += max(count, 1)
        total[o] = acc
    return base, total, adj


class _PeerLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, st, ids_list, dense, *weights):
        ctx.st = st
        ctx.has_dense = dense is not None
        ctx.dense_width = 0 if dense is None else dense.shape[1]
        outs = st._forward(ids_list, dense)
        ctx.training = st.training
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        st = ctx.st
        st._backward(grad_outs)
        gdense = None
        if ctx.has_dense and ctx.needs_input_grad[2] and grad_outs[0] is not None:
            c0 = st.num_features * st.dims[0]
            gdense = grad_outs[0][:, c0:c0 + ctx.dense_width]
        return (None, None, gdense) + (None,) * len(st.dims)


class PeerShardedTables(nn.Module):
    """The tables of one lookup group, row-sharded over the ranks of ``transport``.  ``dims`` lists the widths that
    share the ids (DeepFM: ``[emb_dim, 1]``); width ``w`` of table ``f`` is ``full_tables[w][f]`` (``[V_f, dims[w]]``).

    One training lookup is in flight per module at a time: the routing lists, gradient buffers and the requester-side
    plan are per-module peer buffers that every rank reads between the two barriers of ``backward``, so a second
    training ``forward`` must come after the ``backward`` of the first (forward / backward / forward / backward, as
    the reference loop ``torchctr/trainer.py:291-303`` and gradient accumulation over micro-batches do).  All ranks
    must call ``forward`` / ``backward`` the same number of times with the same batch size."""

    def __init__(self, names, full_tables, transport, device=None, dedup: bool = False, init_seed: int | None = None,
                 init_std: float = 1.0):
        """``dedup``: every distinct row of this rank's batch crosses NVLink once per direction (the requester sorts
        its slots, fetches the distinct rows into a staging matrix and pools from there; backward it reduces its own
        duplicates before the owners pull) instead of once per id slot.  Pays a local sort on the forward path, saves
        most of the NVLink traffic under Zipf ids; worth it from about four ranks on."""
        super().__init__()
        self.dedup = bool(dedup)
        self.transport = transport
        self.world, self.rank = transport.world, transport.rank
        if self.world > _lib.MAX_WORLD:
            raise ValueError(f"at most {_lib.MAX_WORLD} ranks")
        dev = torch.device(device if device is not None else transport.device)
        self.device = dev
        self.names = list(names)
        self.num_features = len(self.names)
        first = full_tables[0]
        # Vocabulary-indexed tables GROW while training (torchctr/transformer.py:451-498): their shard geometry is laid out for
        # a fixed row capacity (``vocab_max_rows``) and ``live_rows`` says how many rows exist so far.  The key -> row map is
        # REPLICATED: every step the ranks all-gather their batch keys and each fits the same global key sequence (rank 0's
        # keys first), so all maps agree and the rows are exactly those a single process assigns to the global batch.
        self.index_kinds = [t.index_kind for t in first]
        self.vocabs = [t.vocab if t.index_kind == "vocab" else None for t in first]
        self.live_rows = [int(t.num_embeddings) for t in first]
        self.num_rows = [int(getattr(t, "vocab_max_rows", 0) or (1 << 20)) if t.index_kind == "vocab" else int(t.num_embeddings)
                         for t in first]
        self.hash_seeds = [t.hash_seed for t in first]
        self._grow_seed = torch.initial_seed() if init_seed is None else init_seed
        for t in first:
            if t.pooling != "sum" or t.use_id_weight:
                raise NotImplementedError("sharded lookup: sum pooling without per-id weights")
        self.dims = [int(tabs[0].embedding_dim) for tabs in full_tables]
        self.base, self.total, adj = shard_geometry(self.num_rows, self.world)
        self.register_buffer("adj", adj.to(dev), persistent=False)
        self._shard_struct = ops.make_shard(self.world, self.rank, self.adj)
        # this rank's fused shard of every width, in peer-visible memory
        self._shard_bufs, self._table_ptrs = [], []
        self.shards = nn.ParameterList()
        rows = self.total[self.rank]
        from ..nn.embedding import EmbeddingTable
        seed0 = torch.initial_seed() if init_seed is None else init_seed
        for wi, (tabs, D) in enumerate(zip(full_tables, self.dims)):
            buf = PeerBuffer(rows * D * 4, dev)
            w = buf.tensor(torch.float32, (rows, D))
            for f, t in enumerate(tabs):
                fr, n = owned_rows(self.num_rows[f], f, self.rank, self.world)
                if n:
                    b = self.base[self.rank][f]
                    if self.index_kinds[f] == "vocab":       # the rows that exist so far; the rest of the capacity is filled as it grows
                        live = self.live_rows[f]
                        cnt = 0 if fr >= live else (live - fr + self.world - 1) // self.world
                        if cnt and not t.weight.is_meta:
                            w[b:b + cnt] = t.weight.detach()[fr:live:self.world].to(dev)
                    elif t.weight.is_meta:
                        # shard-native: this rank draws ITS rows of the table with the counter-based generator; no rank ever
                        # holds the full table (BASELINE config 4: 2^30 rows x 64 floats = 256 GiB)
                        ops.normal_fill_rows_strided(w, b, n, 0.0, init_std, EmbeddingTable.counter_seed(seed0, f, wi), fr, self.world)
                    else:
                        w[b:b + n] = t.weight.detach()[fr::self.world].to(dev)
            self._shard_bufs.append(buf)
            self.shards.append(nn.Parameter(w, requires_grad=True))
            self._table_ptrs.append(ops.ptr_array(transport.share(buf)))
        self.vocab_modules = nn.ModuleList([v for v in self.vocabs if v is not None]).to(dev)   # checkpointed with the model
        self.opt_state0 = [None] * len(self.dims)
        self.opt_state1 = [None] * len(self.dims)
        self.bindings = [None] * len(self.dims)
        self._side = None           # side stream of the owner bucketing
        self._route_pending = False
        self._owner_planned = False
        self.early_owner_plan = True    # direct mode: meet the other ranks and sort on the side stream during the tower
        self._B = None              # batch geometry the peer buffers were sized for
        self._route_ws = None
        self._plan_ws = None

    # ---- optimizer ---------------------------------------------------------------------------------------------
    def bind_optimizer(self, optimizer, kind=None):
        self.bindings = []
        for w in range(len(self.dims)):
            holder = _ShardHolder(self.shards[w])
            self.bindings.append(SparseOptimizerBinding(optimizer, [holder], kind))

    def _ensure_state(self, w, kind, init):
        p = self.shards[w]
        if kind in ("adagrad", "adam") and self.opt_state0[w] is None:
            self.opt_state0[w] = torch.full_like(p.data, init if kind == "adagrad" else 0.0)
        if kind == "rowwise_adagrad" and self.opt_state0[w] is None:
            self.opt_state0[w] = torch.full((p.shape[0],), init, device=p.device)
        if kind == "adam" and self.opt_state1[w] is None:
            self.opt_state1[w] = torch.zeros_like(p.data)

    # ---- per-batch peer buffers ----------------------------------------------------------------------------------
    def _ensure_buffers(self, ids_list):
        """Peer buffers (routing lists, gradient matrices) are sized ONCE, collectively, for the largest batch any rank has
        seen in this call -- every rank enters the exchange together on the first training forward -- and later batches
        that fit (an uneven last batch, a smaller micro-batch) reuse them without any collective.  A batch that does not
        fit raises on that rank instead of silently entering a collective the other ranks are not in."""
        B = ids_list[0].shape[0]
        Ls = tuple(i.shape[1] for i in ids_list) + (self._dense_width,)
        if self._B is not None:
            capB, capLs = self._B
            if Ls == capLs and B <= capB:
                return
            raise RuntimeError(f"sharded lookup: batch geometry {(B, Ls)} does not fit the peer buffers sized for {self._B}; "
                               "all ranks must call reserve(max_batch) together before a larger batch is used")
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("peer buffers must be sized (one eager step) before the step is captured into a CUDA graph")
        geoms = self.transport.all_gather_object((B, Ls))
        if any(g[1] != Ls for g in geoms):
            raise RuntimeError(f"sharded lookup: ranks disagree on the feature layout: {geoms}")
        B = max(g[0] for g in geoms)
        S = B * sum(Ls[:-1])
        dev = self.device
        tr = self.transport
        self._route_buf = PeerBuffer(256 + 8 * max(S, 1), dev)          # [counts 64 words][keys S][slots S]
        ptrs = tr.share(self._route_buf)
        self._peer_counts = ops.ptr_array(ptrs)
        self._peer_keys = ops.ptr_array([p + 256 for p in ptrs])
        self._peer_slots = ops.ptr_array([p + 256 + 4 * max(S, 1) for p in ptrs])
        self._grad_bufs, self._peer_grads, self._strides = [], [], []
        for w, D in enumerate(self.dims):
            stride = (self.num_features * D + (self._dense_width if w == 0 else 0) + 3) // 4 * 4
            buf = PeerBuffer(B * stride * 4, dev)
            self._grad_bufs.append(buf)
            self._peer_grads.append(ops.ptr_array(tr.share(buf)))
            self._strides.append(stride)
        if self.dedup:
            cap = max(S, 1)
            self._uniq_buf = PeerBuffer(256 + 8 * cap, dev)            # [num_unique i64, pad][uniq_feature i32 S][uniq_row i32 S]
            ptrs = tr.share(self._uniq_buf)
            self._peer_nu = ops.ptr_array(ptrs)
            self._peer_uf = ops.ptr_array([p + 256 for p in ptrs])
            self._peer_ur = ops.ptr_array([p + 256 + 4 * cap for p in ptrs])
            self._rowgrad_bufs, self._peer_rowgrads = [], []
            for D in self.dims:
                buf = PeerBuffer(cap * D * 4, dev)
                self._rowgrad_bufs.append(buf)
                self._peer_rowgrads.append(ops.ptr_array(tr.share(buf)))
            self._one_id = torch.zeros(1, 1, dtype=torch.int64, device=dev)
        # owner-side groups are sized for the CAPACITY batch (a peer may send more slots than this rank's own batch holds);
        # their ids are never read -- the owner pulls the peers' routing lists
        self._owner_ids = [torch.zeros(B, L, dtype=torch.int64, device=dev) for L in Ls[:-1]]
        self._S = S
        self._B = (B, Ls)

    _dense_width = 0

    def _request_specs(self, ids_list, w):
        D = self.dims[w]
        return [ops.FeatureSpec(ids=ids, table=None, num_rows=v, D=D, out_col=f * D, index_kind=k, hash_seed=s,
                                vocab=None if vo is None else vo.handle())
                for f, (ids, v, k, s, vo) in enumerate(zip(ids_list, self.num_rows, self.index_kinds, self.hash_seeds, self.vocabs))]

    def grow_vocabularies(self, feats, std: float = 0.01) -> None:
        """Collective, once per training step before ``forward``: admit the new keys of the GLOBAL batch into the
        (replicated) vocabularies and create the rows they got -- N(0, std) from the counter-based generator keyed by the
        global row, as ``DynamicEmbedding`` draws them (``torchctr/nn/embedding.py:74-78``) -- on the ranks that own them."""
        from ..nn.embedding import EmbeddingTable
        for f, (name, vo) in enumerate(zip(self.names, self.vocabs)):
            if vo is None:
                continue
            keys = feats[name].to(self.device, dtype=torch.int64, non_blocking=True).contiguous()
            vo.fit(self.transport.all_gather_tensor(keys.reshape(keys.shape[0], -1)))
            n = vo.num_embeddings()
            if int(vo._last_status.item()) & _lib.STATUS_MAP_FULL:
                raise RuntimeError("vocabulary map overflow")
            if n > self.num_rows[f]:
                raise RuntimeError(f"table {name!r} outgrew its sharded capacity of {self.num_rows[f]} rows (vocab_max_rows)")
            old = self.live_rows[f]
            if n > old:
                for wi in range(len(self.dims)):
                    # rows r in [old, n) with (r + f) % world == rank, i.e. r = first_owned + k * world
                    fr, _ = owned_rows(self.num_rows[f], f, self.rank, self.world)
                    k0 = 0 if old <= fr else (old - fr + self.world - 1) // self.world
                    k1 = 0 if n <= fr else (n - fr + self.world - 1) // self.world
                    if k1 > k0:
                        b = self.base[self.rank][f]
                        ops.normal_fill_rows_strided(self.shards[wi].data, b + k0, k1 - k0, 0.0, std,
                                                     EmbeddingTable.counter_seed(self._grow_seed, f, wi), fr + k0 * self.world, self.world)
                self.live_rows[f] = n
        self.transport.barrier()

    def _owner_specs(self, ids_list, w, with_state, dedup=False):
        D = self.dims[w]
        shard = self.shards[w].data
        specs = []
        for f, ids in enumerate(ids_list):
            _, n = owned_rows(self.num_rows[f], f, self.rank, self.world)
            n = max(n, 1)
            b = self.base[self.rank][f]
            s0 = self.opt_state0[w] if with_state else None
            s1 = self.opt_state1[w] if with_state else None
            specs.append(ops.FeatureSpec(ids=ids, table=shard[b:b + n], num_rows=n, D=D, out_col=0 if dedup else f * D,
                                         state0=None if s0 is None else s0[b:b + n],
                                         state1=None if s1 is None else s1[b:b + n]))
        return specs

    # ---- forward / backward ----------------------------------------------------------------------------------------
    def forward(self, feats, dense=None):
        dev = self.device
        ids = []
        for n in self.names:
            t = feats[n]
            if t.dim() == 1:
                t = t.unsqueeze(1)
            ids.append(t.to(dev, dtype=torch.int64, non_blocking=True).contiguous())
        if dense is not None:
            dense = dense.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        return _PeerLookupFn.apply(self, ids, dense, *list(self.shards))

    def _forward(self, ids_list, dense):
        B = ids_list[0].shape[0]
        self._dense_width = 0 if dense is None else dense.shape[1]
        if self.training:
            self._ensure_buffers(ids_list)
        self._ids = ids_list
        if self.dedup:
            return self._forward_dedup(ids_list, dense)
        dev = self.device
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        outs = []
        for w, D in enumerate(self.dims):
            width = self.num_features * D + (self._dense_width if w == 0 else 0)
            stride = (width + 3) // 4 * 4
            out = torch.empty(B, stride, dtype=torch.float32, device=dev)
            call = ops.make_group(self._request_specs(ids_list, w), B, out, stride, dense=dense if w == 0 else None,
                                  dense_col=self.num_features * D, zero_from=width if stride > width else -1, status=status)
            ops.emb_pool_fwd_sharded(call, self._shard_struct, self._table_ptrs[w])
            outs.append(out)
        if self.training:
            # bucketing the slots by owner is only needed by the backward: it runs on a side stream next to the
            # tower (inside a captured CUDA graph this becomes a parallel branch)
            call = ops.make_group(self._request_specs(ids_list, 0), B, None, self.num_features * self.dims[0], status=status)
            need = ops.route_p2p_workspace_bytes(call)
            if self._route_ws is None or self._route_ws.numel() < need:
                self._route_ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            if not torch.cuda.is_current_stream_capturing():
                for t in ids_list:
                    t.record_stream(self._side)      # the ids are read on the side stream
            main = torch.cuda.current_stream(dev)
            self._side.wait_stream(main)
            rp = self._route_buf.ptr
            # the owner-side plan needs the routing lists of every rank, which depend on the ids only: build them,
            # meet the other ranks and sort -- all on the side stream, next to the tower, as the single-GPU path does
            # with its early sort.  backward() then only waits for the gradients.
            plan_call = ops.make_group(self._owner_specs(self._owner_ids, 0, False), self._B[0], None, self._strides[0])
            need_plan = ops.emb_bwd_p2p_workspace_bytes(plan_call, self.world)
            if self._plan_ws is None or self._plan_ws.numel() < need_plan:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("run one eager step before capturing: the owner-side workspace is not sized yet")
                self._plan_ws = torch.empty(need_plan + 256, dtype=torch.uint8, device=dev)
                self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                ops.route_p2p_build(call, self._shard_struct, rp, rp + 256, rp + 256 + 4 * max(self._S, 1), self._route_ws)
                if self.early_owner_plan:
                    self.transport.barrier()             # every rank's routing lists are in place
                    ops.emb_bwd_plan_p2p(plan_call, self._shard_struct, self._peer_counts, self._peer_keys, self._peer_slots,
                                         self._plan_ws)
                    self._owner_planned = True
            self._route_pending = True
        self.status = status
        return outs

    # ---- de-duplicated exchange ------------------------------------------------------------------------------------
    def _forward_dedup(self, ids_list, dense):
        B = ids_list[0].shape[0]
        dev = self.device
        S = sum(i.numel() for i in ids_list)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        # the plan of the single-GPU backward (sort + runs), on the requester's own slots, before the lookup
        plan_call = ops.make_group(self._request_specs(ids_list, 0), B, None, self.num_features * self.dims[0], status=status)
        ws = torch.empty(ops.emb_bwd_workspace_bytes(plan_call) + 256, dtype=torch.uint8, device=dev)
        ops.emb_bwd_plan(plan_call, ws, runs=True)
        uidx = torch.empty(S, dtype=torch.int64, device=dev)
        views, off = [], 0
        for ids in ids_list:
            views.append(uidx[off:off + ids.numel()].view(ids.shape))
            off += ids.numel()
        outs = []
        for w, D in enumerate(self.dims):
            staging = torch.empty(max(S, 1), D, dtype=torch.float32, device=dev)
            call = ops.make_group(self._request_specs(ids_list, w), B, None, self.num_features * D, status=status)
            ops.unique_fetch(call, self._shard_struct, self._table_ptrs[w], ws, staging, uidx if w == 0 else None)
            width = self.num_features * D + (self._dense_width if w == 0 else 0)
            stride = (width + 3) // 4 * 4
            out = torch.empty(B, stride, dtype=torch.float32, device=dev)
            specs = [ops.FeatureSpec(ids=v, table=staging, num_rows=max(S, 1), D=D, out_col=f * D) for f, v in enumerate(views)]
            pool = ops.make_group(specs, B, out, stride, dense=dense if w == 0 else None, dense_col=self.num_features * D,
                                  zero_from=width if stride > width else -1, status=status)
            ops.emb_pool_fwd(pool)
            outs.append(out)
        self._plan_local = ws
        self.status = status
        return outs

    def _backward_dedup(self, grad_outs):
        ids_list = self._ids
        B = ids_list[0].shape[0]
        dev = self.device
        cap = max(self._S, 1)
        live = [w for w, g in enumerate(grad_outs) if g is not None]
        ub = self._uniq_buf
        nu = ub.tensor(torch.int64, (1,))
        uf = ub.tensor(torch.int32, (cap,), 256)
        ur = ub.tensor(torch.int32, (cap,), 256 + 4 * cap)
        none_opt = ops.make_opt("none")
        for w in live:                                   # this rank's duplicates are summed here, once
            g = grad_outs[w]
            if g.stride(1) != 1 or g.stride(0) != g.shape[1]:
                g = g.contiguous()
            D = self.dims[w]
            call = ops.make_group(self._request_specs(ids_list, w), B, g, g.shape[1])
            rg = self._rowgrad_bufs[w].tensor(torch.float32, (cap, D))
            ops.emb_bwd_apply(call, self._plan_local, none_opt, uf, ur, rg, nu)
        self.transport.barrier()                         # every rank's distinct-row lists and gradients are in place
        world_cap = self.world * cap
        for w in live:
            bind = self.bindings[w]
            self._ensure_state(w, bind.kind, bind.initial_accumulator_value())
        owner_ids = [self._one_id] * self.num_features
        plan_call = ops.make_group(self._owner_specs(owner_ids, 0, False, dedup=True), 1, None, self.dims[0])
        need = ops.emb_bwd_p2p_unique_workspace_bytes(plan_call, world_cap)
        if self._plan_ws is None or self._plan_ws.numel() < need:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("run one eager step before capturing: the owner-side workspace is not sized yet")
            self._plan_ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        ops.emb_bwd_plan_p2p_unique(plan_call, self._shard_struct, self._peer_nu, self._peer_uf, self._peer_ur, world_cap, self._plan_ws)
        for w in live:
            call = ops.make_group(self._owner_specs(owner_ids, w, True, dedup=True), 1, None, self.dims[w])
            ops.emb_bwd_apply_p2p_unique(call, self._shard_struct, self._plan_ws, self.bindings[w].next_opt(),
                                         self._peer_rowgrads[w], world_cap)
        self.transport.barrier()

    def _backward(self, grad_outs):
        if any(b is None for b in self.bindings):
            raise RuntimeError("sharded tables need bind_optimizer(): there is no dense or sparse .grad to hand back")
        if self.dedup:
            return self._backward_dedup(grad_outs)
        ids_list = self._ids
        B = ids_list[0].shape[0]
        live = [w for w, g in enumerate(grad_outs) if g is not None]
        if self._route_pending:                          # the routing lists must be complete before the barrier
            torch.cuda.current_stream(self.device).wait_stream(self._side)
            self._route_pending = False
        for w in live:                                   # gradients into the peer-visible buffers
            g = grad_outs[w]
            if g.data_ptr() != self._grad_bufs[w].ptr:   # (the tower's first block may have written it in place)
                self._grad_bufs[w].tensor(torch.float32, (B, self._strides[w])).copy_(g)
        self.transport.barrier()                         # every rank's routing lists and gradients are in place
        if not self._owner_planned:                      # (no early plan was started in forward)
            plan_call = ops.make_group(self._owner_specs(self._owner_ids, 0, False), self._B[0], None, self._strides[0])
            need = ops.emb_bwd_p2p_workspace_bytes(plan_call, self.world)
            if self._plan_ws is None or self._plan_ws.numel() < need:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("run one eager step before capturing: the owner-side workspace is not sized yet")
                self._plan_ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
            ops.emb_bwd_plan_p2p(plan_call, self._shard_struct, self._peer_counts, self._peer_keys, self._peer_slots, self._plan_ws)
        self._owner_planned = False
        for w in live:
            bind = self.bindings[w]
            self._ensure_state(w, bind.kind, bind.initial_accumulator_value())
            call = ops.make_group(self._owner_specs(self._owner_ids, w, True), self._B[0], None, self._strides[w])
            ops.emb_bwd_apply_p2p(call, self._shard_struct, self._plan_ws, bind.next_opt(), self._peer_grads[w])
        self.transport.barrier()                         # nobody overwrites lists / gradients / reads rows too early

    def grad_buffer_provider(self, w):
        """A callable returning this rank's peer-visible gradient matrix of width ``w`` as a tensor [B, stride]: whoever
        produces dL/d(pooled output) may write it there directly and hand it back through autograd."""
        def provider():
            if self._B is None or self.dedup:
                return None
            return self._grad_bufs[w].tensor(torch.float32, (self._B[0], self._strides[w]))
        return provider

    # ---- checkpoints in the reference's (unsharded) format ------------------------------------------------------------
    def export_full_tables(self, w: int = 0):
        """Collective.  Every rank gets the full ``[V_f, dims[w]]`` table of every feature (CPU tensors), i.e. what
        ``model.embeddings[name].weight`` holds in the reference and in an unsharded model: a checkpoint written from
        them loads into ``torchctr.models.DNN`` (``torchctr/trainer.py:353-496``, ``nn/embedding.py:89-95``)."""
        torch.cuda.synchronize(self.device)
        mine = [self.local_rows_of(w, f)[1].detach().cpu() for f in range(self.num_features)]
        parts = self.transport.all_gather_object(mine)
        full = []
        for f, v in enumerate(self.num_rows):
            t = torch.empty(v, self.dims[w], dtype=torch.float32)
            for r in range(self.world):
                fr, n = owned_rows(v, f, r, self.world)
                if n:
                    t[fr::self.world] = parts[r][f]
            full.append(t[:self.live_rows[f]] if self.vocabs[f] is not None else t)     # a growing table: the rows that exist
        return full

    def export_full_optimizer_state(self, w: int = 0):
        """Collective.  The fused-update state of width ``w`` gathered like ``export_full_tables``: a list (one entry per
        feature) of ``(state0, state1)`` full tensors or None when the width has no state yet."""
        torch.cuda.synchronize(self.device)
        out = []
        for which, states in enumerate((self.opt_state0, self.opt_state1)):
            st = states[w]
            mine = None
            if st is not None:
                mine = []
                for f in range(self.num_features):
                    _, n = owned_rows(self.num_rows[f], f, self.rank, self.world)
                    b = self.base[self.rank][f]
                    mine.append(st[b:b + n].detach().cpu())
            parts = self.transport.all_gather_object(mine)
            if parts[0] is None:
                out.append(None)
                continue
            full = []
            for f, v in enumerate(self.num_rows):
                t = torch.zeros((v,) + tuple(parts[0][f].shape[1:]), dtype=torch.float32)
                for r in range(self.world):
                    fr, n = owned_rows(v, f, r, self.world)
                    if n:
                        t[fr::self.world] = parts[r][f]
                full.append(t)
            out.append(full)
        return out

    def load_full_optimizer_state(self, state, w: int = 0) -> None:
        """Inverse of ``export_full_optimizer_state`` (every rank passes the same full tensors)."""
        for which, name in enumerate(("opt_state0", "opt_state1")):
            full = state[which]
            if full is None:
                continue
            rows = self.total[self.rank]
            buf = torch.zeros((rows,) + tuple(full[0].shape[1:]), dtype=torch.float32, device=self.device)
            for f, t in enumerate(full):
                fr, n = owned_rows(self.num_rows[f], f, self.rank, self.world)
                if n:
                    b = self.base[self.rank][f]
                    buf[b:b + n] = t[fr::self.world].to(self.device)
            getattr(self, name)[w] = buf

    def load_full_tables(self, full, w: int = 0) -> None:
        """Scatter full tables (one ``[V_f, dims[w]]`` tensor per feature, e.g. from a reference checkpoint) into
        this rank's shard.  Every rank calls it with the same tensors; optimizer state of the width is reset."""
        with torch.no_grad():
            for f, t in enumerate(full):
                growing = self.vocabs[f] is not None
                if t.shape[1] != self.dims[w] or (t.shape[0] != self.num_rows[f] and not (growing and t.shape[0] <= self.num_rows[f])):
                    raise ValueError(f"table {f}: expected {(self.num_rows[f], self.dims[w])}, got {tuple(t.shape)}")
                fr, rows = self.local_rows_of(w, f)
                src = t[fr::self.world].to(self.device)
                if src.shape[0]:
                    rows[:src.shape[0]].copy_(src)
                if growing:
                    self.live_rows[f] = int(t.shape[0])
        self.opt_state0[w] = None
        self.opt_state1[w] = None
        self.transport.barrier()

    # ---- inspection (tests, checkpoints): this rank's rows of table f, width w ------------------------------------------
    def local_rows_of(self, w, f):
        fr, n = owned_rows(self.num_rows[f], f, self.rank, self.world)
        b = self.base[self.rank][f]
        return fr, self.shards[w].data[b:b + n]


class _ShardHolder:
    """What SparseOptimizerBinding needs from a table module: its ``weight``."""

    def __init__(self, weight):
        self.weight = weight
