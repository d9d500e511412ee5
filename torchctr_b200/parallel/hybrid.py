"""Hybrid table placement for Criteo-shaped models: small tables replicated, large tables row-sharded.

The reference is replicas-only (``Accelerator().prepare`` -> DDP, ``torchctr/trainer.py:128-130``): every rank holds every
table and dense ``[V, D]`` gradients of ALL tables are all-reduced.  ``PeerShardedTables`` is the other extreme: every table
row-sharded, every id slot crosses NVLink twice per step.  On a Criteo-shaped batch most slots belong to SMALL tables
(18 of 26 tables have <= 1e5 rows and take 69 % of the slots), whose full gradient is a few MB -- cheaper to all-reduce
than to route -- while the large tables' gradients are far too sparse for that.  So:

  replicated  tables with <= ``replicate_max_rows`` rows live on every rank in one contiguous block.  The lookup reads them
              locally; backward, the fused sweep runs in ``CTR_OPT_GRAD_OUT`` mode (summed row gradients into a zeroed dense
              buffer, FM / first-order terms included), ONE NCCL all-reduce sums the buffer over the ranks and
              ``ctr_rows_dense_apply`` gives every replica the same sgd / adagrad update (zero gradients skipped);
  sharded     the other tables are cut as in ``PeerShardedTables`` (row r of the j-th sharded table on rank (r + j) mod P):
              rows are read from the owner's shard through its NVLink peer mapping INSIDE the same lookup kernel, and the
              owners pull the (row, slot) lists, sort them on a side stream during the tower and run the fused sweep reading
              every slot's gradient (and, for DeepFM, its dL/d extra and FM sum) from the rank that produced it.

  hot rows    a large table with direct ids is cut in two: its first ``hot_rows`` rows (the hot ones of a frequency-ordered
              vocabulary: 16 K rows of a 1e7-row Zipf(1.05) table take 69 % of its slots) are replicated like a small table, only
              the tail is sharded.  Both parts are features over the same ids and the same output columns
              (``CTR_INDEX_WINDOW``); a slot belongs to exactly one of them.

One lookup launch and, per rank, two sweeps per step -- DeepFM's first-order tables and FM term ride inside them exactly as
on one GPU (``ctr_group_t.extra``).  The all-reduce doubles as the closing barrier of the owner-side update.
Single-id features of one width (D = 16, 32 or 64), direct or hashed ids, sum pooling, sgd / adagrad.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, ops
from ..nn.embedding import SparseOptimizerBinding
from .peer import PeerBuffer, _ShardHolder, owned_rows, shard_geometry


def hybrid_eligible(tables, twins=None) -> bool:
    """Can ``HybridShardedTables`` hold these tables?"""
    if not tables:
        return False
    D = tables[0].embedding_dim
    if D not in (16, 32, 64):
        return False
    for t in list(tables) + list(twins or []):
        if t.pooling != "sum" or t.use_id_weight or t.index_kind not in ("direct", "hash"):
            return False
    if any(t.embedding_dim != D for t in tables) or any(t.embedding_dim != 1 for t in (twins or [])):
        return False
    return True


class _HybridLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, st, ids_list, dense, *weights):
        ctx.st = st
        ctx.has_dense = dense is not None
        ctx.dense_width = 0 if dense is None else dense.shape[1]
        x, extra = st._forward(ids_list, dense)
        if extra is None:
            return (x,)
        return x, extra

    @staticmethod
    def backward(ctx, gx, gextra=None):
        st = ctx.st
        st._backward(gx, gextra)
        gdense = None
        if ctx.has_dense and ctx.needs_input_grad[2] and gx is not None:
            c0 = st.num_features * st.D
            gdense = gx[:, c0:c0 + ctx.dense_width]
        return (None, None, gdense) + (None,) * len(st.shards)


class HybridShardedTables(nn.Module):
    """``tables``: the full tables of the model in feature order (``twins``: DeepFM's one-column first-order tables, same
    order, or None); ``fm``: also produce the FM second-order term.  ``forward(feats, dense)`` returns ``(x,)`` or, with
    twins / fm, ``(x, extra)`` with ``extra`` f32 [B] = sum of the first-order weights (+ FM term).

    One training lookup is in flight per module at a time (forward / backward / forward / backward); every rank calls
    ``forward`` / ``backward`` the same number of times."""

    def __init__(self, names, tables, twins, transport, device=None, fm: bool = False, replicate_max_rows: int = 1 << 17,
                 init_seed: int | None = None, init_std: float = 1.0, hot_rows: int = 0):
        super().__init__()
        if not hybrid_eligible(tables, twins):
            raise NotImplementedError("hybrid placement: single-id sum-pooled tables of one width (16, 32 or 64), direct / hashed ids")
        from ..nn.embedding import EmbeddingTable
        self.transport = transport
        self.world, self.rank = transport.world, transport.rank
        if self.world > _lib.MAX_WORLD:
            raise ValueError(f"at most {_lib.MAX_WORLD} ranks")
        dev = torch.device(device if device is not None else transport.device)
        self.device = dev
        self.names = list(names)
        F = self.num_features = len(self.names)
        self.D = D = int(tables[0].embedding_dim)
        self.has_twins = twins is not None
        self.fm = bool(fm)
        self.fused_extra = self.has_twins or self.fm
        self.dims = [D, 1] if self.has_twins else [D]
        self.num_rows = [int(t.num_embeddings) for t in tables]
        self.vocabs = [None] * F
        # parts: (feature, first row, rows, index kind, hash seed | window start).  A table is one part, or -- a large table
        # with direct ids and hot_rows > 0 -- a replicated head [0, hot_rows) and a sharded tail [hot_rows, V)
        self.parts, self.sh, self.rp = [], [], []
        for f, t in enumerate(tables):
            V, kind = self.num_rows[f], t.index_kind
            if V <= replicate_max_rows:
                self.rp.append(len(self.parts)); self.parts.append((f, 0, V, kind, t.hash_seed))
            elif kind == "direct" and 0 < hot_rows < V:
                self.rp.append(len(self.parts)); self.parts.append((f, 0, hot_rows, "window", 0))
                self.sh.append(len(self.parts)); self.parts.append((f, hot_rows, V - hot_rows, "window", hot_rows))
            else:
                self.sh.append(len(self.parts)); self.parts.append((f, 0, V, kind, t.hash_seed))
        if len(self.parts) > _lib.MAX_FEATURES:
            raise ValueError(f"{len(self.parts)} table parts exceed the {_lib.MAX_FEATURES} features of a launch group; use hot_rows=0")
        self.order = self.sh + self.rp            # part order of the lookup group: sharded parts first (= the shard geometry's index)
        P, Fs = len(self.parts), len(self.sh)
        seed0 = torch.initial_seed() if init_seed is None else init_seed
        widths = [tables] + ([twins] if self.has_twins else [])

        # ---- sharded parts: this rank's rows (row r of the j-th sharded part on rank (r + j) mod P), in peer-visible memory ----
        if Fs:
            self.base, self.total, adj_s = shard_geometry([self.parts[p][2] for p in self.sh], self.world)
        else:
            self.base, self.total, adj_s = [], [1] * self.world, torch.zeros(1, dtype=torch.int64)
        adj_l = torch.zeros(self.world * P, dtype=torch.int64)           # the same for the lookup group (P features)
        for o in range(self.world):
            for j in range(Fs):
                adj_l[o * P + j] = adj_s[o * Fs + j]
        self.register_buffer("adj_s", adj_s.to(dev), persistent=False)
        self.register_buffer("adj_l", adj_l.to(dev), persistent=False)
        self._shard_s = ops.make_shard(self.world, self.rank, self.adj_s)
        self._shard_l = ops.make_shard(self.world, self.rank, self.adj_l)
        rows = max(self.total[self.rank], 1) if Fs else 1
        self._shard_bufs, self._table_ptrs = [], []
        params = []
        for wi, (tabs, Dw) in enumerate(zip(widths, self.dims)):
            buf = PeerBuffer(rows * Dw * 4, dev)
            w = buf.tensor(torch.float32, (rows, Dw))
            for j, p in enumerate(self.sh):
                f, first, cnt = self.parts[p][:3]
                t = tabs[f]
                fr, n = owned_rows(cnt, j, self.rank, self.world)
                if not n:
                    continue
                b = self.base[self.rank][j]
                if t.weight.is_meta:       # shard-native: this rank draws ITS rows; no rank ever holds the full table
                    ops.normal_fill_rows_strided(w, b, n, 0.0, init_std, EmbeddingTable.counter_seed(seed0, f, wi), first + fr, self.world)
                else:
                    w[b:b + n] = t.weight.detach()[first + fr:first + cnt:self.world].to(dev)
            self._shard_bufs.append(buf)
            params.append(nn.Parameter(w, requires_grad=True))
            self._table_ptrs.append(ops.ptr_array(transport.share(buf)))

        # ---- replicated parts: whole, in one block per width ----
        self.rep_off, acc = {}, 0
        for p in self.rp:
            self.rep_off[p] = acc
            acc += self.parts[p][2]
        self.R = R = (acc + 3) // 4 * 4 if acc else 0
        for wi, (tabs, Dw) in enumerate(zip(widths, self.dims)):
            blk = torch.zeros(max(R, 4), Dw, dtype=torch.float32, device=dev)
            for p in self.rp:
                f, first, cnt = self.parts[p][:3]
                t, o = tabs[f], self.rep_off[p]
                if t.weight.is_meta:
                    ops.normal_fill_rows_strided(blk, o, cnt, 0.0, init_std, EmbeddingTable.counter_seed(seed0, f, wi), first, 1)
                else:
                    blk[o:o + cnt] = t.weight.detach()[first:first + cnt].to(dev)
            params.append(nn.Parameter(blk, requires_grad=True))
        # shards[0 .. nw): this rank's shard of every width; shards[nw .. 2 nw): the replicated block of every width
        self.shards = nn.ParameterList(params)
        # gradient buffer of the replicated block: [R, D] then [R] (twins), ONE flat tensor = one all-reduce
        self._rep_elems = max(R, 4) * (D + (1 if self.has_twins else 0))
        self._rep_grad = torch.zeros(self._rep_elems, dtype=torch.float32, device=dev)
        self.opt_state = [None] * (2 * len(self.dims))      # adagrad sums, same shapes as self.shards
        self.bindings = [None]
        self.binding = None
        self._side = None
        self._route_pending = False
        self._cap = None                        # (B, dense width) the peer buffers were sized for
        self._route_ws = self._plan_ws = self._rep_ws = None
        self._dense_width = 0
        self.status = None

    # ---- optimizer ---------------------------------------------------------------------------------------------
    def bind_optimizer(self, optimizer, kind=None):
        b = SparseOptimizerBinding(optimizer, [_ShardHolder(self.shards[0])], kind)
        if b.kind not in ("sgd", "adagrad"):
            raise NotImplementedError(f"hybrid placement updates its replicated tables densely: sgd / adagrad only, not {b.kind!r} "
                                      "(use PeerShardedTables: shard_model(..., hybrid=False))")
        self.binding = b
        self.bindings = [b]

    def _ensure_state(self):
        if self.binding.kind != "adagrad" or self.opt_state[0] is not None:
            return
        init = self.binding.initial_accumulator_value()
        self.opt_state = [torch.full_like(p.data, init) for p in self.shards]

    # ---- per-batch peer buffers ----------------------------------------------------------------------------------
    def _ensure_buffers(self, B):
        key = self._dense_width
        if self._cap is not None:
            capB, capkey = self._cap
            if key == capkey and B <= capB:
                return
            raise RuntimeError(f"hybrid lookup: batch geometry {(B, key)} does not fit the peer buffers sized for {self._cap}")
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("peer buffers must be sized (one eager step) before the step is captured into a CUDA graph")
        geoms = self.transport.all_gather_object((B, key))
        if any(g[1] != key for g in geoms):
            raise RuntimeError(f"hybrid lookup: ranks disagree on the feature layout: {geoms}")
        B = max(g[0] for g in geoms)
        dev, tr, D, F = self.device, self.transport, self.D, self.num_features
        Fs = len(self.sh)
        S = max(B * Fs, 1)
        self._S = S
        self._route_buf = PeerBuffer(256 + 8 * S, dev)                  # [counts 64 words][keys S][slots S]
        ptrs = tr.share(self._route_buf)
        self._peer_counts = ops.ptr_array(ptrs)
        self._peer_keys = ops.ptr_array([p + 256 for p in ptrs])
        self._peer_slots = ops.ptr_array([p + 256 + 4 * S for p in ptrs])
        self._stride = (F * D + self._dense_width + 3) // 4 * 4
        self._grad_buf = PeerBuffer(B * self._stride * 4, dev)          # dL/dx of this rank's bags
        self._peer_grads = ops.ptr_array(tr.share(self._grad_buf))
        self._peer_extra = None
        if self.fused_extra:
            self._extra_buf = PeerBuffer(B * 4, dev)                    # dL/d extra of this rank's bags
            self._peer_extra = ops.ptr_array(tr.share(self._extra_buf))
            self._gextra = self._extra_buf.tensor(torch.float32, (B,))
        self._peer_pack = None
        if self.fm:
            self._fm_sum = torch.empty(B, D, dtype=torch.float32, device=dev)   # sum over the fields of the pooled vectors
            if Fs:
                # the sharded parts' gradients with the "c * fm_sum" part of the FM gradient folded in, packed [B, Fs * D]: what
                # the owners pull (one D-float piece per slot instead of the gradient slice plus the bag's fm_sum row)
                self._pack_buf = PeerBuffer(B * Fs * D * 4, dev)
                self._peer_pack = ops.ptr_array(tr.share(self._pack_buf))
        self._owner_ids = torch.zeros(B, 1, dtype=torch.int64, device=dev)   # never read: the owner pulls the peers' lists
        self._cap = (B, key)

    # ---- feature specs ---------------------------------------------------------------------------------------------
    def _part_kw(self, p, ids_list_pos, ids_list):
        f, first, cnt, kind, seed = self.parts[p]
        return dict(ids=ids_list[ids_list_pos], num_rows=cnt, D=self.D, out_col=f * self.D, index_kind=kind, hash_seed=seed)

    def _rep_slices(self, p):
        """(table rows, twin rows | None) of replicated part ``p`` inside the replicated blocks."""
        nw = len(self.dims)
        o, cnt = self.rep_off[p], self.parts[p][2]
        rep_w = self.shards[nw].data
        rep_t = self.shards[nw + 1].data if self.has_twins else None
        return rep_w[o:o + cnt], (None if rep_t is None else rep_t[o:o + cnt])

    def _lookup_specs(self, ids_list):
        """Sharded parts first (their position is the index the shard geometry uses), then the replicated ones."""
        specs = []
        for j, p in enumerate(self.order):
            kw = self._part_kw(p, j, ids_list)
            if j < len(self.sh):
                specs.append(ops.FeatureSpec(table=None, **kw))
            else:
                w, tw = self._rep_slices(p)
                specs.append(ops.FeatureSpec(table=w, twin_table=tw, **kw))
        return specs

    def _rep_bwd_specs(self, ids_list, plan: bool = False):
        """The replicated parts for the GRAD_OUT sweep: state0 / twin_state0 are the dense gradient buffers
        (``plan``: ids and row counts only, for the sort)."""
        D = self.D
        R = max(self.R, 4)
        g_main = self._rep_grad[:R * D].view(R, D)
        g_twin = self._rep_grad[R * D:R * D + R] if self.has_twins else None
        specs = []
        for j, p in enumerate(self.order):
            if j < len(self.sh):
                continue
            kw = self._part_kw(p, j, ids_list)
            if plan:
                specs.append(ops.FeatureSpec(table=None, **kw))
                continue
            o, cnt = self.rep_off[p], self.parts[p][2]
            w, tw = self._rep_slices(p)
            specs.append(ops.FeatureSpec(table=w, state0=g_main[o:o + cnt], twin_table=tw,
                                         twin_state0=None if g_twin is None else g_twin[o:o + cnt], **kw))
        return specs

    def _route_specs(self, ids_list):
        return [ops.FeatureSpec(table=None, **self._part_kw(p, j, ids_list)) for j, p in enumerate(self.sh)]

    def _owner_specs(self, with_state, packed: bool = False):
        """This rank's rows of the sharded parts.  ``packed``: the gradients come from the peers' packed matrices
        ([B, Fs * D], part j at column j * D) instead of their full dL/dx matrices."""
        D = self.D
        shard = self.shards[0].data
        twin = self.shards[1].data.view(-1) if self.has_twins else None
        s0 = self.opt_state[0] if with_state else None
        t0 = self.opt_state[1].view(-1) if (with_state and self.has_twins and self.opt_state[1] is not None) else None
        specs = []
        for j, p in enumerate(self.sh):
            f, first, cnt = self.parts[p][:3]
            _, n = owned_rows(cnt, j, self.rank, self.world)
            n = max(n, 1)
            b = self.base[self.rank][j]
            specs.append(ops.FeatureSpec(ids=self._owner_ids, table=shard[b:b + n], num_rows=n, D=D, out_col=j * D if packed else f * D,
                                         state0=None if s0 is None else s0[b:b + n],
                                         twin_table=None if twin is None or not with_state else twin[b:b + n],
                                         twin_state0=None if t0 is None else t0[b:b + n]))
        return specs

    # ---- forward / backward ----------------------------------------------------------------------------------------
    def _ids_in_order(self, feats):
        dev = self.device
        cache, ids = {}, []
        for p in self.order:
            f = self.parts[p][0]
            if f not in cache:
                t = feats[self.names[f]]
                if t.dim() == 1:
                    t = t.unsqueeze(1)
                if t.shape[1] != 1:
                    raise ValueError(f"hybrid placement holds single-id features; {self.names[f]!r} has {t.shape[1]} ids per row")
                cache[f] = t.to(dev, dtype=torch.int64, non_blocking=True).contiguous()
            ids.append(cache[f])
        return ids

    def forward(self, feats, dense=None):
        ids = self._ids_in_order(feats)
        if dense is not None:
            dense = dense.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        self._want_backward = self.training and torch.is_grad_enabled()      # (grad mode is off inside Function.forward)
        return _HybridLookupFn.apply(self, ids, dense, *list(self.shards))

    _want_backward = None

    def _forward(self, ids_list, dense):
        B = ids_list[0].shape[0]
        dev, D, F = self.device, self.D, self.num_features
        self._dense_width = 0 if dense is None else dense.shape[1]
        train = self._want_backward if self._want_backward is not None else (self.training and torch.is_grad_enabled())
        self._want_backward = None
        if train:
            self._ensure_buffers(B)
        self._ids = ids_list
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        width = F * D + self._dense_width
        stride = (width + 3) // 4 * 4
        x = torch.empty(B, stride, dtype=torch.float32, device=dev)
        extra = torch.empty(B, dtype=torch.float32, device=dev) if self.fused_extra else None
        fm_sum = None
        if self.fm:
            fm_sum = self._fm_sum[:B] if train else torch.empty(B, D, dtype=torch.float32, device=dev)
        call = ops.make_group(self._lookup_specs(ids_list), B, x, stride, dense=dense, dense_col=F * D,
                              zero_from=width if stride > width else -1, status=status, extra=extra, fm_sum=fm_sum, fm=self.fm)
        ops.emb_pool_fwd_sharded(call, self._shard_l, self._table_ptrs[0], self._table_ptrs[1] if self.has_twins else None)
        if train:
            # everything the backward needs that depends on the ids only runs on a side stream next to the tower: the sort of
            # the replicated parts' slots, the bucketing of the sharded parts' slots by owner, the meeting with the other
            # ranks and the owner-side gather + sort
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            capturing = torch.cuda.is_current_stream_capturing()
            if not capturing:
                for t in ids_list:
                    t.record_stream(self._side)
            main = torch.cuda.current_stream(dev)
            route_call = plan_call = rep_call = None
            if self.sh:
                route_call = ops.make_group(self._route_specs(ids_list), B, None, stride, status=status)
                need = ops.route_p2p_workspace_bytes(route_call)
                if self._route_ws is None or self._route_ws.numel() < need:
                    self._route_ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
                plan_call = ops.make_group(self._owner_specs(False), self._cap[0], None, self._stride)
                need = ops.emb_bwd_p2p_workspace_bytes(plan_call, self.world)
                if self._plan_ws is None or self._plan_ws.numel() < need:
                    if capturing:
                        raise RuntimeError("run one eager step before capturing: the owner-side workspace is not sized yet")
                    self._plan_ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            if self.rp:
                rep_call = ops.make_group(self._rep_bwd_specs(ids_list, plan=True), B, None, stride, status=status)
                need = ops.emb_bwd_workspace_bytes(rep_call)
                if self._rep_ws is None or self._rep_ws.numel() < need:
                    if capturing:
                        raise RuntimeError("run one eager step before capturing: the plan workspace is not sized yet")
                    self._rep_ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                if route_call is not None:
                    rp = self._route_buf.ptr
                    ops.route_p2p_build(route_call, self._shard_s, rp, rp + 256, rp + 256 + 4 * self._S, self._route_ws)
                    self.transport.barrier()                 # every rank's routing lists are in place
                    ops.emb_bwd_plan_p2p(plan_call, self._shard_s, self._peer_counts, self._peer_keys, self._peer_slots, self._plan_ws)
                if rep_call is not None:
                    ops.emb_bwd_plan(rep_call, self._rep_ws, runs=False)
            self._route_pending = True
        self.status = status
        return x, extra

    def _backward(self, gx, gextra):
        if self.binding is None:
            raise RuntimeError("hybrid tables need bind_optimizer(): there is no dense or sparse .grad to hand back")
        dev, D = self.device, self.D
        ids_list = self._ids
        B = ids_list[0].shape[0]
        if self._route_pending:
            torch.cuda.current_stream(dev).wait_stream(self._side)
            self._route_pending = False
        gbuf = self._grad_buf.tensor(torch.float32, (B, self._stride))
        if gx is None:
            gbuf.zero_()
        elif gx.data_ptr() != self._grad_buf.ptr:          # (the tower's first block may have written it in place)
            gbuf.copy_(gx)
        ge = None
        if self.fused_extra:
            ge = self._gextra[:B]
            if gextra is None:
                ge.zero_()
            else:
                ge.copy_(gextra.reshape(-1))
        self._ensure_state()
        opt = self.binding.next_opt()
        nw = len(self.dims)
        fm_sum = self._fm_sum[:B] if self.fm else None
        main = torch.cuda.current_stream(dev)
        packed = self._peer_pack is not None
        if packed:
            ops.fm_pack_grads(gbuf, ge, fm_sum, [self.parts[p][0] * D for p in self.sh], D,
                              self._pack_buf.tensor(torch.float32, (B, len(self.sh) * D)))
        if self.sh:
            self.transport.barrier()                       # every rank's gradients (and routing lists) are in place
        side = main
        if self.rp:
            # replicated parts, on a second stream NEXT TO the owner-side update of the sharded parts: this rank's summed row
            # gradients -> dense buffer, all-reduce (the tower's flat gradient buffer rides along), dense update of every replica
            if self.sh:
                if self._side2 is None:
                    self._side2 = torch.cuda.Stream(device=dev)
                side = self._side2
                side.wait_stream(main)
            with torch.cuda.stream(side):
                call = ops.make_group(self._rep_bwd_specs(ids_list), B, gbuf, self._stride, extra=ge, fm_sum=fm_sum, fm=self.fm)
                ops.emb_bwd_apply(call, self._rep_ws, ops.make_opt("grad_out"))
                from ..nn.tower import wait_deferred
                wait_deferred(dev)              # the tower's weight gradients (second stream) ride in this all-reduce
                self.transport.all_reduce(self._rep_grad)
                R = max(self.R, 4)
                st = self.opt_state
                ops.rows_dense_apply(self.shards[nw].data, self._rep_grad[:R * D], None if st[nw] is None else st[nw], opt)
                if self.has_twins:
                    ops.rows_dense_apply(self.shards[nw + 1].data, self._rep_grad[R * D:R * D + R],
                                         None if st[nw + 1] is None else st[nw + 1], opt)
        if self.sh:
            capB = self._cap[0]
            if packed:
                call = ops.make_group(self._owner_specs(True, packed=True), capB, None, len(self.sh) * D, extra=self._gextra,
                                      fm_sum=self._fm_sum, fm=True)           # (fm_sum: a marker here, the owner does not read it)
                ops.emb_bwd_apply_p2p(call, self._shard_s, self._plan_ws, opt, self._peer_pack, peer_extra=self._peer_extra)
            else:
                call = ops.make_group(self._owner_specs(True), capB, None, self._stride, extra=self._gextra if self.fused_extra else None)
                ops.emb_bwd_apply_p2p(call, self._shard_s, self._plan_ws, opt, self._peer_grads, peer_extra=self._peer_extra)
            if side is not main:
                main.wait_stream(side)
            self.transport.barrier()                       # closing: nobody overwrites what an owner still reads / reads rows too early
        # (no sharded part: the all-reduce above is the only collective, and it is on this stream)

    _side2 = None

    def adopt_dense_grads(self, numel: int):
        """Room for ``numel`` more floats behind the replicated tables' gradients: the model's flat dense-gradient buffer
        (``CTRModelBase.enable_flat_dense_grads``) then travels in the same all-reduce, which runs at the end of backward -- the
        lookup is the first node of the graph, so every dense gradient has been accumulated by then.  Returns the view, or
        None when there is no replicated table (no all-reduce here)."""
        if not self.rp:
            return None
        flat = torch.zeros(self._rep_elems + (numel + 3) // 4 * 4, dtype=torch.float32, device=self.device)
        flat[:self._rep_elems].copy_(self._rep_grad)
        self._rep_grad = flat
        return flat[self._rep_elems:self._rep_elems + numel]

    def grad_buffer_provider(self, w):
        """A callable returning this rank's peer-visible gradient matrix as a tensor [B, stride]: whoever produces
        dL/d(pooled output) may write it there directly and hand it back through autograd."""
        def provider():
            if self._cap is None or w != 0:
                return None
            return self._grad_buf.tensor(torch.float32, (self._cap[0], self._stride))
        return provider

    # ---- checkpoints in the reference's (unsharded) format ------------------------------------------------------------
    def _gather_width(self, shard_t, rep_t, wi):
        """Full per-feature tensors (CPU) of width ``wi`` from a shard-shaped and a replicated-block-shaped tensor."""
        torch.cuda.synchronize(self.device)
        Dw = self.dims[wi]
        mine = []
        for j, p in enumerate(self.sh):
            _, n = owned_rows(self.parts[p][2], j, self.rank, self.world)
            b = self.base[self.rank][j]
            mine.append(shard_t[b:b + n].detach().cpu())
        gathered = self.transport.all_gather_object(mine)
        full = [torch.empty(v, Dw, dtype=torch.float32) for v in self.num_rows]
        for j, p in enumerate(self.sh):
            f, first, cnt = self.parts[p][:3]
            for r in range(self.world):
                fr, n = owned_rows(cnt, j, r, self.world)
                if n:
                    full[f][first + fr:first + cnt:self.world] = gathered[r][j].view(-1, Dw)
        for p in self.rp:
            f, first, cnt = self.parts[p][:3]
            o = self.rep_off[p]
            full[f][first:first + cnt] = rep_t[o:o + cnt].detach().cpu().view(cnt, Dw)
        return full

    def export_full_tables(self, w: int = 0):
        """Collective.  Every rank gets the full ``[V_f, dims[w]]`` table of every feature (CPU tensors, feature order), i.e.
        what ``model.embeddings[name].weight`` holds in the reference (``torchctr/trainer.py:353-496``)."""
        nw = len(self.dims)
        return self._gather_width(self.shards[w].data, self.shards[nw + w].data, w)

    def export_full_optimizer_state(self, w: int = 0):
        nw = len(self.dims)
        if self.opt_state[w] is None:
            self.transport.all_gather_object(None)
            return None, None
        return self._gather_width(self.opt_state[w], self.opt_state[nw + w], w), None

    def _scatter_width(self, full, shard_t, rep_t, wi):
        Dw = self.dims[wi]
        for f, t in enumerate(full):
            if tuple(t.shape) != (self.num_rows[f], Dw):
                raise ValueError(f"table {f}: expected {(self.num_rows[f], Dw)}, got {tuple(t.shape)}")
        for j, p in enumerate(self.sh):
            f, first, cnt = self.parts[p][:3]
            fr, n = owned_rows(cnt, j, self.rank, self.world)
            if n:
                b = self.base[self.rank][j]
                shard_t[b:b + n].copy_(full[f][first + fr:first + cnt:self.world].to(self.device).view(n, -1))
        for p in self.rp:
            f, first, cnt = self.parts[p][:3]
            o = self.rep_off[p]
            rep_t[o:o + cnt].copy_(full[f][first:first + cnt].to(self.device).view(cnt, -1))

    def load_full_tables(self, full, w: int = 0) -> None:
        """Scatter full tables (one ``[V_f, dims[w]]`` tensor per feature, e.g. from a reference checkpoint) into this rank's
        shard and replicated block.  Every rank calls it with the same tensors; optimizer state is reset."""
        nw = len(self.dims)
        with torch.no_grad():
            self._scatter_width(full, self.shards[w].data, self.shards[nw + w].data, w)
        self.opt_state = [None] * (2 * nw)
        self.transport.barrier()

    def load_full_optimizer_state(self, state, w: int = 0) -> None:
        full = state[0]
        if full is None:
            return
        nw = len(self.dims)
        if self.opt_state[0] is None:
            self.opt_state = [torch.zeros_like(p.data) for p in self.shards]
        self._scatter_width(full, self.opt_state[w], self.opt_state[nw + w], w)
