"""Row-sharded embedding tables with an all-to-all exchange.

The reference is replicas-only: ``Accelerator().prepare`` wraps the model in DDP
(``torchctr/trainer.py:128-130``), every rank holds every table and dense ``[V, D]`` gradients are
all-reduced.  Here one process per GPU holds ``1/P`` of every table (``owner(row) = row mod P``,
all tables of a group fused into one shard per rank), the batch stays data-parallel, and a step is

  forward   route ids -> all_gather(counts) -> all_to_all(rows) -> owner gather -> all_to_all(vectors)
            -> local pool + concat (K1 over the received vectors)
  backward  gather grad_out per sent vector -> all_to_all(grads) -> owner-side K2 (dedup across ALL
            requesters + fused optimizer update of the shard)

The dense tower is replicated; its gradients are summed with one all-reduce (``reduce_dense_grads``)
after ``(loss / P).backward()``, which also gives the table gradients the 1/P of a global-batch mean.

The collectives go through ``torch.distributed`` (NCCL over NVLink on GPUs).  The device work is behind
a five-method backend so that the exchange logic is exercised on CPU with gloo in the tests
(``tests/test_sharded_gloo.py`` supplies a torch-CPU stand-in; the product backend below has no CPU path).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops
from ..nn.embedding import EmbeddingTable, SparseOptimizerBinding, _Workspace, _layout


def local_rows(num_rows: int, rank: int, world: int) -> int:
    return (num_rows - rank + world - 1) // world


def shard_bases(num_rows_list, world: int) -> torch.Tensor:
    """base[o, f] = first virtual row of table f inside owner o's fused shard (i64 [world, F])."""
    base = torch.zeros(world, len(num_rows_list), dtype=torch.int64)
    for o in range(world):
        acc = 0
        for f, v in enumerate(num_rows_list):
            base[o, f] = acc
            acc += local_rows(v, o, world)
    return base


class CudaBackend:
    """Device work of the sharded lookup on libctr_b200 kernels."""

    def route(self, st, ids_list):
        B = ids_list[0].shape[0]
        specs = [ops.FeatureSpec(ids=ids, table=None, num_rows=v, D=st.dims[0], out_col=c, index_kind=k, hash_seed=s)
                 for ids, v, c, k, s in zip(ids_list, st.num_rows, st.cols(0), st.index_kinds, st.hash_seeds)]
        S = sum(i.numel() for i in ids_list)
        call = ops.make_group(specs, B, None, st.cols(0)[-1] + st.dims[0])
        counts, send_rows, inv, ws = ops.route_build(call, st.world, st.base_dev, S)
        invs, off = [], 0
        for ids in ids_list:
            invs.append(inv[off:off + ids.numel()].view(ids.shape))
            off += ids.numel()
        return counts, send_rows, invs, (ws, ids_list)

    def gather(self, rows, shard):
        return ops.rows_gather(rows, shard)

    def pool(self, st, w, invs, got, dense):
        B = invs[0].shape[0]
        D = st.dims[w]
        n = max(got.shape[0], 1)
        if got.shape[0] == 0:
            got = torch.zeros(1, D, device=got.device)
        cols = st.cols(w)
        width = cols[-1] + D + (0 if dense is None else dense.shape[1])
        stride = (width + 3) // 4 * 4
        out = torch.empty(B, stride, dtype=torch.float32, device=got.device)
        specs = [ops.FeatureSpec(ids=inv, table=got, num_rows=n, D=D, out_col=c) for inv, c in zip(invs, cols)]
        call = ops.make_group(specs, B, out, stride, dense=dense, dense_col=cols[-1] + D,
                              zero_from=width if stride > width else -1)
        ops.emb_pool_fwd(call)
        return out

    def grad_gather(self, st, w, handle, grad_out, n_send):
        ws, ids_list = handle
        B = ids_list[0].shape[0]
        specs = [ops.FeatureSpec(ids=ids, table=None, num_rows=v, D=st.dims[w], out_col=c, index_kind=k, hash_seed=s)
                 for ids, v, c, k, s in zip(ids_list, st.num_rows, st.cols(w), st.index_kinds, st.hash_seeds)]
        call = ops.make_group(specs, B, grad_out, grad_out.shape[1])
        return ops.route_grad_gather(call, st.world, ws, n_send, st.dims[w])

    def update(self, shard_mod, rows, grads, binding, cache=None):
        """Owner side of the backward: K2 over the received (row, gradient) pairs.  ``cache`` lets the
        widths of one step, which received the same rows, share one sort."""
        n = rows.shape[0]
        if n == 0:
            return
        shard_mod._ensure_state(binding.kind, binding.initial_accumulator_value())
        spec = ops.FeatureSpec(ids=rows.view(n, 1), table=shard_mod.weight.data, num_rows=shard_mod.num_embeddings,
                               D=shard_mod.embedding_dim, out_col=0, state0=getattr(shard_mod, "opt_state0", None),
                               state1=getattr(shard_mod, "opt_state1", None))
        call = ops.make_group([spec], n, grads, grads.shape[1])
        if cache is not None and "ws" in cache:
            ws = cache["ws"]
        else:
            widest = ops.FeatureSpec(ids=rows.view(n, 1), table=None, num_rows=shard_mod.num_embeddings,
                                     D=(cache or {}).get("max_dim", shard_mod.embedding_dim), out_col=0)
            need = ops.emb_bwd_workspace_bytes(ops.make_group([widest], n, None, widest.D))
            ws = torch.empty(need + 256, dtype=torch.uint8, device=grads.device) if cache is not None \
                else _Workspace.get(grads.device, need)
            ops.emb_bwd_plan(call, ws, runs=False)
            if cache is not None:
                cache["ws"] = ws
        ops.emb_bwd_apply(call, ws, binding.next_opt())


class _ShardedLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, st, feats_ids, dense, *shards):
        be, pg, P, rank = st.backend, st.pg, st.world, st.rank
        counts, send_rows, invs, handle = be.route(st, feats_ids)
        matrix = torch.empty(P * (P + 1), dtype=torch.int64, device=counts.device)
        dist.all_gather_into_tensor(matrix, counts, group=pg)
        matrix = matrix.view(P, P + 1).cpu()                       # the one host read of the exchange
        send_splits = matrix[rank, :P].tolist()
        recv_splits = matrix[:, rank].tolist()
        n_send, n_recv = sum(send_splits), sum(recv_splits)
        recv_rows = torch.empty(n_recv, dtype=torch.int64, device=counts.device)
        dist.all_to_all_single(recv_rows, send_rows[:n_send].contiguous(), recv_splits, send_splits, group=pg)
        outs = []
        for w, shard in enumerate(shards):
            vec = be.gather(recv_rows, shard.detach())
            got = torch.empty(n_send, st.dims[w], dtype=torch.float32, device=vec.device)
            dist.all_to_all_single(got, vec, send_splits, recv_splits, group=pg)
            outs.append(be.pool(st, w, invs, got, dense if w == 0 else None))
        ctx.st, ctx.handle, ctx.recv_rows = st, handle, recv_rows
        ctx.splits = (send_splits, recv_splits, n_send, n_recv)
        ctx.has_dense = dense is not None
        ctx.dense_width = 0 if dense is None else dense.shape[1]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        st = ctx.st
        be, pg = st.backend, st.pg
        send_splits, recv_splits, n_send, n_recv = ctx.splits
        cache = {"max_dim": max(st.dims)}
        for w, g in enumerate(grad_outs):
            if g is None:
                continue
            g = g.contiguous()
            g_send = be.grad_gather(st, w, ctx.handle, g, n_send)
            g_recv = torch.empty(n_recv, st.dims[w], dtype=torch.float32, device=g.device)
            dist.all_to_all_single(g_recv, g_send, recv_splits, send_splits, group=pg)
            if st.bindings[w] is None:
                raise RuntimeError("sharded tables need bind_optimizer(): there is no dense or sparse .grad to hand back")
            be.update(st.shards[w], ctx.recv_rows, g_recv, st.bindings[w], cache)
        gdense = None
        if ctx.has_dense and ctx.needs_input_grad[2] and grad_outs[0] is not None:
            c0 = st.cols(0)[-1] + st.dims[0]
            gdense = grad_outs[0][:, c0:c0 + ctx.dense_width]
        return (None, None, gdense) + (None,) * len(st.shards)


class ShardedTables(nn.Module):
    """The tables of one lookup group, row-sharded over ``pg``.  ``dims`` lists the widths that share the
    ids (DeepFM: ``[emb_dim, 1]``); width ``w`` of table ``f`` is ``full_tables[w][f]`` (``[V_f, dims[w]]``)."""

    def __init__(self, names, full_tables, pg=None, backend=None):
        super().__init__()
        self.pg = pg
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        self.backend = backend or CudaBackend()
        self.names = list(names)
        first = full_tables[0]
        self.num_rows = [int(t.num_embeddings) for t in first]
        self.index_kinds = [t.index_kind for t in first]
        self.hash_seeds = [t.hash_seed for t in first]
        for t in first:
            if t.pooling != "sum" or t.use_id_weight or t.index_kind == "vocab":
                raise NotImplementedError("sharded lookup v1: sum pooling, direct / hashed ids")
        self.dims = [int(tabs[0].embedding_dim) for tabs in full_tables]
        base = shard_bases(self.num_rows, self.world)
        self.register_buffer("base_dev", base.reshape(-1).clone(), persistent=False)
        total = int(sum(local_rows(v, self.rank, self.world) for v in self.num_rows))
        self.shards = nn.ModuleList()
        for tabs, D in zip(full_tables, self.dims):
            shard = EmbeddingTable(max(total, 1), D)
            with torch.no_grad():
                for f, t in enumerate(tabs):
                    n = local_rows(self.num_rows[f], self.rank, self.world)
                    b = int(base[self.rank, f])
                    shard.weight[b:b + n] = t.weight.detach()[self.rank::self.world].to(shard.weight.device)
            self.shards.append(shard)
        self.bindings = [None] * len(self.dims)

    def cols(self, w):
        return [f * self.dims[w] for f in range(len(self.names))]

    def bind_optimizer(self, optimizer, kind=None):
        self.bindings = [SparseOptimizerBinding(optimizer, [s], kind) for s in self.shards]

    def forward(self, feats, dense=None):
        dev = self.shards[0].weight.device
        ids = []
        for n in self.names:
            t = feats[n]
            if t.dim() == 1:
                t = t.unsqueeze(1)
            ids.append(t.to(dev, dtype=torch.int64, non_blocking=True).contiguous())
        if dense is not None:
            dense = dense.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        return _ShardedLookupFn.apply(self, ids, dense, *[s.weight for s in self.shards])


def reduce_dense_grads(params, pg=None):
    """Sum the replicated (tower) gradients over the ranks in one flat all-reduce."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=pg)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
