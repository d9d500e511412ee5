"""Worker of the multi-rank tests (spawned by tests/test_sharded_*.py with torchrun-style env).

Checks ``torchctr_b200.parallel.ShardedTables`` against the single-process oracle on the GLOBAL batch:
forward = each rank's pooled output equals the oracle lookup on its local batch; after backward with a
known grad_out the union of shards equals the oracle's touched-row update with all ranks' gradients.

backend "cuda": NCCL + libctr_b200 kernels (the product path).  backend "cpu": gloo + a torch-CPU
stand-in for the five device steps, which exists ONLY here to exercise the exchange logic without a GPU.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import embedding as oe          # noqa: E402
from oracle import optim as oo              # noqa: E402
from torchctr_b200.nn.embedding import EmbeddingTable          # noqa: E402
from torchctr_b200.parallel.sharded import ShardedTables, local_rows, shard_bases   # noqa: E402


class TorchCpuBackend:
    """Test stand-in for CudaBackend (same five methods, plain torch on the host)."""

    def route(self, st, ids_list):
        P = st.world
        base = st.base_dev.view(P, -1)
        owners, vrows = [], []
        for f, ids in enumerate(ids_list):
            flat = ids.reshape(-1)
            valid = flat >= 0
            row = flat.clamp(min=0)
            o = torch.where(valid, row % P, torch.full_like(row, P))
            v = torch.where(valid, base[(row % P), f] + row // P, torch.full_like(row, -1))
            owners.append(o); vrows.append(v)
        owner, vrow = torch.cat(owners), torch.cat(vrows)
        order = torch.sort(owner, stable=True).indices
        counts = torch.bincount(owner, minlength=P + 1)
        n = int(counts[:P].sum())
        send_rows = vrow[order]
        inv = torch.full_like(owner, -1)
        inv[order[:n]] = torch.arange(n)
        invs, off = [], 0
        for ids in ids_list:
            invs.append(inv[off:off + ids.numel()].view(ids.shape)); off += ids.numel()
        return counts, send_rows, invs, (order, ids_list)

    def gather(self, rows, shard):
        return shard[rows]

    def pool(self, st, w, invs, got, dense):
        parts = [oe.pooled_lookup(inv, got if got.shape[0] else torch.zeros(1, st.dims[w])) for inv in invs]
        if dense is not None:
            parts.append(dense)
        out = torch.cat(parts, 1)
        pad = (-out.shape[1]) % 4
        return torch.nn.functional.pad(out, (0, pad))

    def grad_gather(self, st, w, handle, grad_out, n_send):
        order, ids_list = handle
        D = st.dims[w]
        rows = []
        for f, ids in enumerate(ids_list):
            B, L = ids.shape
            rows.append(grad_out[:, f * D:(f + 1) * D].unsqueeze(1).expand(B, L, D).reshape(B * L, D))
        return torch.cat(rows)[order[:n_send]].contiguous()

    def update(self, shard_mod, rows, grads, binding, cache=None):
        opt = binding.next_opt()
        uniq, inverse = torch.unique(rows, return_inverse=True)
        g = torch.zeros(uniq.numel(), grads.shape[1]).index_add_(0, inverse, grads)
        oo.sgd_rows(shard_mod.weight.data, uniq, g, opt.lr)


def main():
    backend = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if backend in ("cuda", "peer"):
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dev = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    gen = torch.Generator().manual_seed(5)              # identical on every rank: the GLOBAL problem
    Vs, Ls, D, B = [37, 101, 8, 1000], [1, 6, 1, 3], 16, 200
    full16 = [torch.randn(v, D, generator=gen) for v in Vs]
    full1 = [torch.randn(v, 1, generator=gen) for v in Vs]
    ids_all, gout16_all, gout1_all = [], [], []
    for r in range(world):
        ids = []
        for v, L in zip(Vs, Ls):
            t = torch.randint(0, v, (B, L), generator=gen)
            if L > 1:
                t[torch.rand(B, L, generator=gen) < 0.3] = -100
            t[:, 0] = torch.randint(0, min(v, 3), (B,), generator=gen)       # hot rows owned by few ranks
            ids.append(t)
        ids_all.append(ids)
        gout16_all.append(torch.randn(B, len(Vs) * D + 4, generator=gen))
        gout1_all.append(torch.randn(B, 4, generator=gen))
    dense = torch.randn(B, 3, generator=gen)

    names = [f"f{i}" for i in range(len(Vs))]
    tabs16 = [EmbeddingTable(v, D, _weight=w.clone()) for v, w in zip(Vs, full16)]
    tabs1 = [EmbeddingTable(v, 1, _weight=w.clone()) for v, w in zip(Vs, full1)]
    if backend == "peer":          # rows read / gradients pulled through CUDA IPC peer mappings, no all-to-all
        from torchctr_b200.parallel.peer import IpcTransport, PeerShardedTables
        st = PeerShardedTables(names, [tabs16, tabs1], IpcTransport(None, dev), dev).train()
        shard_params = list(st.shards)
    else:
        with torch.device(dev):
            st = ShardedTables(names, [tabs16, tabs1], None, TorchCpuBackend() if backend == "cpu" else None)
        st = st.to(dev)
        shard_params = [s.weight for s in st.shards]
    opt = torch.optim.SGD(shard_params, lr=0.5)
    st.bind_optimizer(opt, kind="sgd")

    feats = {n: t for n, t in zip(names, ids_all[rank])}
    for p in shard_params:
        p.requires_grad_(True)
    x, first = st(feats, dense.to(dev))
    ref_x = torch.cat([oe.pooled_lookup(i, w) for i, w in zip(ids_all[rank], full16)] + [dense], 1)
    ref_1 = torch.cat([oe.pooled_lookup(i, w) for i, w in zip(ids_all[rank], full1)], 1)
    width = ref_x.shape[1]
    assert x.shape[1] % 4 == 0 and torch.allclose(x[:, :width].cpu(), ref_x, rtol=1e-5, atol=1e-6), "forward D=16"
    assert torch.allclose(first[:, :len(Vs)].cpu(), ref_1, rtol=1e-5, atol=1e-6), "forward D=1"

    g16 = gout16_all[rank][:, :x.shape[1]].contiguous().to(dev)
    g16 = torch.nn.functional.pad(g16, (0, x.shape[1] - g16.shape[1])) if g16.shape[1] < x.shape[1] else g16
    torch.autograd.backward([x, first], [g16, gout1_all[rank][:, :first.shape[1]].contiguous().to(dev)])
    if dev.type == "cuda":
        torch.cuda.synchronize()

    # oracle: SGD on the touched rows with the gradients of ALL ranks
    base = shard_bases(Vs, world)
    for w, (fulls, gouts, Dw) in enumerate(((full16, gout16_all, D), (full1, gout1_all, 1))):
        for f, (v, wt) in enumerate(zip(Vs, fulls)):
            dense_grad = torch.zeros(v, Dw)
            for r in range(world):
                dense_grad += oe.dense_table_grad(ids_all[r][f], gouts[r][:, f * Dw:(f + 1) * Dw], v)
            expect = wt - 0.5 * dense_grad
            if backend == "peer":
                fr, rows = st.local_rows_of(w, f)
                got, want = rows.cpu(), expect[fr::world]
            else:
                n = local_rows(v, rank, world)
                b = int(base[rank, f])
                got, want = st.shards[w].weight.detach().cpu()[b:b + n], expect[rank::world]
            assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-5, atol=2e-5), f"update width {Dw} table {f}"
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK", backend, world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
