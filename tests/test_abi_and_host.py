"""CPU-side checks (no GPU, no compute calls): the C-ABI library builds, loads and exports every
symbol include/ctr_b200.h declares; struct layouts agree between the header and the ctypes
mirror; host-side argument validation and model plumbing behave like the reference."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ctr_b200.h")


def declared_in_header():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctr_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from torchctr_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    names = declared_in_header()
    assert len(names) >= 20
    for name in names:
        assert hasattr(handle, name), f"{name} declared in ctr_b200.h but not exported"
    assert set(_lib.declared_symbols()) == set(names), "ctypes signature table out of sync with the header"
    assert _lib.lib().ctr_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof / offsetof with ctypes."""
    from torchctr_b200 import _lib
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "ctr_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(ctr_feature_t), sizeof(ctr_group_t), sizeof(ctr_opt_t), sizeof(ctr_vocab_map_t), sizeof(ctr_hyper_t));
  printf("%zu %zu %zu %zu\n", offsetof(ctr_feature_t, num_rows), offsetof(ctr_feature_t, hash_seed), offsetof(ctr_group_t, status), offsetof(ctr_opt_t, device_hyper));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = [ctypes.sizeof(x) for x in (_lib.Feature, _lib.Group, _lib.Opt, _lib.VocabMap, _lib.Hyper)]
    offs = [_lib.Feature.num_rows.offset, _lib.Feature.hash_seed.offset, _lib.Group.status.offset, _lib.Opt.device_hyper.offset]
    assert [int(v) for v in out] == sizes + offs


def test_no_cpu_fallback_and_reference_error_messages():
    from torchctr_b200.models import DNN, DeepFM
    from torchctr_b200.models.base import split_feature_configs
    fc = [{"name": "a", "type": "sparse", "num_embeddings": 10, "emb_dim": 16}, {"name": "d", "type": "dense"},
          {"name": "s", "type": "dense", "islist": True}]
    sparse, dense_width = split_feature_configs(fc)
    assert len(sparse) == 1 and dense_width == 4              # dnn.py:24-27: 3 per dense list feature
    with pytest.raises(ValueError, match="emb_dim must be specified for sparse features."):
        DNN([{"name": "a", "type": "sparse", "num_embeddings": 3}])
    with pytest.raises(ValueError, match="Unsupported feature type: weird"):
        DNN([{"name": "a", "type": "weird"}])
    m = DNN(fc[:2], [8])
    assert list(m.state_dict().keys())[:1] == ["embeddings.a.weight"]
    assert m.feat_configs is fc[:2] or m.feat_configs == fc[:2]
    with pytest.raises(RuntimeError, match="no CPU path"):
        m({"a": torch.zeros(4, 1, dtype=torch.long), "dense_features": torch.zeros(4, 1)})
    with pytest.raises(ValueError, match="common emb_dim"):
        DeepFM([{"name": "a", "type": "sparse", "num_embeddings": 3, "emb_dim": 4},
                {"name": "b", "type": "sparse", "num_embeddings": 3, "emb_dim": 8}])


def test_state_dict_keys_match_oracle_models():
    from oracle import models as om
    from torchctr_b200.models import DCNv2, DeepFM, DNN
    fc = [{"name": "a", "type": "sparse", "num_embeddings": 10, "emb_dim": 8},
          {"name": "b", "type": "sparse", "num_embeddings": 7, "emb_dim": 8}, {"name": "d", "type": "dense"}]
    for ours, ref in ((DNN(fc, [8, 4]), om.OracleDNN(fc, [8, 4])), (DeepFM(fc, [8]), om.OracleDeepFM(fc, [8])),
                      (DCNv2(fc, [8], 2), om.OracleDCNv2(fc, [8], 2))):
        assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
        ours.load_state_dict(ref.state_dict())


def test_optimizer_binding_reads_torch_hyperparameters():
    from torchctr_b200 import ops
    from torchctr_b200.models import DNN
    fc = [{"name": "a", "type": "sparse", "num_embeddings": 10, "emb_dim": 8}, {"name": "d", "type": "dense"}]
    m = DNN(fc, [4])
    opt = torch.optim.Adagrad(m.parameters(), lr=0.5, lr_decay=0.1, eps=1e-9)
    b = m.bind_optimizer(opt)._lookup.binding
    assert b.kind == "adagrad"
    o1, o2 = b.next_opt(), b.next_opt()
    assert o1.step == 1 and abs(o1.lr - 0.5) < 1e-12 and abs(o2.lr - 0.5 / 1.1) < 1e-12 and o1.eps == 1e-9
    opt.param_groups[0]["lr"] = 0.25                             # an lr scheduler's effect is picked up
    assert abs(b.next_opt().lr - 0.25 / 1.2) < 1e-12
    with pytest.raises(ValueError, match="weight_decay"):
        m.bind_optimizer(torch.optim.SGD(m.parameters(), lr=0.1, weight_decay=1e-3))
    adam = ops.opt_hyper(ops.make_opt("adam", lr=0.01, eps=1e-8, betas=(0.9, 0.999), step=3))
    import math
    assert abs(adam[4] - 0.01 * math.sqrt(1 - 0.999 ** 3) / (1 - 0.9 ** 3)) < 1e-9
    assert adam[2] == pytest.approx(0.1, rel=1e-6) and adam[3] == pytest.approx(0.001, rel=1e-6)


def test_graft_entry_build():
    import __graft_entry__ as g
    g.build()


def test_peer_shard_geometry_is_a_bijection():
    """Row r of table f -> (owner (r + f) % P, adj[o, f] + (r + f) // P): every row lands on exactly one slot of
    its owner's fused shard, inside that table's range, in row order (what the kernels' shard_vrow computes)."""
    from torchctr_b200.parallel.peer import owned_rows, shard_geometry
    for world in (1, 2, 3, 4, 8, 16):
        Vs = [37, 101, 2, 1000, 5, 1]
        base, total, adj = shard_geometry(Vs, world)
        F = len(Vs)
        for o in range(world):
            assert total[o] == base[o][-1] + max(owned_rows(Vs[-1], F - 1, o, world)[1], 1)
        for f, v in enumerate(Vs):
            seen = set()
            for r in range(v):
                o = (r + f) % world
                vr = int(adj[o * F + f]) + (r + f) // world
                fr, n = owned_rows(v, f, o, world)
                assert base[o][f] <= vr < base[o][f] + n
                assert (r - fr) % world == 0 and vr - base[o][f] == (r - fr) // world
                assert (o, vr) not in seen
                seen.add((o, vr))
            assert len(seen) == v


def test_fused_adagrad_on_cpu_parameters_is_torch_adagrad():
    """Off the GPU (or with lr_decay / weight decay / sparse gradients) FusedAdagrad must take torch's own code path."""
    from torchctr_b200.optim import FusedAdagrad
    gen = torch.Generator().manual_seed(0)
    p1 = [torch.randn(5, 3, generator=gen, requires_grad=True), torch.randn(4, generator=gen, requires_grad=True)]
    p2 = [p.detach().clone().requires_grad_(True) for p in p1]
    o1 = FusedAdagrad(p1, lr=0.1, lr_decay=0.01)
    o2 = torch.optim.Adagrad(p2, lr=0.1, lr_decay=0.01)
    for _ in range(3):
        for a, b in zip(p1, p2):
            g = torch.randn(a.shape, generator=gen)
            a.grad, b.grad = g.clone(), g.clone()
        o1.step()
        o2.step()
    for a, b in zip(p1, p2):
        assert torch.equal(a, b)
        assert torch.equal(o1.state[a]["sum"], o2.state[b]["sum"]) and float(o1.state[a]["step"]) == float(o2.state[b]["step"])


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md must say, for every symbol of the header, which reference lines it stands in for."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_in_header() if n not in doc]
    assert not missing, missing


def test_dropout_seed_is_snapshotted_per_forward():
    """backward must see the seed value ITS forward used (fwd, fwd, bwd orders): every training forward gets its own
    copy of the device counter instead of a reference to the tensor that the next forward increments in place."""
    import torch
    from torchctr_b200.models import DNN
    m = DNN([{"name": "a", "type": "sparse", "num_embeddings": 10, "emb_dim": 16}, {"name": "d", "type": "dense"}], [16])
    s1 = m._step_seed(torch.device("cpu"))
    s2 = m._step_seed(torch.device("cpu"))
    assert s1.data_ptr() != s2.data_ptr() and int(s2) == int(s1) + 1
    m._step_seed(torch.device("cpu"))
    assert int(s2) == int(s1) + 1                      # later forwards do not change what earlier ones hold


def test_first_block_preparation_host_logic(monkeypatch):
    """``nn.tower.prepare_first_block`` / ``CTRModelBase._take_prepared`` with the CUDA stream calls stubbed out (CPU tensors):
    the prepared tensors are exactly what the in-line path builds (F.pad of the weight, its transpose, counter + 1), and a
    preparation that does not fit the tower input is joined and dropped, never used."""
    import contextlib

    import torch.nn as nn
    import torch.nn.functional as F
    from torchctr_b200.models import DNN
    from torchctr_b200.nn import tower as T

    class FakeStream:
        waited = None

        def wait_stream(self, other):
            pass

        def wait_event(self, ev):
            self.waited = ev

    class FakeEvent:
        def record(self, stream):
            self.stream = stream

    main, side = FakeStream(), FakeStream()
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: main)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(T, "_wgrad_stream", lambda device: side)
    lin = nn.Linear(429, 256)
    counter = torch.tensor([5])
    snap, w, wt, ready = T.prepare_first_block(lin, 3, counter, want_wt=True)
    assert counter.item() == 6 and snap.item() == 6 and ready.stream is side
    assert torch.equal(w, F.pad(lin.weight.detach(), (0, 3))) and torch.equal(wt, w.t().contiguous())
    snap, w, wt, _ = T.prepare_first_block(lin, 0, counter, want_wt=False)
    assert wt is None and w.data_ptr() == lin.weight.data_ptr() and snap.item() == 7

    fc = [{"name": "a", "type": "sparse", "num_embeddings": 10, "emb_dim": 4}, {"name": "d", "type": "dense"}]
    m = DNN(fc, [8, 4])
    layers = list(m.tower)[:-1]
    x = torch.zeros(2, 8)                              # 4 + 1 columns padded to 8: the first block sees 3 padding columns
    assert m._take_prepared(x, layers) is None and m._first_block_pad == 3
    m._prepare_tower()                                 # tables / tower on the CPU: nothing is prepared
    assert m._prepared is None
    made = T.prepare_first_block(m.tower[0], 3, m._seed_counter(torch.device("cpu")), True)
    entry = (made, 3, T.matmul_precision(), m.tower[0].weight._version, m.tower[0])
    m._prepared = entry
    assert m._take_prepared(x, layers) is made and m._prepared is None
    m._prepared = entry
    assert m._take_prepared(torch.zeros(2, 12), layers) is None and main.waited is made[3]      # other width: joined, dropped
    m._prepared = entry
    with torch.no_grad():
        m.tower[0].weight.mul_(1.0)                    # an in-place update after the preparation (version bump): stale, dropped
    assert m._take_prepared(x, layers) is None
