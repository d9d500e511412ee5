import sys, traceback, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_models import _criteo_like, no_dropout
from torchctr_b200.graph import GraphedTrainStep
from torchctr_b200.models import DeepFM
from torchctr_b200.nn import embedding as E
orig = E._LookupCall.run_backward
def wrapped(self, *a, **k):
    try:
        return orig(self, *a, **k)
    except Exception:
        traceback.print_exc()
        raise
E._LookupCall.run_backward = wrapped
from torchctr_b200.nn import tower as T
ob = T._TowerBlockFn.backward
def tb(ctx, gy):
    try:
        return ob(ctx, gy)
    except Exception:
        traceback.print_exc(); raise
T._TowerBlockFn.backward = staticmethod(tb)
gen = torch.Generator().manual_seed(17)
fc, f0, l0 = _criteo_like(gen, 1024, 5, 300, 16, 3)
torch.manual_seed(1)
b = no_dropout(DeepFM(fc, [32, 16])).cuda().train()
o = torch.optim.Adagrad(b.dense_parameters(), lr=0.05)
b.bind_optimizer(o, kind="adagrad")
g = GraphedTrainStep(b, o, (f0, l0), warmup=1)
print("captured ok", g(( f0, l0)).item())
