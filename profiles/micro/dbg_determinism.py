"""Two identical models, same batch, one step each: which tensors are not bit-identical?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_reference_trainer as T
from torchctr_b200.models import DNN
torch.backends.cuda.matmul.allow_tf32 = False
fc = T._feat_configs()
gen = torch.Generator().manual_seed(5)
train = T._batches(gen, 3, 256, fc)
def make():
    torch.manual_seed(0)
    m = T._no_dropout(DNN(fc, [32, 16])).cuda().train()
    o = torch.optim.Adagrad(m.dense_parameters(), lr=0.05)
    m.bind_optimizer(o, kind="adagrad")
    return m, o
a, oa = make(); b, ob = make()
for step in range(3):
    for m, o in ((a, oa), (b, ob)):
        o.zero_grad(); m.training_step(train[step], step).backward()
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        if p.grad is not None and not torch.equal(p.grad, q.grad):
            print("step", step, "grad differs", n, float((p.grad - q.grad).abs().max()), float(p.grad.abs().max()))
    oa.step(); ob.step()
    for k, v in a.state_dict().items():
        if not torch.equal(v, b.state_dict()[k]):
            print("step", step, "state differs", k, float((v - b.state_dict()[k]).abs().max()))
print("done")
