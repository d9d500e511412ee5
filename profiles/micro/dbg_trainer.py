import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_reference_trainer as T
from baseline import refshim
ref = refshim.load_reference()
torch.backends.cuda.matmul.allow_tf32 = False
fc = T._feat_configs()
gen = torch.Generator().manual_seed(3)
train, evalb = T._batches(gen, 6, 512, fc), T._batches(gen, 2, 512, fc)
theirs, ours = T._pair(ref, fc)
ot = torch.optim.Adagrad(theirs.parameters(), lr=0.05); oo = torch.optim.Adagrad(ours.parameters(), lr=0.05)
theirs.train(); ours.train()
for k, b in enumerate(train[:3]):
    ot.zero_grad(); oo.zero_grad()
    lt = theirs.training_step(b, k); lo = ours.training_step(b, k)
    lt.backward(); lo.backward(); ot.step(); oo.step()
    print(k, lt.item(), lo.item())
sd_o, sd_t = ours.state_dict(), theirs.state_dict()
for k in sd_t:
    if sd_t[k].dtype.is_floating_point:
        err = float((sd_o[k].cpu() - sd_t[k]).abs().max()); print(f"{k:32s} err={err:.3e} max={float(sd_t[k].abs().max()):.3e}")
    else:
        print(k, sd_o[k].item(), sd_t[k].item())
theirs.eval(); ours.eval()
with torch.no_grad():
    a = theirs(evalb[0][0]); b = ours(evalb[0][0]).cpu()
print("eval logits err", float((a - b).abs().max()), float(a.abs().max()))
