"""Run bench-like epochs and report the first epoch at which the status word becomes non-zero."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from bench import lorenz_poisson
from vjf_b200.model import VJF
B = int(os.environ.get("PB", 4096)); T = 256; D = 200
torch.manual_seed(1234)
m = VJF.make_model(D, 3, 0, 50, [64], "poisson", max_trials=B, seed=99)
y = lorenz_poisson(T, B, D, seed=1000, device="cuda")
for ep in range(int(os.environ.get("EPOCHS", 40))):
    mu, lv, ls = m.run(y)
    st = m.status()
    P = m.w_precision
    print(ep, "loss %.4f" % ls[:, 0].mean().item(), "status", st, "tr_logvar %.4f" % m.transition.logvar.item(),
          "max|W| %.3f" % m.w_mean.abs().max().item(), "lv [%.2f, %.2f]" % (lv.min().item(), lv.max().item()),
          "mu max %.2f" % mu.abs().max().item(), "P max %.3e" % P.abs().max().item(),
          "nonfinite losses", int((~torch.isfinite(ls)).sum().item()), "first bad step", (~torch.isfinite(ls[:,0])).nonzero()[:1].flatten().tolist(), flush=True)
    if st:
        bad = (ls[:, 2] == 0).nonzero().flatten()[:5].tolist()
        print("  steps with zeroed dyn term:", bad, "loss rows:", ls[bad[0] - 1:bad[0] + 2].tolist() if bad else None)
        break
