"""Development check of the tile pipeline against the fp64 oracle: prints the errors instead of asserting.
   python scripts/dev_tile.py [B] [T] [lik] [D] [d] [R] [H]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vjf_oracle as O  # noqa: E402
from vjf_b200.model import VJF  # noqa: E402

a = sys.argv[1:]
B = int(a[0]) if len(a) > 0 else 200
T = int(a[1]) if len(a) > 1 else 3
lik = a[2] if len(a) > 2 else "poisson"
D = int(a[3]) if len(a) > 3 else 200
d = int(a[4]) if len(a) > 4 else 3
R = int(a[5]) if len(a) > 5 else 50
H = [int(a[6])] if len(a) > 6 else [64]
rng = np.random.default_rng(5)
torch.manual_seed(5)
m = VJF.make_model(D, d, 0, R, H, lik, lr=1e-3, max_trials=B, seed=4321)
o = O.OracleVJF(D, d, 0, R, H, lik, lr=1e-3, dtype=np.float64)
st0 = {k: v.detach().cpu().numpy() for k, v in m.full_state().items()}
o.set_state(st0)
if lik == "poisson":
    t = np.arange(T)[:, None, None] * 0.05
    ph = rng.uniform(0, 2 * np.pi, (1, B, d))
    x = np.sin(t * (1 + np.arange(d)) + ph)
    Cm = rng.normal(size=(d, D)) / np.sqrt(d)
    y = rng.poisson(np.exp(np.clip(x @ Cm - 1.0, None, 3.0))).astype(np.float32)
else:
    y = rng.normal(size=(T, B, D)).astype(np.float32)
eps = rng.normal(size=(T, 2, B, d)).astype(np.float32)
t0 = time.time()
mu, lv, losses = m.run(torch.as_tensor(y), None, None, eps=torch.as_tensor(eps))
torch.cuda.synchronize()
print(f"gpu run {time.time() - t0:.3f}s status {m.status()}")
omu, olv, ol = o.run(y.astype(np.float64), None, eps=eps.astype(np.float64))


def err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()), float(np.abs(b).max())


for tt in range(T):
    print(f"t={tt}: mu {err(mu[tt].cpu().numpy(), omu[tt])}  logvar {err(lv[tt].cpu().numpy(), olv[tt])}  losses gpu {losses[tt].cpu().numpy()} oracle {ol[tt]}")
got = {k: v.detach().cpu().numpy() for k, v in m.full_state().items()}
want = o.get_state()
for k, v in want.items():
    if k in got:
        print(f"  {k:45s} err {err(got[k], v)}")
