"""Short program for ncu: the tile pipeline at the headline shape (C2, 4096 trials) and in the throughput regime (65536 trials)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vjf_b200 import _lib
from vjf_b200.model import VJF

T = int(os.environ.get("PT", 8))
torch.manual_seed(0)
lib = _lib.load()
for B, reps in ((4096, 2), (65536, 2)):
    m = VJF.make_model(200, 3, 0, 50, [64], "poisson", max_trials=B)
    y = torch.poisson(torch.full((T, B, 200), 0.5, device="cuda"))
    for _ in range(reps):
        mu, lv, ls = m.run(y)
    torch.cuda.synchronize()
    print("B", B, "kind", lib.vjf_last_launch_kind(), "loss", ls[-1].tolist(), "status", m.status())
