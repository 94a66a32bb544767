"""Quick device-side timing of the persistent run at a few shapes (development aid, not the benchmark)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vjf_b200.model import VJF

def t_run(D, d, u, R, H, lik, B, T, reps=3):
    m = VJF.make_model(D, d, u, R, H, lik, max_trials=B)
    y = torch.poisson(torch.full((T, B, D), 0.5, device="cuda")) if lik == "poisson" else torch.randn(T, B, D, device="cuda")
    uu = torch.randn(T, B, u, device="cuda") if u else None
    m.run(y, uu); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.run(y, uu); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = m.status()
    from vjf_b200 import _lib
    kind = _lib.load().vjf_last_launch_kind()
    print(f"[kind {kind}] D={D} d={d} R={R} H={H} {lik} B={B} T={T}: {best/T*1e3:.2f} us/step  {B*T/best*1e3:.3e} trial-steps/s status={st}", flush=True)

if __name__ == "__main__":
    t_run(200, 3, 0, 50, [64], "poisson", 4096, 64)
    t_run(200, 3, 0, 50, [64], "poisson", 16384, 32)
    t_run(200, 3, 0, 50, [64], "poisson", 65536, 16)
    t_run(50, 4, 0, 32, [32], "gaussian", 1024, 256)
    t_run(20, 2, 0, 100, [20], "gaussian", 1, 256)
    t_run(2000, 8, 0, 64, [128], "poisson", 8192, 8)
