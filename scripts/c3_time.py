"""C3 (n_rbf = 1024, 16 384 trials): time per step of the large-n_rbf launch sequence."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import C3, bench_state, rotation_gaussian, c3_tensor_flops_per_step, time_runs
from vjf_b200.model import VJF
dev = torch.device("cuda")
B = int(os.environ.get("PB", C3["trials_per_gpu"])); T = int(os.environ.get("PT", 4))
m = VJF.make_model(C3["ydim"], C3["xdim"], 0, C3["n_rbf"], C3["hidden"], "gaussian", max_trials=B, seed=99)
m.load_full_state(bench_state(C3))
y = rotation_gaussian(T, B, C3["ydim"], C3["xdim"], dev, 17)
st = m._flat.clone()
def step():
    m._flat.copy_(st); m.run(y)
ms = time_runs(step, int(os.environ.get("REPS", 3)))
fl = c3_tensor_flops_per_step(B, C3["n_rbf"])
print(f"C3 B={B}: {ms / T * 1e3:.1f} us/step, {B * T / ms * 1e3:.3e} trial-steps/s, tensor {fl / (ms / T * 1e-3) / 1e12:.1f} TFLOP/s issued, status {m.status()}, kind {m._lib.vjf_last_launch_kind()}")
if os.environ.get("STAMPS"):
    import ctypes as C, numpy as np
    ptr = m._lib.vjf_bigr_buffer(m._h, 4)
    from vjf_b200 import _lib
    buf = torch.as_tensor(_lib.DevBuf(ptr, 1024), device="cuda").cpu().numpy().view(np.uint64)[8:8 + 64].reshape(16, 4).astype(np.int64)
    for kb in range(16):
        a = buf[kb]
        nxt = buf[kb + 1][0] if kb + 1 < 16 else a[3]
        print(f"panel {kb:2d}: chol {(a[1]-a[0])/1e3:6.1f} us  rows {(a[2]-a[1])/1e3:6.1f}  barrier {(a[3]-a[2])/1e3:6.1f}  tiles+barrier {(nxt-a[3])/1e3:6.1f}")
