"""C4 (ydim 2000 Poisson, xdim 8): time per step of the wide-observation launch sequence (csrc/wide.cu) vs the general kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import C4, bench_state, synthetic_counts_gpu, time_runs, ALGO_BYTES_PER_TRIAL_STEP
from vjf_b200 import _lib
from vjf_b200.model import VJF
dev = torch.device("cuda")
B = int(os.environ.get("PB", C4["global_trials"] // 8)); T = int(os.environ.get("PT", C4["T"]))
m = VJF.make_model(C4["ydim"], C4["xdim"], 0, C4["n_rbf"], C4["hidden"], C4["likelihood"], max_trials=B, seed=99)
m.load_full_state(bench_state(C4))
y = synthetic_counts_gpu(T, B, C4["ydim"], C4["xdim"], dev, 31)
st = m._flat.clone()
UPD = os.environ.get("UPD", "1") == "1"
def step():
    m._flat.copy_(st); m.run(y, update=UPD)
for mode in [int(v) for v in os.environ.get("MODES", "0,1").split(",")]:
    _lib.check(m._lib.vjf_set_tile_mode(mode))
    ms = time_runs(step, int(os.environ.get("REPS", 3)))
    tps = B * T / ms * 1e3
    print(f"C4 B={B} mode={mode}: {ms / T * 1e3:.1f} us/step, {tps:.3e} trial-steps/s, HBM {ALGO_BYTES_PER_TRIAL_STEP(C4) * tps / 1e9:.1f} GB/s algorithmic, "
          f"status {m.status()}, kind {m._lib.vjf_last_launch_kind()}", flush=True)
if os.environ.get("VJF_WIDE_STAMPS"):
    import ctypes as C, numpy as np
    ptr = m._lib.vjf_wide_stamps(m._h)
    buf = torch.as_tensor(_lib.DevBuf(ptr, 128), device="cuda").cpu().numpy().view(np.uint64).astype(np.int64).reshape(8, 8)
    t_last = (m._step_index - 1) & 7
    base = buf[(t_last - 3) & 7][2]
    for k in (3, 2, 1, 0):
        r = buf[(t_last - k) & 7]
        print(f"step -{k}: mid start {(r[2]-base)/1e3:7.1f}  wait {(r[3]-base)/1e3:7.1f} .. {(r[4]-base)/1e3:7.1f}  mid end(CTA0) {(r[5]-base)/1e3:7.1f} (last CTA {(r[7]-base)/1e3:7.1f})  sgd end {(r[6]-base)/1e3:7.1f}"
              f"  rls {(r[0]-base)/1e3:7.1f} .. {(r[1]-base)/1e3:7.1f}")
