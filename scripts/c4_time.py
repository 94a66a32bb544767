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
def step():
    m._flat.copy_(st); m.run(y)
for mode in [int(v) for v in os.environ.get("MODES", "0,1").split(",")]:
    _lib.check(m._lib.vjf_set_tile_mode(mode))
    ms = time_runs(step, int(os.environ.get("REPS", 3)))
    tps = B * T / ms * 1e3
    print(f"C4 B={B} mode={mode}: {ms / T * 1e3:.1f} us/step, {tps:.3e} trial-steps/s, HBM {ALGO_BYTES_PER_TRIAL_STEP(C4) * tps / 1e9:.1f} GB/s algorithmic, "
          f"status {m.status()}, kind {m._lib.vjf_last_launch_kind()}", flush=True)
