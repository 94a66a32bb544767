"""Per-phase time breakdown of the persistent kernel (CTA 0 globaltimer stamps)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vjf_b200 import _lib
from vjf_b200.model import VJF
B = int(os.environ.get("PB", 4096)); T = int(os.environ.get("PT", 64))
D, d, R, H = int(os.environ.get("PD", 200)), int(os.environ.get("Pd", 3)), int(os.environ.get("PR", 50)), [int(os.environ.get("PH", 64))]
lib = _lib.load()
m = VJF.make_model(D, d, 0, R, H, os.environ.get("PLIK", "poisson"), max_trials=B)
y = torch.poisson(torch.full((T, B, D), 0.5, device="cuda"))
warm = VJF.make_model(D, d, 0, R, H, os.environ.get("PLIK", "poisson"), max_trials=B)
for _ in range(20):
    warm.run(y)   # brings the clocks up; the measured model below starts from fresh RLS state
m.run(y[:8])
torch.cuda.synchronize()
dbg = torch.zeros(T, 64, dtype=torch.int64, device="cuda")
lib.vjf_debug_set_stamps.argtypes = [C.c_void_p]; lib.vjf_debug_set_stamps.restype = None
lib.vjf_debug_set_stamps(C.c_void_p(dbg.data_ptr()))
lib.vjf_debug_set_cta.argtypes = [C.c_int]; lib.vjf_debug_set_cta.restype = None
lib.vjf_debug_set_cta(int(os.environ.get('PCTA', 1)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.run(y); e1.record(); torch.cuda.synchronize()
lib.vjf_debug_set_stamps(None)
s = dbg.cpu().double()
ov = bool((s[5:, 22] > s[5:, 4]).all())  # front stamps of step t+1 land after barrier 2 of step t => overlapped schedule
names = ["back(t) | phaseA", "barrier1", "B1(reduce+sgd)", "barrier2", "B2(rls) || front(t+1)", "barrier3"]
dif = (s[:, 1:7] - s[:, 0:6])[4:-1]
print(f"B={B} D={D} R={R} H={H}: event time per step {e0.elapsed_time(e1)/T*1e3:.1f} us; step (stamps) {(s[-1,6]-s[4,0]).item()/(T-4)/1e3:.1f} us; overlapped={ov} status={m.status()}")
for i, n in enumerate(names):
    print(f"  {n:24s} mean {dif[:, i].mean().item()/1e3:7.2f} us   min {dif[:, i].min().item()/1e3:7.2f}  max {dif[:, i].max().item()/1e3:7.2f}")
front = [(22, 8, "S0 tile loads + cp.async wait"), (8, 9, "S1 xs"), (9, 11, "S2 phi"), (11, 12, "S4 mlp fwd"), (12, 13, "heads+xt"),
         (13, 18, "S5/S6 decoder+grads"), (18, 19, "S9 gram+b")]
back = [(10, 15, "S3 quadform+pm+plv"), (15, 16, "S7 dyn/entropy"), (16, 17, "S8 heads bwd+gpre"), (17, 20, "wgrad+colsum+scalars")]
print(" front half (first trial CTA):")
for i, j, n in front:
    print(f"  {n:30s} {(s[5:-1, j] - s[5:-1, i]).mean().item()/1e3:7.2f} us")
print(" back half:")
for i, j, n in back:
    print(f"  {n:30s} {(s[5:-1, j] - s[5:-1, i]).mean().item()/1e3:7.2f} us")
b2 = [(4, 30, "P, W, PW preload + stats wait"), (30, 25, "A, b -> P', g"), (25, 26, "rows init+publish"), (26, 27, "LDL sweep"), (27, 28, "scale+commit"),
      (28, 29, "W'=Uz"), (29, 5, "residual/var + losses")]
# timeline of one step relative to CTA 0's step start (mean over steps)
tl = [("cta0 arrive bar1", 1, 0), ("cta0 bar1 released", 2, 0), ("cta0 B2 start", 4, 0), ("cta0 B2 end", 5, 0), ("cta0 bar3 released", 6, 0),
      ("trial back start (after cp.async wait)", 10, 0), ("trial back end", 20, 0), ("trial B1 end", 23, 0), ("trial barrier released", 21, 0), ("trial front prologue issued", 7, 0), ("trial front(t+1) tile loads issued", 22, 1), ("trial front(t+1) staged", 8, 1), ("trial front(t+1) end", 19, 1)]
print(f" timeline (us after step start; trial CTA {os.environ.get('PCTA', 1)}):")
for n, i, dt in tl:
    a = s[5 + dt:-2 + dt if -2 + dt else None, i] - s[5:-2, 0]
    print(f"  {n:40s} {a.mean().item()/1e3:7.2f}")
fine = [(10, 49, "S3 quadform (mma)"), (49, 50, "S3 p_mean dots"), (50, 51, "S3 sync"), (51, 15, "S3 p_logvar + sync"),
        (17, 52, "wgrad: bias colsum"), (52, 53, "wgrad: fences + sync"), (53, 54, "wgrad: 12 tcgen05.mma + commit + wait"), (54, 43, "wgrad: TMEM -> slot epilogue"),
        (13, 41, "decoder rows (warp 0)"), (41, 42, "decoder sync"), (42, 18, "decoder grads -> slot"),
        (17, 43, "colsum + umma wgrad"), (43, 44, "scalar warp sums + sync"), (44, 45, "scalar store + sync"), (45, 20, "y prefetch issue")]
print(" fine stamps (trial CTA):")
for i, j, n in fine:
    print(f"  {n:30s} {(s[5:-1, j] - s[5:-1, i]).mean().item()/1e3:7.2f} us")
print(f"  prologue: proxy fences {(s[0,48]-s[0,47]).item()/1e3:.2f} us (last step)")
print(" phase B2 stages (CTA 0):")
for i, j, n in b2:
    print(f"  {n:30s} {(s[5:-1, j] - s[5:-1, i]).mean().item()/1e3:7.2f} us")

dc = (s[-1, 40] - s[4, 40]).item(); dg = (s[-1, 0] - s[4, 0]).item()
print(f" effective SM clock of CTA 0 during the run: {dc / dg * 1e3:.0f} MHz (clock64 ticks / globaltimer ns)")

import numpy as np
tk = (C.c_longlong * 160)()
lib.vjf_debug_read_sweep.argtypes = [C.c_void_p]; lib.vjf_debug_read_sweep.restype = C.c_int
lib.vjf_debug_read_sweep(tk)
tk = np.array(tk[:R], dtype=np.int64)
print(" sweep: cycles per column (clock64, last step):", np.diff(tk).tolist())
