"""Stage timeline of the tile pipeline (first tile of one trial CTA per step; debug library with globaltimer stamps).
   VJF_B200_LIB=vjf_b200/lib/libvjf_b200_dbg.so python scripts/tile_stamps.py   (env: PB trials, PT steps, PCTA cta)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vjf_b200 import _lib
from vjf_b200.model import VJF
B = int(os.environ.get("PB", 4096)); T = int(os.environ.get("PT", 64))
D, d, R, H = int(os.environ.get("PD", 200)), int(os.environ.get("Pd", 3)), int(os.environ.get("PR", 50)), [int(os.environ.get("PH", 64))]
lik = os.environ.get("PLIK", "poisson")
lib = _lib.load()
m = VJF.make_model(D, d, 0, R, H, lik, max_trials=B)
y = torch.poisson(torch.full((T, B, D), 0.5, device="cuda")) if lik == "poisson" else torch.randn(T, B, D, device="cuda")
warm = VJF.make_model(D, d, 0, R, H, lik, max_trials=B)
for _ in range(10):
    warm.run(y)
torch.cuda.synchronize()
dbg = torch.zeros(T, 64, dtype=torch.int64, device="cuda")
lib.vjf_debug_set_stamps.argtypes = [C.c_void_p]; lib.vjf_debug_set_stamps.restype = None
lib.vjf_debug_set_cta.argtypes = [C.c_int]; lib.vjf_debug_set_cta.restype = None
lib.vjf_debug_set_stamps(C.c_void_p(dbg.data_ptr()))
lib.vjf_debug_set_cta(int(os.environ.get("PCTA", 1)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.run(y); e1.record(); torch.cuda.synchronize()
lib.vjf_debug_set_stamps(None)
s = dbg.cpu().double()
print(f"B={B} D={D} R={R} H={H} {lik}: {e0.elapsed_time(e1) / T * 1e3:.1f} us per step, kind {lib.vjf_last_launch_kind()} status {m.status()}")
lo, hi = 4, T - 2
names = {0: "tile start", 1: "extras+eps", 2: "xs", 3: "wait DW(prev)", 4: "wait y (TMA)", 5: "X cols + lo image (CX)", 6: "phi written (CPHI)", 7: "wait D1 (FWD MMA)",
         8: "E1 tanh", 9: "heads/xt", 10: "permute in", 11: "LK decoder+likelihood", 12: "wait FL (QUAD MMA)", 13: "phi permute + dx + FL epilogue",
         14: "wait ctrl3 (RLS tail)", 15: "S7 dyn/entropy", 16: "wait GRAM", 17: "GS g_pre + head grads (CG)"}
print(" compute warps, first tile (us after run_tiles entry; delta):")
base = s[lo:hi, 48]
prev = None
for i in range(18):
    v = (s[lo:hi, i] - base).mean().item() / 1e3
    print(f"  {i:2d} {names[i]:36s} {v:8.2f}  {'' if prev is None else f'+{v - prev:6.2f}'}")
    prev = v
print(" LK detail (warp 0): " + " ".join(f"{(s[lo:hi, i] - s[lo:hi, 10]).mean().item() / 1e3:.2f}" for i in range(20, 32)))
cn = {32: "step start (y TMA + ring prefill issued)", 33: "CX seen", 34: "FWD issued", 35: "CPHI seen", 36: "UK loaded", 37: "QUAD issued", 38: "CPHIT seen", 39: "GRAM issued",
      40: "CG seen", 41: "DW issued"}
print(" control warp:")
for i in range(32, 42):
    v = (s[lo:hi, i] - base).mean().item() / 1e3
    print(f"  {i:2d} {cn[i]:40s} {v:8.2f}")
sn = {49: "tiles done (before flush)", 50: "flush done", 51: "barrier 1 released", 52: "B1 done", 53: "trial barrier released"}
print(" step level (us after run_tiles entry of the SAME step's tiles):")
sn.update({54: "flush: tensor-memory accumulators stored", 55: "flush: register accumulators in scratch"})
for i in (49, 54, 55, 50):
    print(f"  {sn[i]:40s} {(s[lo:hi, i] - base).mean().item() / 1e3:8.2f}")
# stamps 51..53 of step t follow the tiles of step t (entered during step t-1)
for i in (51, 52, 53):
    print(f"  {sn[i]:40s} {(s[lo:hi, i] - base).mean().item() / 1e3:8.2f}")
print(f"  next run_tiles entry                     {(s[lo + 1:hi + 1, 48] - base).mean().item() / 1e3:8.2f}")
rn = {30: "RLS: statistics arrived (independent part P, P W done before)", 25: "RLS: P' / g built", 26: "RLS: sweep starts", 27: "RLS: sweep done",
      28: "RLS: columns scaled, w_pchol / w_chol / images committed", 29: "RLS: W' done, ctrl5 published", 24: "B2: losses read out (deferred: after the factorisation)"}
print(" RLS CTA (us after barrier 1 of the same step):")
b1 = s[lo:hi, 51]
for i in (30, 25, 26, 27, 28, 29, 24):
    print(f"  {i:2d} {rn[i]:60s} {(s[lo:hi, i] - b1).mean().item() / 1e3:8.2f}")
print(f"  trial CTAs: B1 done {(s[lo:hi, 52] - b1).mean().item() / 1e3:8.2f}  trial barrier released {(s[lo:hi, 53] - b1).mean().item() / 1e3:8.2f}")
print(f"  next barrier 1 released {(s[lo + 1:hi + 1, 51] - b1).mean().item() / 1e3:8.2f}")
