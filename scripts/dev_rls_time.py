import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vjf_b200.model import VJF
dev = torch.device("cuda")
for R in (32, 50, 100):
    for bits in (32, 64):
        m = VJF.make_model(50, 3, 0, R, [32], "gaussian", max_trials=64, rls_precision=bits)
        xs = torch.randn(64, 3, device=dev); xt = xs + 0.1 * torch.randn(64, 3, device=dev)
        cen = torch.rand(R, 3, device=dev) * 4 - 2
        for _ in range(5): m.initialize_transition(xt, xs, centroid=cen)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        n = 200
        from vjf_b200 import _lib
        import ctypes as C
        for _ in range(n):
            _lib.check(m._lib.vjf_rls_initialize(m._h, 64, C.c_void_p(xs.data_ptr()), C.c_void_p(xt.data_ptr()), None, m._stream()))
        b.record(); torch.cuda.synchronize()
        print(f"R={R} bits={bits}: {a.elapsed_time(b) / n * 1e3:.1f} us per initialize (3 launches), status {m.status()}")
