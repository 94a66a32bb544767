"""C5 (BASELINE.json configs[4]): long-horizon latency path -- xdim 4, ydim 50 Gaussian, 32 RBFs, hidden [32], 1024 trials,
T up to 100 000 time steps with NO state reset.  Reports us/step and the status word per chunk for the fp32 and the
double-precision RLS (vjf_set_rls_precision)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import C5, bench_state, limit_cycle_gaussian
from vjf_b200.model import VJF, Gaussian

B = int(os.environ.get("PB", C5["trials_per_gpu"])); T = int(os.environ.get("PT", 100000)); chunk = int(os.environ.get("CHUNK", 10000))
dev = torch.device("cuda")
for bits in (64, 32):
    m = VJF.make_model(C5["ydim"], C5["xdim"], 0, C5["n_rbf"], C5["hidden"], "gaussian", max_trials=B, seed=3, rls_precision=bits)
    m.load_full_state(bench_state(C5))
    q = None
    tot = 0.0
    for c0 in range(0, T, chunk):
        y = limit_cycle_gaussian(c0, min(chunk, T - c0), B, C5["ydim"], dev, seed=7)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mu, lv, ls = m.run(y, None, q)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        tot += dt
        q = Gaussian(mu[-1].clone(), lv[-1].clone())
        st = m.status()
        print(f"rls{bits} steps {c0 + y.shape[0]:7d}  {dt / y.shape[0] * 1e6:6.1f} us/step  status {st}  loss {ls[:, 0].mean().item():9.4f} "
              f"tr_logvar {m.transition.logvar.item():7.3f} lik_logvar {m.likelihood.logvar.item():7.3f} Pmax {m.w_precision.abs().max().item():.3e} "
              f"|W|max {m.w_mean.abs().max().item():.3f} kind {m._lib.vjf_last_launch_kind()}", flush=True)
        if st & 16 and bits == 32:
            print("  fp32 RLS: Cholesky failed -> RLS state frozen from here on"); break
    print(f"rls{bits}: {tot / (c0 + y.shape[0]) * 1e6:.1f} us/step overall")
