import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import C5, C2, bench_state, limit_cycle_gaussian, lorenz_poisson
from oracle.vjf_oracle import OracleVJF
from vjf_b200.model import VJF, Gaussian
dev = torch.device("cuda")
# (1) C5 shapes, 2000 steps vs fp64 oracle
B, T = int(os.environ.get("PB", 64)), int(os.environ.get("PT", 2000))
st = {k: v.numpy() for k, v in bench_state(C5).items()}
y = limit_cycle_gaussian(0, T, B, C5["ydim"], dev, seed=7)
eps = torch.randn(T, 2, B, C5["xdim"], generator=torch.Generator().manual_seed(1))
for bits in (64, 32):
    m = VJF.make_model(C5["ydim"], C5["xdim"], 0, C5["n_rbf"], C5["hidden"], "gaussian", max_trials=B, seed=3, rls_precision=bits, lr=1e-3)
    m.load_full_state(bench_state(C5))
    mu, lv, ls = m.run(y, None, None, eps=eps)
    print("bits", bits, "status", m.status(), "loss last", ls[-5:, 0].tolist(), "Pmax %.3e" % m.w_precision.abs().max().item())
    if bits == 64:
        o = OracleVJF(C5["ydim"], C5["xdim"], 0, C5["n_rbf"], C5["hidden"], "gaussian", dtype=np.float64, lr=1e-3)
        s0 = o.get_state(); s0.update(st); o.set_state(s0)
        t0 = time.time()
        omu, olv, ols = o.run(y.cpu().numpy().astype(np.float64), eps=eps.numpy().astype(np.float64))
        print("oracle", time.time() - t0, "s status", o.status)
    for name, a, b in (("mu", mu.cpu().numpy(), omu), ("lv", lv.cpu().numpy(), olv), ("loss", ls.cpu().numpy(), ols)):
        err = np.abs(a - b)
        for t in (10, 100, 500, 1000, T - 1):
            print(f"  {name} t={t}: max err {err[t].max():.3e} (scale {np.abs(b[t]).max():.3e})")
    print("  w_mean err", np.abs(m.w_mean.cpu().numpy() - o.w_mean).max(), "scale", np.abs(o.w_mean).max(),
          "tr_logvar", m.transition.logvar.item(), o.tr_logvar, "lik", m.likelihood.logvar.item(), o.lik_logvar)
    print("  P err rel", (np.abs(m.w_precision.cpu().numpy() - o.w_precision).max() / np.abs(o.w_precision).max()))
# (2) C2 shapes: fp32 RLS breaks, fp64 survives
Bc = 4096
yc = lorenz_poisson(256, Bc, 200, seed=1000).to(dev)
for bits in (32, 64):
    m = VJF.make_model(200, 3, 0, 50, [64], "poisson", max_trials=Bc, seed=99, rls_precision=bits)
    m.load_full_state(bench_state(C2))
    for ep in range(24):
        t0 = time.perf_counter()
        mu, lv, ls = m.run(yc)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        stt = m.status()
        if stt or ep % 4 == 3:
            print("C2 bits", bits, "epoch", ep, "status", stt, "loss %.3f" % ls[:, 0].mean().item(), "Pmax %.3e" % m.w_precision.abs().max().item(),
                  "|W| %.3f" % m.w_mean.abs().max().item(), "%.1f us/step" % (dt / 256 * 1e6), "kind", m._lib.vjf_last_launch_kind(), flush=True)
        if stt: break
