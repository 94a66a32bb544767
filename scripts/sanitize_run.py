"""Short runs of every kernel family for compute-sanitizer (scripts/sanitize.sh): persistent kernel, tile pipeline, split phases,
large-n_rbf launch sequence, the rows next to the hot path (RLS initialisation over N >= 32 x SMs samples, weight-space Kalman,
forecast) and the batched Kalman operator."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from vjf_b200 import _lib
from vjf_b200.model import VJF

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda")
torch.manual_seed(0)
rng = np.random.default_rng(0)
lib = _lib.load()


def run(name, fn):
    if which in ("all", name):
        fn(); torch.cuda.synchronize(); print("SANITIZE_RAN", name, flush=True)


def persistent():
    lib.vjf_set_tile_mode(1)
    m = VJF.make_model(40, 3, 1, 20, [16, 8], "gaussian", max_trials=300)
    m.run(torch.randn(3, 300, 40), torch.randn(3, 300, 1)); assert m.status() == 0
    m = VJF.make_model(200, 3, 0, 50, [64], "poisson", max_trials=300)   # overlapped schedule with TMA + tcgen05 weight gradient
    m.run(torch.poisson(torch.full((3, 300, 200), 0.7))); assert m.status() == 0
    lib.vjf_set_tile_mode(0)


def tile():
    for mode, B in ((2, 300), (3, 300), (0, 5000)):
        lib.vjf_set_tile_mode(mode)
        m = VJF.make_model(200, 3, 0, 50, [64], "poisson", max_trials=B)
        m.run(torch.poisson(torch.full((2, B, 200), 0.7))); assert m.status() == 0 and lib.vjf_last_launch_kind() == 1
    lib.vjf_set_tile_mode(0)


def split():
    m = VJF.make_model(30, 2, 0, 12, [8], "gaussian", max_trials=100)
    y = torch.randn(100, 30, device=dev); mean = torch.empty(100, 2, device=dev); lv = torch.empty_like(mean); loss = torch.empty(4, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    fl = _lib.FLAG_SGD | _lib.FLAG_UPDATE | _lib.FLAG_PRIOR_Q0
    _lib.check(lib.vjf_step_phase_a(m._h, 100, 100, p(y), 0, None, None, None, None, 1, 0, 0, fl, p(mean), p(lv), None))
    _lib.check(lib.vjf_step_phase_b(m._h, 100, fl, 1e-3, p(loss), None))


def bigr():
    m = VJF.make_model(24, 3, 0, 200, [16], "gaussian", max_trials=150)
    m.run(torch.randn(2, 150, 24)); assert m.status() == 0 and lib.vjf_last_launch_kind() == 2


def rls64():
    m = VJF.make_model(24, 3, 0, 40, [16], "gaussian", max_trials=150, rls_precision=64)
    m.run(torch.randn(3, 150, 24)); assert m.status() == 0


def aux():
    R, d = 50, 3
    m = VJF.make_model(12, d, 0, R, [8], "gaussian", max_trials=64)
    N = 32 * torch.cuda.get_device_properties(0).multi_processor_count + 77   # every slot in use (out-of-bounds store there)
    xs = torch.randn(N, d, device=dev); xt = xs + 0.1 * torch.randn(N, d, device=dev)
    m.initialize_transition(xt, xs, centroid=torch.rand(R, d, device=dev) * 4 - 2)
    m.transition.velocity.kalman(xs, 0.1 * torch.randn(N, d, device=dev), 0.3, diffusion=0.01)
    m.forecast(torch.randn(5, d), None, 4, noise=True)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore"); m.status()


for name, fn in (("persistent", persistent), ("tile", tile), ("split", split), ("bigr", bigr), ("rls64", rls64), ("aux", aux)):
    run(name, fn)
