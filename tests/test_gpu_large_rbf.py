"""Large n_rbf (BASELINE.json configs[2], "C3": n_rbf = 1024): the launch sequence of csrc/bigr.cu -- phi w_chol and phi^T phi as
tcgen05 GEMMs over all trials, blocked multi-CTA Cholesky -- against the fp64 oracle.  Needs a B200: pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import vjf_oracle as O
from tests.helpers import assert_close, compare_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda_mod():
    from tests import gpu_helpers
    assert torch.cuda.is_available(), "GPU tests selected without a CUDA device"
    return gpu_helpers


def _case(cuda_mod, lik, B, D, d, u, R, H, T, seed=5, lr=1e-3, flags=None):
    from vjf_b200.model import VJF
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    m = VJF.make_model(D, d, u, R, H, lik, lr=lr, max_trials=B)
    o = O.OracleVJF(D, d, u, R, H, lik, lr=lr, dtype=np.float64)
    o.set_state(cuda_mod.state_np(m))
    y = rng.poisson(0.7, (T, B, D)).astype(np.float32) if lik == "poisson" else rng.normal(size=(T, B, D)).astype(np.float32)
    uu = rng.normal(size=(T, B, u)).astype(np.float32) if u else None
    eps = rng.normal(size=(T, 2, B, d)).astype(np.float32)
    kw = flags or {}
    mu, lv, losses = m.run(torch.as_tensor(y), None if uu is None else torch.as_tensor(uu), None, eps=torch.as_tensor(eps), **kw)
    assert m._lib.vjf_last_launch_kind() == 2
    omu, olv, olosses = o.run(y.astype(np.float64), uu, eps=eps.astype(np.float64), **kw)
    assert_close(mu.cpu().numpy(), omu, 2e-4, 2e-5, "mu")
    assert_close(lv.cpu().numpy(), olv, 2e-4, 2e-5, "logvar")
    assert_close(losses.cpu().numpy(), olosses, 2e-4, 2e-3, "losses")
    compare_state(cuda_mod.state_np(m), o.get_state(), rtol=1e-3, atol=1e-4)
    assert m.status() == 0
    return m, o


@pytest.mark.parametrize("lik,B,D,d,u,R,H,T", [("gaussian", 300, 40, 5, 0, 256, [32], 4), ("poisson", 130, 64, 3, 2, 164, [16, 8], 3),
                                               ("gaussian", 77, 20, 10, 0, 420, [24], 3)])
def test_large_rbf_steps_match_oracle(cuda_mod, lik, B, D, d, u, R, H, T):
    """Seeded steps above the single-CTA limit (n_rbf > 160): every output and the whole state against the fp64 oracle; odd sizes
    (n_rbf not a multiple of the 64-column panel / 256-column GEMM tile, trials not a multiple of the 128-row tile)."""
    _case(cuda_mod, lik, B, D, d, u, R, H, T)


@pytest.mark.parametrize("flags", [dict(warm_up=True), dict(sgd=False), dict(update=False)])
def test_large_rbf_step_flags(cuda_mod, flags):
    _case(cuda_mod, "gaussian", 200, 30, 4, 0, 192, [16], 3, flags=flags)


def test_c3_shapes_2048_trials_match_oracle(cuda_mod):
    """BASELINE config 3 dimensions (xdim 10, ydim 500 Gaussian, 1024 RBFs, hidden [128]) at 2048 trials, 3 steps."""
    _case(cuda_mod, "gaussian", 2048, 500, 10, 0, 1024, [128], 3)


def test_large_rbf_philox_equals_tape(cuda_mod):
    from vjf_b200 import _lib
    from vjf_b200.model import VJF
    torch.manual_seed(3)
    B, D, d, R, T = 150, 24, 3, 200, 3
    a = VJF.make_model(D, d, 0, R, [16], "gaussian", max_trials=B, seed=77)
    b = VJF.make_model(D, d, 0, R, [16], "gaussian", max_trials=B, seed=77)
    b.load_full_state(a.full_state())
    y = torch.randn(T, B, D)
    e = torch.empty(T, 2, B, d, device="cuda")
    for t in range(T):
        _lib.check(a._lib.vjf_philox_normal(77, t, 0, B, d, C.c_void_p(e[t].data_ptr()), None))
    mu_a, lv_a, l_a = a.run(y)
    mu_b, lv_b, l_b = b.run(y, eps=e)
    assert torch.equal(mu_a, mu_b) and torch.equal(lv_a, lv_b) and torch.equal(a._flat, b._flat)
