"""Parity of the throughput tile pipeline (vjf_b200/csrc/tile_kernels.cuh: every contraction on tcgen05, observations by TMA
tensor copies) against the fp64 oracle and against the general persistent kernel, in every variant of its plan: general
observations (full lo image, 32-trial tiles), exact spike counts (32- and 64-trial tiles), Gaussian likelihood, control
input, ragged last tile, more tiles than CTAs, warm-up / frozen decoder / sgd=False / update=False flags, in-kernel Philox,
and the non-finite ELBO-term path (vjf/model.py:138-145).  Needs a B200: pytest -m gpu."""
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


@pytest.fixture(autouse=True)
def _auto_mode():
    from vjf_b200 import _lib
    yield
    _lib.check(_lib.load().vjf_set_tile_mode(0))


def _data(rng, lik, T, B, D, d, udim):
    if lik == "poisson":
        t = np.arange(T)[:, None, None] * 0.05
        ph = rng.uniform(0, 2 * np.pi, (1, B, d))
        x = np.sin(t * (1 + np.arange(d)) + ph)
        Cm = rng.normal(size=(d, D)) / np.sqrt(d)
        y = rng.poisson(np.exp(np.clip(x @ Cm - 1.0, None, 3.0))).astype(np.float32)
    else:
        y = rng.normal(size=(T, B, D)).astype(np.float32)
    u = rng.normal(size=(T, B, udim)).astype(np.float32) if udim else None
    return y, u


def _run(cuda_mod, mode, lik, B, D, d, R, H, T, udim=0, seed=3, lr=1e-3, y_override=None, flags=None, want_kind=1):
    """One run of the CUDA model under tile mode `mode` and of the fp64 oracle on the same inputs."""
    from vjf_b200 import _lib
    from vjf_b200.model import VJF
    lib = _lib.load()
    _lib.check(lib.vjf_set_tile_mode(mode))
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    m = VJF.make_model(D, d, udim, R, H, lik, lr=lr, max_trials=B, seed=99)
    o = O.OracleVJF(D, d, udim, R, H, lik, lr=lr, dtype=np.float64)
    o.set_state(cuda_mod.state_np(m))
    y, u = _data(rng, lik, T, B, D, d, udim)
    if y_override is not None:
        y = y_override(y)
    eps = rng.normal(size=(T, 2, B, d)).astype(np.float32)
    kw = dict(sgd=True, update=True, warm_up=False)
    kw.update(flags or {})
    mu, lv, losses = m.run(torch.as_tensor(y), None if u is None else torch.as_tensor(u), None, eps=torch.as_tensor(eps), **kw)
    torch.cuda.synchronize()
    assert lib.vjf_last_launch_kind() == want_kind, "the launch did not take the expected kernel"
    omu, olv, ol = o.run(y.astype(np.float64), None if u is None else u.astype(np.float64), eps=eps.astype(np.float64), **kw)
    return m, o, (mu.cpu().numpy(), lv.cpu().numpy(), losses.cpu().numpy()), (omu, olv, ol)


def _check(cuda_mod, m, o, got, want, rls_tol=2e-3):
    assert_close(got[0], want[0], 2e-4, 2e-5, "mu")
    assert_close(got[1], want[1], 2e-4, 2e-5, "logvar")
    assert_close(got[2], want[2], 2e-4, 2e-3, "losses")
    compare_state(cuda_mod.state_np(m), o.get_state(), rtol=rls_tol, atol=rls_tol / 10)


# (mode, B): 2 = general observations; 0 = automatic (exact counts, 32-trial tiles at small B); 3 = exact with 64-trial tiles
@pytest.mark.parametrize("mode,B", [(2, 300), (0, 300), (3, 300), (3, 97), (2, 33)])
def test_poisson_c2_shapes_match_oracle(cuda_mod, mode, B):
    m, o, got, want = _run(cuda_mod, mode, "poisson", B, 200, 3, 50, [64], 6)
    _check(cuda_mod, m, o, got, want)
    assert m.status() == 0


def test_more_tiles_than_ctas_64_trial_tiles(cuda_mod):
    """20 000 trials: 313 tiles of 64 trials over 147 trial CTAs -- accumulators in tensor memory across the tiles of a CTA."""
    m, o, got, want = _run(cuda_mod, 0, "poisson", 20000, 200, 3, 50, [64], 3)
    assert_close(got[0], want[0], 2e-4, 2e-5, "mu")
    assert_close(got[1], want[1], 2e-4, 2e-5, "logvar")
    assert_close(got[2], want[2], 2e-4, 2e-3, "losses")
    compare_state(cuda_mod.state_np(m), o.get_state(), rtol=3e-3, atol=3e-4, skip=("w_mean", "w_chol", "w_precision", "transition.logvar"))
    assert m.status() == 0


def test_non_integer_poisson_observations_take_the_general_plan(cuda_mod):
    """The reference's own smoke test feeds N(0,1) 'counts' to the Poisson model (test/test_model.py:32-44): the exactness
    scan must send such data through the full (hi, lo) image."""
    m, o, got, want = _run(cuda_mod, 0, "poisson", 200, 200, 3, 50, [64], 4, y_override=lambda y: (y + 0.123456789).astype(np.float32))
    _check(cuda_mod, m, o, got, want)


@pytest.mark.parametrize("udim", [0, 2])
def test_gaussian_likelihood_and_control_input(cuda_mod, udim):
    m, o, got, want = _run(cuda_mod, 0, "gaussian", 150, 48, 4, 32, [32], 6, udim=udim)
    _check(cuda_mod, m, o, got, want)
    assert m.status() == 0


def test_wide_hidden_layer_and_xdim8(cuda_mod):
    m, o, got, want = _run(cuda_mod, 0, "poisson", 130, 96, 8, 40, [128], 4)
    _check(cuda_mod, m, o, got, want)


@pytest.mark.parametrize("flags", [dict(warm_up=True), dict(sgd=False), dict(update=False)])
def test_step_flags(cuda_mod, flags):
    m, o, got, want = _run(cuda_mod, 0, "poisson", 120, 200, 3, 50, [64], 5, flags=flags)
    _check(cuda_mod, m, o, got, want)


def test_tile_pipeline_equals_persistent_kernel(cuda_mod):
    """Same inputs through the tile pipeline and through the general persistent kernel: identical up to summation order."""
    a, _, ga, _ = _run(cuda_mod, 0, "poisson", 500, 200, 3, 50, [64], 8)
    b, _, gb, _ = _run(cuda_mod, 1, "poisson", 500, 200, 3, 50, [64], 8, want_kind=0)
    assert_close(ga[0], gb[0], 5e-5, 5e-6, "mu")
    assert_close(ga[1], gb[1], 5e-5, 5e-6, "logvar")
    assert_close(ga[2], gb[2], 5e-5, 5e-4, "losses")
    # (the RLS solution amplifies summation-order differences of the statistics: compared more loosely)
    rls = ("w_mean", "w_chol", "w_precision", "w_pchol", "transition.logvar")
    sa, sb_ = cuda_mod.state_np(a), cuda_mod.state_np(b)
    compare_state(sa, sb_, rtol=1e-4, atol=1e-5, skip=rls)
    compare_state({k: sa[k] for k in rls}, {k: sb_[k] for k in rls}, rtol=5e-3, atol=1e-3)


def test_philox_in_kernel_equals_tape(cuda_mod):
    from vjf_b200 import _lib
    from vjf_b200.model import VJF
    lib = _lib.load()
    B, D, d, T = 260, 200, 3, 4
    rng = np.random.default_rng(1)
    y, _ = _data(rng, "poisson", T, B, D, d, 0)
    torch.manual_seed(1)
    a = VJF.make_model(D, d, 0, 50, [64], "poisson", lr=1e-3, max_trials=B, seed=77)
    b = VJF.make_model(D, d, 0, 50, [64], "poisson", lr=1e-3, max_trials=B, seed=77)
    b.load_full_state(a.full_state())
    e = torch.empty(T, 2, B, d, device="cuda")
    for t in range(T):
        _lib.check(lib.vjf_philox_normal(77, t, 0, B, d, C.c_void_p(e[t].data_ptr()), None))
    mu_a, lv_a, ls_a = a.run(torch.as_tensor(y))
    assert lib.vjf_last_launch_kind() == 1
    mu_b, lv_b, ls_b = b.run(torch.as_tensor(y), None, None, eps=e)
    assert torch.equal(mu_a, mu_b) and torch.equal(lv_a, lv_b) and torch.equal(ls_a, ls_b)
    assert torch.equal(a._flat, b._flat)


def test_nonfinite_term_is_zeroed_without_gradient(cuda_mod):
    """vjf/model.py:138-145: a non-finite ELBO term becomes the constant 0 and carries no gradient -- the tile pipeline redoes
    the tiles of the step without it.  Scenario of the oracle test: fp32 overflow of the trace term exp(p_logvar + l_t - gamma)
    with l_t ~ 120 (exp(l_t / 2) is still finite, so the other two terms and their gradients stay finite)."""
    from vjf_b200 import _lib
    from vjf_b200.model import VJF, Gaussian
    lib = _lib.load()
    B, D, d = 100, 200, 3
    rng = np.random.default_rng(2)
    y, _ = _data(rng, "poisson", 1, B, D, d, 0)
    eps = np.zeros((1, 2, B, d), np.float32)
    torch.manual_seed(2)
    m = VJF.make_model(D, d, 0, 50, [64], "poisson", lr=1e-2, max_trials=B)
    with torch.no_grad():
        m.recognition.logvar.bias.fill_(120.0)
    o = O.OracleVJF(D, d, 0, 50, [64], "poisson", lr=1e-2, dtype=np.float32)
    o.set_state(cuda_mod.state_np(m))
    z = np.zeros((B, d), np.float32)
    mu, lv, losses = m.run(torch.as_tensor(y), None, Gaussian(torch.as_tensor(z), torch.as_tensor(z)), eps=torch.as_tensor(eps), update=False)
    torch.cuda.synchronize()
    assert lib.vjf_last_launch_kind() == 1
    omu, olv, ol = o.run(y, None, eps=eps, update=False, q0=O.Gaussian(z, z))
    assert m.status() & _lib.ST_DYN_NONFINITE
    got = losses.cpu().numpy()
    assert got[0, 2] == 0.0 and ol[0, 2] == 0.0
    assert_close(got[0, [0, 1, 3]], ol[0, [0, 1, 3]], 1e-4, 1e-4, "loss terms")
    assert_close(mu.cpu().numpy(), omu, 1e-4, 1e-5, "mu")
    compare_state(cuda_mod.state_np(m), o.get_state(), rtol=2e-4, atol=2e-6)
