"""Parity of the wide-observation path (csrc/wide.cu: ydim above the tile pipeline's limit; BASELINE configs[3], "C4":
ydim 2000 Poisson, xdim 8, 64 RBFs, hidden [128]) against the fp64 oracle and against the general persistent kernel:
the two contractions over the observation columns run as tcgen05 GEMMs over all trials of the step (the forward one reads the
observations K-major, the weight-gradient one reads the same rows MN-major through 32-byte-atom TMA boxes).  Needs a B200."""
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


def _make(lik, B, D, d, udim, R, H, lr=1e-3):
    from vjf_b200.model import VJF
    torch.manual_seed(5)
    return VJF.make_model(D, d, udim, R, H, lik, lr=lr, max_trials=B, seed=4321)


def _run(m, y, u, eps, mode):
    from vjf_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.vjf_set_tile_mode(mode))
    out = m.run(torch.as_tensor(y), None if u is None else torch.as_tensor(u), None, eps=None if eps is None else torch.as_tensor(eps))
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out], lib.vjf_last_launch_kind()


@pytest.mark.parametrize("lik,B,D,d,udim,R,H,T", [
    ("poisson", 300, 512, 3, 0, 20, [64], 5),     # ragged: sub-tiles of 32 + a tail, 2 trials of padding in a TMA box
    ("poisson", 1111, 1000, 8, 2, 64, [128], 4),  # control input, odd trial count, observation columns not a multiple of 128
    ("gaussian", 515, 640, 4, 1, 33, [96], 4),    # general observations: lo image of y in both GEMMs; H below the tile width
    ("poisson", 37, 2000, 8, 0, 64, [128], 3),    # fewer trials than one GEMM tile
])
def test_wide_path_matches_oracle_and_general_kernel(cuda_mod, lik, B, D, d, udim, R, H, T):
    rng = np.random.default_rng(3)
    y, u = _data(rng, lik, T, B, D, d, udim)
    eps = rng.normal(size=(T, 2, B, d)).astype(np.float32)
    m = _make(lik, B, D, d, udim, R, H)
    st0 = {k: v.clone() for k, v in m.full_state().items()}
    (mu, lv, losses), kind = _run(m, y, u, eps, 0)
    assert kind == 3, "the launch did not take the wide-observation path"
    assert m.status() == 0
    got = cuda_mod.state_np(m)
    # the fp64 oracle
    o = O.OracleVJF(D, d, udim, R, H, lik, lr=1e-3, dtype=np.float64)
    o.set_state({k: v.cpu().numpy() for k, v in st0.items()})
    omu, olv, olosses = o.run(y.astype(np.float64), None if u is None else u.astype(np.float64), eps=eps.astype(np.float64))
    assert_close(mu, omu, 2e-4, 2e-5, "mu")
    assert_close(lv, olv, 2e-4, 2e-5, "logvar")
    assert_close(losses, olosses, 2e-4, 2e-3, "losses")
    rls = ("w_mean", "w_chol", "w_precision", "transition.logvar")
    compare_state(got, o.get_state(), rtol=3e-3, atol=3e-4, skip=rls)
    # the general persistent kernel from the same initial state
    m.load_full_state(st0)
    (mu2, lv2, losses2), kind2 = _run(m, y, u, eps, 1)
    assert kind2 == 0
    assert_close(mu, mu2, 1e-4, 1e-5, "mu vs general kernel")
    assert_close(lv, lv2, 1e-4, 1e-5, "logvar vs general kernel")
    assert_close(losses, losses2, 1e-4, 1e-3, "losses vs general kernel")
    got2 = cuda_mod.state_np(m)
    compare_state(got, got2, rtol=2e-3, atol=2e-4, skip=("w_chol", "w_precision"))


def test_wide_path_philox_equals_tape(cuda_mod):
    """In-kernel Philox draws of the wide path == the same numbers handed in as a tape (vjf_philox_normal)."""
    from vjf_b200 import _lib
    lik, B, D, d, R, H, T = "poisson", 200, 768, 8, 32, [64], 3
    rng = np.random.default_rng(9)
    y, _ = _data(rng, lik, T, B, D, d, 0)
    m = _make(lik, B, D, d, 0, R, H)
    st0 = {k: v.clone() for k, v in m.full_state().items()}
    (mu, lv, losses), kind = _run(m, y, None, None, 0)
    assert kind == 3
    lib = _lib.load()
    e = torch.empty(T, 2, B, d, device="cuda")
    for t in range(T):
        _lib.check(lib.vjf_philox_normal(4321, t, 0, B, d, C.c_void_p(e[t].data_ptr()), None))
    torch.cuda.synchronize()
    m.load_full_state(st0)
    (mu2, lv2, losses2), _ = _run(m, y, None, e.cpu().numpy(), 0)
    assert np.array_equal(mu, mu2) and np.array_equal(lv, lv2) and np.array_equal(losses, losses2)


def test_c4_shape_takes_the_wide_path(cuda_mod):
    """BASELINE configs[3] shapes at 4096 trials: the launch kind is 3 and the run is deterministic (fixed-order sums)."""
    lik, B, D, d, R, H, T = "poisson", 4096, 2000, 8, 64, [128], 3
    rng = np.random.default_rng(4)
    y, _ = _data(rng, lik, T, B, D, d, 0)
    eps = rng.normal(size=(T, 2, B, d)).astype(np.float32)
    m = _make(lik, B, D, d, 0, R, H)
    st0 = {k: v.clone() for k, v in m.full_state().items()}
    (mu, lv, losses), kind = _run(m, y, None, eps, 0)
    assert kind == 3 and m.status() == 0
    s1 = cuda_mod.state_np(m)
    m.load_full_state(st0)
    (mu2, lv2, losses2), _ = _run(m, y, None, eps, 0)
    s2 = cuda_mod.state_np(m)
    assert np.array_equal(mu, mu2) and np.array_equal(losses, losses2)
    for k in s1:
        assert np.array_equal(s1[k], s2[k]), k


def test_wide_path_host_entry_equals_device_entry(cuda_mod):
    """vjf_run_host (pinned host buffers, chunked H2D) on a wide shape == vjf_run on device buffers, bit for bit."""
    from vjf_b200 import _lib
    lik, B, D, d, R, H, T = "poisson", 260, 640, 4, 24, [64], 7
    rng = np.random.default_rng(12)
    y, _ = _data(rng, lik, T, B, D, d, 0)
    eps = rng.normal(size=(T, 2, B, d)).astype(np.float32)
    a = _make(lik, B, D, d, 0, R, H)
    st0 = {k: v.clone() for k, v in a.full_state().items()}
    (mu, lv, losses), kind = _run(a, y, None, eps, 0)
    assert kind == 3
    b = _make(lik, B, D, d, 0, R, H)
    b.load_full_state(st0)
    yh, eh = torch.as_tensor(y).contiguous().pin_memory(), torch.as_tensor(eps).contiguous().pin_memory()
    mu_h = torch.empty(T, B, d).pin_memory(); lv_h = torch.empty(T, B, d).pin_memory(); ls_h = torch.empty(T, 4).pin_memory()
    p = lambda t: C.c_void_p(t.data_ptr())
    flags = _lib.FLAG_SGD | _lib.FLAG_UPDATE | _lib.FLAG_PRIOR_Q0
    _lib.check(_lib.load().vjf_run_host(b._h, T, B, p(yh), 0, None, p(eh), 0, 0, flags, b.lr, p(mu_h), p(lv_h), p(ls_h), 3))
    assert _lib.load().vjf_last_launch_kind() == 3
    assert np.array_equal(mu, mu_h.numpy()) and np.array_equal(lv, lv_h.numpy()) and np.array_equal(losses, ls_h.numpy())
    assert torch.equal(a._flat, b._flat)
