"""Parity of the CUDA path at the shapes bench.py actually times: the C2 headline
configuration (4096 trials: 28-trial tiles, tcgen05 weight gradient with 4 k-steps, early TMA issue, in-kernel Philox),
the multi-tile regime (16 384 trials), the C4 shape at 2048 trials, pre-clip gradients against the reference's autograd
gradients, and a freshly constructed model against the reference's own initial values.  Needs a B200: pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import vjf_oracle as O
from tests.helpers import assert_close, compare_state, load_golden, sub

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda_mod():
    from tests import gpu_helpers
    assert torch.cuda.is_available(), "GPU tests selected without a CUDA device"
    return gpu_helpers


def _poisson_counts(rng, T, B, D, d):
    """Counts driven by a smooth low-dimensional latent (rates ~0.1-3), like the benchmark's Lorenz recipe."""
    t = np.arange(T)[:, None, None] * 0.05
    ph = rng.uniform(0, 2 * np.pi, (1, B, d))
    x = np.sin(t * (1 + np.arange(d)) + ph)
    Cm = rng.normal(size=(d, D)) / np.sqrt(d)
    return rng.poisson(np.exp(np.clip(x @ Cm - 1.0, None, 3.0))).astype(np.float32)


def _run_vs_oracle(cuda_mod, lik, B, D, d, R, H, T, noise, seed=11, lr=1e-3):
    from vjf_b200 import _lib
    from vjf_b200.model import VJF
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    m = VJF.make_model(D, d, 0, R, H, lik, lr=lr, max_trials=B, seed=4321)
    o = O.OracleVJF(D, d, 0, R, H, lik, lr=lr, dtype=np.float64)
    st0 = cuda_mod.state_np(m)
    o.set_state(st0)
    y = _poisson_counts(rng, T, B, D, d) if lik == "poisson" else rng.normal(size=(T, B, D)).astype(np.float32)
    if noise == "tape":
        eps = rng.normal(size=(T, 2, B, d)).astype(np.float32)
        mu, lv, losses = m.run(torch.as_tensor(y), None, None, eps=torch.as_tensor(eps))
    else:
        # in-kernel Philox draws; the oracle gets the same numbers from vjf_philox_normal (the same device code)
        lib = _lib.load()
        e = torch.empty(T, 2, B, d, device="cuda")
        for t in range(T):
            _lib.check(lib.vjf_philox_normal(4321, t, 0, B, d, C.c_void_p(e[t].data_ptr()), None))
        torch.cuda.synchronize()
        eps = e.cpu().numpy()
        mu, lv, losses = m.run(torch.as_tensor(y))
    omu, olv, olosses = o.run(y.astype(np.float64), None, eps=eps.astype(np.float64))
    assert_close(mu.cpu().numpy(), omu, 2e-4, 2e-5, "mu")
    assert_close(lv.cpu().numpy(), olv, 2e-4, 2e-5, "logvar")
    assert_close(losses.cpu().numpy(), olosses, 2e-4, 2e-3, "losses")
    # Everything except the RLS solution at the fixed fp32 tolerance.  T*B samples go into the fp32 information-form RLS and
    # its rounding error grows with the sample count and the conditioning of P: the RLS tensors are judged like
    # SURVEY.md section 8c asks -- no further from the fp64 run than an fp32 run of the reference algorithm is (the fp32
    # oracle on the same inputs), with a factor 4 of slack.
    rls = ("w_mean", "w_chol", "w_precision", "transition.logvar")
    got, want = cuda_mod.state_np(m), o.get_state()
    compare_state(got, want, rtol=3e-3, atol=3e-4, skip=rls)
    o32 = O.OracleVJF(D, d, 0, R, H, lik, lr=lr, dtype=np.float32)
    o32.set_state(st0)
    o32.run(y, None, eps=eps.astype(np.float32))
    ref32 = o32.get_state()
    for k in ("w_mean", "w_precision", "transition.logvar"):
        scale = float(np.abs(want[k]).max())
        e_ref = float(np.abs(np.asarray(ref32[k], np.float64) - want[k]).max())
        e_got = float(np.abs(np.asarray(got[k], np.float64) - want[k]).max())
        assert e_got <= 4 * e_ref + 2e-5 * max(scale, 1.0), (k, e_got, e_ref, scale)
    assert m.status() == 0
    return m


@pytest.mark.parametrize("noise", ["tape", "philox"])
def test_c2_bench_configuration_matches_oracle(cuda_mod, noise):
    """Exactly what bench.py times at N=1: C2 shapes, 4096 trials per step, 8 steps."""
    _run_vs_oracle(cuda_mod, "poisson", 4096, 200, 3, 50, [64], 8, noise)


def test_c2_multi_tile_16384_trials_matches_oracle(cuda_mod):
    _run_vs_oracle(cuda_mod, "poisson", 16384, 200, 3, 50, [64], 4, "tape")


def test_c4_shape_2048_trials_matches_oracle(cuda_mod):
    """BASELINE configs[3] shapes: ydim 2000 Poisson, xdim 8, 64 RBFs, hidden [128]."""
    _run_vs_oracle(cuda_mod, "poisson", 2048, 2000, 8, 64, [128], 4, "tape")


def test_c3_dims_gaussian_small_rbf_matches_oracle(cuda_mod):
    """BASELINE configs[2] shapes except n_rbf: ydim 500 Gaussian, xdim 10, hidden [128]."""
    _run_vs_oracle(cuda_mod, "gaussian", 1024, 500, 10, 96, [128], 4, "tape")


@pytest.mark.parametrize("name", ["grads_pois", "grads_gauss", "grads_gauss_warm"])
def test_preclip_gradients_match_reference_autograd(cuda_mod, name):
    """GPU hand-derived backward vs the reference's autograd gradients read before the clip (vjf/model.py:209-210)."""
    from vjf_b200.model import VJF, Gaussian
    g = load_golden(f"{name}_f32")
    ydim, xdim, udim, n_rbf, B, _ = [int(v) for v in g["cfg"]]
    m = VJF.make_model(ydim, xdim, udim, n_rbf, [int(h) for h in g["hidden"]], str(g["lik"]), max_trials=B)
    m.load_full_state(sub(g, "state."))
    qs = Gaussian(torch.as_tensor(g["q_mean"]), torch.as_tensor(g["q_logvar"]))
    grads, qt = m.loss_gradients(g["y"], g.get("u"), qs, warm_up=bool(g["warm_up"]), eps=g["eps"])
    assert_close(qt.mean.cpu().numpy(), g["qt_mean"], 5e-5, 5e-6, "qt.mean")
    assert_close(qt.logvar.cpu().numpy(), g["qt_logvar"], 5e-5, 5e-6, "qt.logvar")
    want = sub(g, "grad.")
    assert set(want) <= set(grads), set(want) - set(grads)
    for k, v in want.items():
        scale = max(1e-3, float(np.abs(v).max()))
        assert_close(grads[k].cpu().numpy(), v, 1e-4, 2e-5 * scale, f"grad {k}")


def test_fresh_model_has_the_reference_initial_values(cuda_mod):
    """VJF.make_model's deterministic initial tensors against the reference's own freshly constructed
    model (golden init.*): W = 0, w_chol = P = I, logwidth = 0, obs logvar = log 0.1, state logvar = 0, prior 0, counters 0;
    the random tensors have the reference's shapes and ranges."""
    from vjf_b200.model import VJF
    for name in ("c1_gauss_f32", "pois_u_f32"):
        g = load_golden(name)
        ydim, xdim, udim, n_rbf, B, _ = [int(v) for v in g["cfg"]]
        m = VJF.make_model(ydim, xdim, udim, n_rbf, [int(h) for h in g["hidden"]], str(g["lik"]), max_trials=B)
        got, want = cuda_mod.state_np(m), sub(g, "init.")
        assert set(want) <= set(got) | {"w_pchol"}, set(want) - set(got)
        for k in ("mean", "logvar", "transition.logvar", "transition.velocity.feature.logwidth", "w_mean", "w_chol", "w_precision",
                  "likelihood.logvar", "likelihood.n_sample", "transition.n_sample"):
            if k in want:
                assert_close(got[k], want[k], 0, 1e-7, k)
        assert np.array_equal(got["w_pchol"], np.eye(n_rbf, dtype=np.float32))
        for k, v in want.items():
            assert got[k].shape == v.shape, (k, got[k].shape, v.shape)
            if k.endswith("weight") or k.endswith("bias") or k.endswith("centroid"):
                lim = 2.0 if k.endswith("centroid") else 1.0 / np.sqrt(v.shape[-1] if k.endswith("weight") else
                                                                     want[k[:-4] + "weight"].shape[-1])
                assert np.abs(got[k]).max() <= lim * (1 + 1e-6), k
                assert np.abs(got[k]).max() > 0.5 * lim or got[k].size < 8, k  # drawn from the full range, not zeros
        assert sorted(m.state_dict().keys()) == sorted(k for k in want if k not in ("w_mean", "w_chol", "w_precision", "likelihood.n_sample", "transition.n_sample"))
