"""Pin the numpy oracle against outputs of the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest

from oracle import vjf_oracle as O
from tests.helpers import assert_close, compare_state, load_golden, oracle_from_golden, run_phases, sub

RUNS = ["c1_gauss", "pois_u", "gauss_phases", "c2_small"]
# fp64: differences are pure summation-order round-off; fp32: round-off amplified over the run
TOL = {"f64": dict(rtol=1e-8, atol=1e-9), "f32": dict(rtol=3e-3, atol=3e-4)}


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("name", RUNS)
def test_run_matches_reference(name, tag):
    g = load_golden(f"{name}_{tag}")
    m = oracle_from_golden(g)
    mu, lv, losses = run_phases(m, g)
    tol = TOL[tag]
    assert_close(mu, g["mu"], what="mu", **tol)
    assert_close(lv, g["logvar"], what="logvar", **tol)
    assert_close(losses, g["losses"], what="losses", rtol=tol["rtol"], atol=tol["atol"] * 10)
    skip = ()
    if tag == "f32" and name == "c1_gauss":
        # R=100 RBFs fed one sample per step with a shrinking state-noise estimate: P reaches 1e4 and the
        # fp32 RLS solution is ill-conditioned -- the reference's own fp32 run is 0.15 away from the fp64
        # run of the same recipe.  Judge the RLS state against that yardstick instead of a fixed tolerance.
        skip = ("w_mean", "w_chol", "w_precision")
        m64 = oracle_from_golden(g, dtype=np.float64)
        run_phases(m64, g)
        ref_err = np.abs(g["final.w_mean"] - m64.w_mean).max()
        our_err = np.abs(m.w_mean - m64.w_mean).max()
        assert our_err <= 2 * ref_err + 1e-4, (our_err, ref_err)
    compare_state(m.get_state(), sub(g, "final."), skip=skip, **tol)
    assert m.status == 0


@pytest.mark.parametrize("name", RUNS)
def test_first_step_f32_tight(name):
    """One step from identical state: fp32 differences are a few ulp, no amplification yet."""
    g = load_golden(f"{name}_f32")
    m = oracle_from_golden(g)
    n, sgd, upd, warm, _ = g["phases"][0]
    u = g.get("u")
    q, loss, a, b, c = m.filter(g["y"][0], None if u is None else u[0], None, eps=g["eps"][0], sgd=bool(sgd),
                                update=bool(upd), verbose=True, warm_up=bool(warm))
    assert_close(q.mean, g["mu"][0], 2e-5, 2e-6, "mu[0]")
    assert_close(q.logvar, g["logvar"][0], 2e-5, 2e-6, "logvar[0]")
    assert_close([loss, a, b, c], g["losses"][0], 2e-5, 2e-5, "losses[0]")


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("name", ["grads_pois", "grads_gauss", "grads_gauss_warm"])
def test_hand_derived_backward_matches_autograd(name, tag):
    g = load_golden(f"{name}_{tag}")
    ydim, xdim, udim, n_rbf, B, _ = [int(v) for v in g["cfg"]]
    m = O.OracleVJF(ydim, xdim, udim, n_rbf, [int(h) for h in g["hidden"]], str(g["lik"]), dtype=g["y"].dtype)
    m.set_state(sub(g, "state."))
    qs = O.Gaussian(g["q_mean"], g["q_logvar"])
    qt, loss, a, b, c, grads = m.filter(g["y"], g.get("u"), qs, eps=g["eps"], sgd=False, update=False, verbose=True,
                                        warm_up=bool(g["warm_up"]), return_grads=True)
    tol = dict(rtol=1e-9, atol=1e-11) if tag == "f64" else dict(rtol=2e-4, atol=2e-6)
    assert_close(qt.mean, g["qt_mean"], what="qt.mean", **tol)
    assert_close(qt.logvar, g["qt_logvar"], what="qt.logvar", **tol)
    assert_close([loss, a, b, c], g["losses"], what="losses", rtol=tol["rtol"], atol=tol["atol"] * 100)
    want = sub(g, "grad.")
    assert want, "fixture holds no gradients"
    for k, v in want.items():
        assert_close(grads[k], v, what="grad " + k, **tol)


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_kalman_matches_reference(tag):
    g = load_golden(f"kalman_{tag}")
    tol = dict(rtol=1e-9, atol=1e-10) if tag == "f64" else dict(rtol=2e-4, atol=2e-5)
    yhat, xhat, Lhat = O.kalman_predict(g["x"], g["L0"], g["A"], g["Q"], g["H"])
    assert_close(yhat, g["yhat"], what="yhat", **tol)
    assert_close(xhat, g["xhat"], what="xhat", **tol)
    assert_close(Lhat, g["Lhat"], what="Lhat", **tol)
    x, L = O.kalman_update(g["y"], g["yhat"], g["xhat"], g["Lhat"], g["H"], g["R"])
    assert_close(x, g["x_upd"], what="x_upd", **tol)
    assert_close(L, g["L_upd"], what="L_upd", **tol)
    x, L = O.kalman_joseph_update(g["y"], g["yhat"], g["xhat"], g["Lhat"], g["H"], g["R"])
    assert_close(x, g["x_jos"], what="x_jos", **tol)
    assert_close(L, g["L_jos"], what="L_jos", **tol)
    assert_close(O.symmetrize(g["sym_in"]), g["sym_out"], what="symmetrize", **tol)
    assert_close(O.positivize(g["pos_in"]), g["pos_out"], what="positivize", rtol=tol["rtol"] * 10, atol=tol["atol"] * 10)


def test_nonfinite_terms_are_zeroed():
    """vjf/model.py:138-145: a non-finite ELBO term becomes the constant 0 and carries no gradient."""
    m = O.OracleVJF(6, 2, 0, 5, [4], "poisson", dtype=np.float64)
    rng = np.random.default_rng(0)
    y = rng.poisson(1.0, (3, 6)).astype(np.float64)
    qs = O.Gaussian(np.zeros((3, 2)), np.full((3, 2), 2000.0))  # exp(p_logvar + l_t - gamma) overflows later
    m.head_v_b[:] = 800.0  # l_t huge -> trace term exp(.) = inf -> l_dyn = inf
    eps = np.zeros((2, 3, 2))
    qt, loss, a, b, c, grads = m.filter(y, None, O.Gaussian(np.zeros((3, 2)), np.zeros((3, 2))), eps=eps, sgd=False,
                                        update=False, verbose=True, return_grads=True)
    assert b == 0 and m.status & O.ST_DYN_NONFINITE
    assert all(np.all(np.isfinite(v)) for v in grads.values())


# ---- rows next to the hot path: fixtures of tests/golden/make_golden_next.py ----
NEXT_TOL = {"f64": dict(rtol=1e-7, atol=1e-9), "f32": dict(rtol=3e-3, atol=3e-4)}


def _oracle_with(g, prefix, dtype=None):
    ydim, xdim, udim, n_rbf, B, T = [int(v) for v in g["cfg"]]
    dt = dtype or g[[k for k in g if k.startswith(prefix)][0]].dtype
    m = O.OracleVJF(ydim, xdim, udim, n_rbf, [int(h) for h in g["hidden"]], str(g["lik"]), dtype=dt)
    m.set_state(sub(g, prefix))
    return m


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_initialize_matches_reference(tag):
    """RBFDS.initialize + LinearRegression.initialize (vjf/model.py:379-388, vjf/module.py:144-150)."""
    g = load_golden(f"init_{tag}")
    m = _oracle_with(g, "before.", g["xs"].dtype)
    st, r = m.initialize_transition(g["xt"], g["xs"], g.get("ut"), centroid=g["after.transition.velocity.feature.centroid"])
    assert st == 0
    skip = ()
    if tag == "f32":
        # wide RBFs (width r) make phi^T phi ill-conditioned: the fp32 RLS solution of the REFERENCE is itself ~1e-3 away from
        # the fp64 solution of the same problem.  Judge the weights against that yardstick, the rest at the usual tolerance.
        skip = ("w_mean", "w_chol", "w_precision")
        m64 = _oracle_with(g, "before.", np.float64)
        m64.initialize_transition(g["xt"], g["xs"], g.get("ut"), centroid=g["after.transition.velocity.feature.centroid"])
        ref_err = np.abs(g["after.w_mean"] - m64.w_mean).max()
        assert np.abs(m.w_mean - m64.w_mean).max() <= 2 * ref_err + 1e-4
        assert_close(m.w_precision, g["after.w_precision"], 1e-4, 1e-5, "w_precision")
    compare_state(m.get_state(), sub(g, "after."), skip=skip, **NEXT_TOL[tag])


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_forecast_matches_reference(tag):
    """VJF.forecast -> RBFDS.forecast with sampled weights (vjf/model.py:321-324, :342-361; vjf/module.py:70-73)."""
    g = load_golden(f"forecast_{tag}")
    m = _oracle_with(g, "state.", g["x0"].dtype)
    n_step = int(g["cfg"][5])
    x, y = m.forecast(g["x0"], g.get("u"), n_step, noise=True, w_eps=g["w_eps"], x_eps=g["x_eps"])
    tol = dict(rtol=1e-9, atol=1e-10) if tag == "f64" else dict(rtol=2e-4, atol=2e-5)
    assert_close(x, g["x"], what="x", **tol)
    assert_close(y, g["y"], what="y", **tol)


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_fit_flow_matches_reference(tag):
    """VJF.fit across the warm-up boundary (vjf/model.py:223-307): same epoch count, same epoch of the warm-up exit,
    decoder frozen from there on, transition re-initialised, lr decayed once per completed epoch."""
    g = load_golden(f"fit_{tag}")
    m = oracle_from_golden(g, lr=float(g["lr"]))
    m.lr_decay = float(g["lr_decay"])
    seen = {}

    def centroid(r):
        seen["epoch"] = True
        return g["init_centroid"]
    mu, lv, epoch_loss, n_epochs = m.fit(g["y"], g.get("u"), eps=g["eps"], centroid_unit=centroid, max_iter=int(g["max_iter"]),
                                         rtol=float(g["rtol"]))
    assert n_epochs == int(g["n_epochs"]) and seen and m.decoder_frozen == bool(g["decoder_frozen"])
    tol = NEXT_TOL[tag]
    assert_close(m.lr, g["final_lr"], 1e-12, 0, "lr")
    assert_close(mu, g["mu"], what="mu", **tol)
    assert_close(lv, g["logvar"], what="logvar", **tol)
    assert_close(epoch_loss, g["epoch_loss"], what="epoch_loss", **tol)
    skip = ()
    if tag == "f32":  # same yardstick as test_initialize_matches_reference for the ill-conditioned RLS weights
        skip = ("w_mean", "w_chol", "w_precision")
        m64 = oracle_from_golden(g, lr=float(g["lr"]), dtype=np.float64)
        m64.lr_decay = float(g["lr_decay"])
        m64.fit(g["y"], g.get("u"), eps=g["eps"], centroid_unit=lambda r: g["init_centroid"], max_iter=int(g["max_iter"]), rtol=float(g["rtol"]))
        ref_err = np.abs(g["final.w_mean"] - m64.w_mean).max()
        assert np.abs(m.w_mean - m64.w_mean).max() <= 2 * ref_err + 1e-4
    compare_state(m.get_state(), sub(g, "final."), skip=skip, **tol)


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_weight_kalman_matches_reference(tag):
    """LinearRegression.kalman (vjf/module.py:114-142)."""
    g = load_golden(f"wkalman_{tag}")
    m = _oracle_with(g, "before.", g["xu"].dtype)
    m.weight_kalman(g["xu"], g["target"], float(g["v"]), float(g["diffusion"]))
    tol = dict(rtol=1e-8, atol=1e-10) if tag == "f64" else dict(rtol=2e-3, atol=2e-5)
    assert_close(m.w_mean, g["w_mean"], what="w_mean", **tol)
    assert_close(m.w_chol @ m.w_chol.T, g["w_chol"] @ g["w_chol"].T, what="w_chol w_chol^T", **tol)
