"""CPU oracle for the VJF per-time-step filter + learning step.  TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the reference algorithm (catniplab/vjf, pure
Python/PyTorch) with a HAND-DERIVED backward pass -- it never calls autograd, so it is an
independent check of the gradient formulas the CUDA kernels implement.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it; the product path (``vjf_b200``) never does and fails loudly without its CUDA library.

Pinning: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference (imported from
/root/reference in the build container) with a shared noise tape and stores its outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle against every one of
them (fp64 to ~1e-10, fp32 to fp32 round-off).  The reference's own tests hold no numerical
golden vectors, so those fixtures are the pin.

Parameter arrays use the reference (torch ``nn.Linear``) layout: ``weight[out, in]``.

Every function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import List, Optional, Sequence

import numpy as np
import scipy.linalg as sla

Gaussian = namedtuple("Gaussian", ["mean", "logvar"])  # vjf/distribution.py:3

# status bits (mirrors include/vjf_b200.h)
ST_RECON_NONFINITE = 1
ST_DYN_NONFINITE = 2
ST_ENTROPY_NONFINITE = 4
ST_MSE_NONFINITE = 8
ST_CHOL_FAILED = 16


# --------------------------------------------------------------------------------------------
# small functional pieces
# --------------------------------------------------------------------------------------------
def reparametrize(q: Gaussian, eps: np.ndarray) -> np.ndarray:
    """vjf/util.py:11-13 with the N(0,1) draw supplied from a noise tape."""
    return q.mean + eps * np.exp(0.5 * q.logvar)


def running_var(acc_var, acc_size: int, new_var, new_size: int, size_cap: int = 1000):
    """vjf/util.py:20-35.  f1/f2 are Python floats (double) as in the reference."""
    acc_size = min(acc_size, size_cap)
    tot = acc_size + new_size
    f1 = acc_size / tot
    f2 = new_size / tot
    dt = np.asarray(new_var).dtype
    return dt.type(f1) * acc_var + dt.type(f2) * new_var, tot


def rbf(x: np.ndarray, c: np.ndarray, w: np.ndarray) -> np.ndarray:
    """vjf/functional.py:11-22: exp(-0.5 * (||x-c|| / w)^2).

    The reference goes through torch.cdist (matmul form + sqrt) and squares again; the direct
    sum of squared differences used here is the more accurate of the two.
    """
    diff = x[:, None, :] - c[None, :, :]
    d2 = np.sum(diff * diff, axis=-1)
    return np.exp(-0.5 * d2 / (w * w)[None, :])


def gaussian_entropy(q: Gaussian):
    """vjf/functional.py:25-29."""
    return 0.5 * q.logvar.sum(-1).mean()


def gaussian_loss(m1, logv1, m2, logv2, logvar):
    """vjf/functional.py:32-75 (expected Gaussian NLL; trace = exp(logv1 + logv2 - logvar))."""
    p = np.exp(-0.5 * logvar)
    mse = (m1 * p - m2 * p) ** 2
    finite = bool(np.all(np.isfinite(mse)))  # functional.py:60 assert
    nll = 0.5 * (mse + logvar)
    if logv1 is None and logv2 is None:
        trace = 0.0
    elif logv2 is None:
        trace = np.exp(logv1 - logvar)
    elif logv1 is None:
        trace = np.exp(logv2 - logvar)
    else:
        trace = np.exp(logv1 + logv2 - logvar)
    with np.errstate(over="ignore", invalid="ignore"):
        nll = nll + 0.5 * trace
    return nll.sum(-1).mean(), finite


# --------------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------------
class OracleVJF:
    """Restatement of ``vjf.model.VJF`` (vjf/model.py:50-324) + RBFDS (:327-391) +
    LinearRegression/RBF (vjf/module.py:14-150) + Recognition (vjf/recognition.py:16-42) +
    likelihoods (vjf/likelihood.py)."""

    def __init__(self, ydim, xdim, udim, n_rbf, hidden_sizes: Sequence[int], likelihood="poisson",
                 lr=1e-4, lr_decay=0.9, dtype=np.float32, rng: Optional[np.random.Generator] = None):
        self.ydim, self.xdim, self.udim, self.n_rbf = ydim, xdim, udim, n_rbf
        self.hidden = list(hidden_sizes)
        self.likelihood = likelihood.lower()
        assert self.likelihood in ("poisson", "gaussian")
        self.lr, self.lr_decay = float(lr), float(lr_decay)
        self.dtype = np.dtype(dtype)
        T = self.dtype.type
        rng = rng or np.random.default_rng(0)
        # vjf/model.py:66-67 prior (zeros, never stepped: not in the optimizer, :69-77)
        self.prior_mean = np.zeros(xdim, self.dtype)
        self.prior_logvar = np.zeros(xdim, self.dtype)
        # vjf/likelihood.py:16-17
        self.lik_logvar = T(math.log(0.1))
        self.lik_n = 0
        # vjf/model.py:331-332
        self.tr_logvar = T(0.0)
        self.tr_n = 0
        # vjf/module.py:20-21,46-54
        self.centroid = (rng.random((n_rbf, xdim + udim)) * 4 - 2).astype(self.dtype)
        self.logwidth = np.zeros(n_rbf, self.dtype)
        self.w_mean = np.zeros((n_rbf, xdim), self.dtype)
        self.w_chol = np.eye(n_rbf, dtype=self.dtype)
        self.w_precision = np.eye(n_rbf, dtype=self.dtype)
        self.w_pchol = np.eye(n_rbf, dtype=self.dtype)
        # vjf/recognition.py:20-28 (init: nn.Linear default U(-1/sqrt(in), 1/sqrt(in)))
        sizes = [ydim + udim + 2 * xdim] + self.hidden
        self.mlp_w: List[np.ndarray] = []
        self.mlp_b: List[np.ndarray] = []
        for i in range(len(self.hidden)):
            k = 1 / math.sqrt(sizes[i])
            self.mlp_w.append(rng.uniform(-k, k, (sizes[i + 1], sizes[i])).astype(self.dtype))
            self.mlp_b.append(rng.uniform(-k, k, sizes[i + 1]).astype(self.dtype))
        k = 1 / math.sqrt(self.hidden[-1])
        self.head_m_w = rng.uniform(-k, k, (xdim, self.hidden[-1])).astype(self.dtype)
        self.head_v_w = rng.uniform(-k, k, (xdim, self.hidden[-1])).astype(self.dtype)
        self.head_v_b = rng.uniform(-k, k, xdim).astype(self.dtype)
        # vjf/model.py:24
        k = 1 / math.sqrt(xdim)
        self.dec_w = rng.uniform(-k, k, (ydim, xdim)).astype(self.dtype)
        self.dec_b = rng.uniform(-k, k, ydim).astype(self.dtype)
        self.decoder_frozen = False  # vjf/model.py:283
        self.status = 0

    # ---- state exchange (names follow the reference state_dict + plain attributes) ----
    STATE_KEYS = ("mean", "logvar", "likelihood.logvar", "transition.logvar",
                  "transition.velocity.feature.centroid", "transition.velocity.feature.logwidth",
                  "recognition.mean.weight", "recognition.logvar.weight", "recognition.logvar.bias",
                  "decoder.decode.weight", "decoder.decode.bias",
                  "w_mean", "w_chol", "w_precision", "likelihood.n_sample", "transition.n_sample")

    def get_state(self) -> dict:
        s = {
            "mean": self.prior_mean, "logvar": self.prior_logvar,
            "likelihood.logvar": np.asarray(self.lik_logvar), "transition.logvar": np.asarray(self.tr_logvar),
            "transition.velocity.feature.centroid": self.centroid,
            "transition.velocity.feature.logwidth": self.logwidth,
            "recognition.mean.weight": self.head_m_w, "recognition.logvar.weight": self.head_v_w,
            "recognition.logvar.bias": self.head_v_b,
            "decoder.decode.weight": self.dec_w, "decoder.decode.bias": self.dec_b,
            "w_mean": self.w_mean, "w_chol": self.w_chol, "w_precision": self.w_precision,
            "likelihood.n_sample": np.asarray(self.lik_n), "transition.n_sample": np.asarray(self.tr_n),
        }
        for i in range(len(self.hidden)):
            s[f"recognition.mlp.{2 * i}.weight"] = self.mlp_w[i]
            s[f"recognition.mlp.{2 * i}.bias"] = self.mlp_b[i]
        return {k: np.array(v, copy=True) for k, v in s.items()}

    def set_state(self, s: dict):
        dt = self.dtype
        g = lambda k: np.array(s[k], dtype=dt, copy=True)
        self.prior_mean, self.prior_logvar = g("mean"), g("logvar")
        self.lik_logvar = dt.type(s["likelihood.logvar"]) if "likelihood.logvar" in s else self.lik_logvar
        self.tr_logvar = dt.type(s["transition.logvar"])
        self.centroid = g("transition.velocity.feature.centroid")
        self.logwidth = g("transition.velocity.feature.logwidth")
        self.head_m_w, self.head_v_w = g("recognition.mean.weight"), g("recognition.logvar.weight")
        self.head_v_b = g("recognition.logvar.bias")
        self.dec_w, self.dec_b = g("decoder.decode.weight"), g("decoder.decode.bias")
        for i in range(len(self.hidden)):
            self.mlp_w[i] = g(f"recognition.mlp.{2 * i}.weight")
            self.mlp_b[i] = g(f"recognition.mlp.{2 * i}.bias")
        if "w_mean" in s:
            self.w_mean, self.w_chol, self.w_precision = g("w_mean"), g("w_chol"), g("w_precision")
        if "likelihood.n_sample" in s:
            self.lik_n = int(s["likelihood.n_sample"])
        if "transition.n_sample" in s:
            self.tr_n = int(s["transition.n_sample"])

    # ---- sub-modules ----
    def feature(self, xu):
        """RBF.forward vjf/module.py:30-34 (no intercept)."""
        return rbf(xu, self.centroid, np.exp(self.logwidth))

    def velocity(self, xu) -> Gaussian:
        """LinearRegression.forward(sampling=False) vjf/module.py:56-77.  The reference forms the
        (B,B) matrix FL FL^T and reads its diagonal (:76); the diagonal is the row-wise squared norm."""
        phi = self.feature(xu)
        FL = phi @ self.w_chol
        with np.errstate(divide="ignore"):
            logvar = np.log(np.sum(FL * FL, axis=1))
        return Gaussian(phi @ self.w_mean, np.tile(logvar[:, None], (1, self.xdim))), phi

    def recognition(self, y, qs: Gaussian, u):
        """Recognition.forward vjf/recognition.py:31-42; returns the activations for the backward."""
        parts = [y] + ([u] if (u is not None and self.udim > 0) else []) + [qs.mean, qs.logvar]
        acts = [np.concatenate(parts, axis=-1)]
        for W, b in zip(self.mlp_w, self.mlp_b):
            acts.append(np.tanh(acts[-1] @ W.T + b))
        h = acts[-1]
        return Gaussian(h @ self.head_m_w.T, h @ self.head_v_w.T + self.head_v_b), acts

    def prior(self, n_batch) -> Gaussian:
        """VJF.prior vjf/model.py:80-95."""
        one = np.ones((n_batch, self.xdim), self.dtype)
        return Gaussian(one * self.prior_mean, one * self.prior_logvar)

    # ---- the hot path ----
    def filter(self, y, u=None, qs: Optional[Gaussian] = None, *, eps, sgd=True, update=True,
               verbose=False, warm_up=False, return_grads=False):
        """VJF.filter vjf/model.py:179-221.  ``eps`` = (2, B, xdim) noise tape: eps[0] is the draw for
        xs (model.py:112), eps[1] for xt (:119)."""
        dt = self.dtype
        y = np.atleast_2d(np.asarray(y, dtype=dt))
        if u is not None:
            u = np.atleast_2d(np.asarray(u, dtype=dt))
            if self.udim == 0:
                u = None
        B = y.shape[0]
        eps = np.asarray(eps, dtype=dt).reshape(2, B, self.xdim)
        if qs is None:
            qs = self.prior(B)
        qs = Gaussian(np.asarray(qs.mean, dt), np.asarray(qs.logvar, dt))

        # ---- forward (model.py:97-122) ----
        xs = reparametrize(qs, eps[0])
        xu = xs if u is None else np.concatenate([xs, u], -1)  # util.py:38-49
        dxq, phi = self.velocity(xu)
        pt = Gaussian(xs + dxq.mean, dxq.logvar)  # RBFDS.forward model.py:334-340, leak = 0
        qt, acts = self.recognition(y, qs, u)
        xt = reparametrize(qt, eps[1])
        eta = xt @ self.dec_w.T + self.dec_b  # LinearDecoder model.py:29-30

        # ---- loss (model.py:124-154) ----
        status = 0
        if self.likelihood == "gaussian":
            l_recon, fin = gaussian_loss(y, None, eta, None, self.lik_logvar)  # likelihood.py:26
            if not fin:
                status |= ST_MSE_NONFINITE
        else:
            etac = np.minimum(eta, dt.type(10.0))  # likelihood.py:60
            with np.errstate(over="ignore", invalid="ignore"):
                l_recon = (np.exp(etac) - y * etac).sum(-1).mean()
        with np.errstate(over="ignore", invalid="ignore"):
            l_dyn, fin = gaussian_loss(pt.mean, pt.logvar, qt.mean, qt.logvar, self.tr_logvar)  # model.py:390-391
        if not fin:
            status |= ST_MSE_NONFINITE
        h = gaussian_entropy(qt)
        r_on = bool(np.isfinite(l_recon))
        d_fin = bool(np.isfinite(l_dyn))
        h_on = bool(np.isfinite(h))
        if not r_on:
            l_recon = dt.type(0); status |= ST_RECON_NONFINITE
        if not d_fin:
            l_dyn = dt.type(0); status |= ST_DYN_NONFINITE
        if not h_on:
            h = dt.type(0); status |= ST_ENTROPY_NONFINITE
        loss = l_recon - h
        if not warm_up:
            loss = loss + l_dyn
        d_on = d_fin and not warm_up

        grads = None
        if sgd or return_grads:
            grads = self._backward(y, qs, pt, qt, acts, xt, eta, eps[1], r_on, d_on, h_on)
        if sgd:
            self._sgd(grads)
        if update:
            status |= self._update(y, eta, xt, xs, xu, phi, warm_up)
        self.status |= status

        out = (qt, loss) + ((-l_recon, -l_dyn, h) if verbose else ())
        if return_grads:
            out = out + (grads,)
        return out

    def _backward(self, y, qs, pt, qt, acts, xt, eta, eps2, r_on, d_on, h_on) -> dict:
        """Hand-derived gradient of VJF.loss (what autograd computes at vjf/model.py:209), see
        SURVEY.md section 8a row A17.  Terms whose loss was non-finite carry no gradient
        (model.py:138-145 replaces them by constants)."""
        dt = self.dtype
        B = y.shape[0]
        invB = dt.type(1.0 / B)
        g = {}
        if self.likelihood == "gaussian":
            il = np.exp(-self.lik_logvar)
            r = y - eta
            g_eta = (-r * il * invB) if r_on else np.zeros_like(eta)
            g["likelihood.logvar"] = (0.5 * (1.0 - r * r * il)).sum(-1).mean() if r_on else dt.type(0)
        else:
            etac = np.minimum(eta, dt.type(10.0))
            g_eta = ((np.exp(etac) - y) * (eta <= 10.0) * invB) if r_on else np.zeros_like(eta)
        g["decoder.decode.weight"] = g_eta.T @ xt
        g["decoder.decode.bias"] = g_eta.sum(0)
        g_xt = g_eta @ self.dec_w
        gam = self.tr_logvar
        g_mt = g_xt.copy()
        g_lt = 0.5 * g_xt * eps2 * np.exp(0.5 * qt.logvar)
        if h_on:
            g_lt = g_lt - 0.5 * invB
        if d_on:
            g_mt = g_mt + (qt.mean - pt.mean) * np.exp(-gam) * invB
            g_lt = g_lt + 0.5 * np.exp(pt.logvar + qt.logvar - gam) * invB
        hL = acts[-1]
        g["recognition.mean.weight"] = g_mt.T @ hL
        g["recognition.logvar.weight"] = g_lt.T @ hL
        g["recognition.logvar.bias"] = g_lt.sum(0)
        g_h = g_mt @ self.head_m_w + g_lt @ self.head_v_w
        for l in range(len(self.hidden) - 1, -1, -1):
            g_pre = g_h * (1.0 - acts[l + 1] ** 2)
            g[f"recognition.mlp.{2 * l}.weight"] = g_pre.T @ acts[l]
            g[f"recognition.mlp.{2 * l}.bias"] = g_pre.sum(0)
            if l > 0:
                g_h = g_pre @ self.mlp_w[l]
        return g

    def _sgd(self, g: dict):
        """clip_grad_value_(1.) + plain SGD, vjf/model.py:210-211 (optimizer :69-77)."""
        lr = self.dtype.type(self.lr)
        step = lambda p, gr: p - lr * np.clip(gr, -1.0, 1.0).astype(self.dtype)
        if self.likelihood == "gaussian":
            self.lik_logvar = self.dtype.type(step(self.lik_logvar, g["likelihood.logvar"]))
        if not self.decoder_frozen:
            self.dec_w = step(self.dec_w, g["decoder.decode.weight"])
            self.dec_b = step(self.dec_b, g["decoder.decode.bias"])
        self.head_m_w = step(self.head_m_w, g["recognition.mean.weight"])
        self.head_v_w = step(self.head_v_w, g["recognition.logvar.weight"])
        self.head_v_b = step(self.head_v_b, g["recognition.logvar.bias"])
        for l in range(len(self.hidden)):
            self.mlp_w[l] = step(self.mlp_w[l], g[f"recognition.mlp.{2 * l}.weight"])
            self.mlp_b[l] = step(self.mlp_b[l], g[f"recognition.mlp.{2 * l}.bias"])

    def rls(self, phi, target, v, shrink=1.0) -> int:
        """LinearRegression.rls vjf/module.py:79-112 (features already evaluated)."""
        dt = self.dtype
        P = self.w_precision
        s = np.sqrt(v)
        sf = phi / s
        st = target / s
        g = (P @ self.w_mean) * dt.type(shrink) + sf.T @ st
        P = P * dt.type(shrink) + sf.T @ sf
        try:
            L = np.linalg.cholesky(P)
        except np.linalg.LinAlgError:
            return ST_CHOL_FAILED  # the reference's fallback calls the removed torch.eig (module.py:106)
        self.w_pchol = L
        self.w_precision = P
        self.w_mean = sla.cho_solve((L, True), g).astype(dt)
        self.w_chol = sla.solve_triangular(L.T, np.eye(L.shape[0], dtype=dt), lower=False).astype(dt)
        return 0

    def _update(self, y, eta, xt, xs, xu, phi, warm_up) -> int:
        """VJF.update vjf/model.py:156-177 -> GaussianLikelihood.update likelihood.py:28-40 and
        RBFDS.update model.py:363-377."""
        dt = self.dtype
        B = y.shape[0]
        st = 0
        if self.likelihood == "gaussian":
            mse = ((y - eta) ** 2).mean()
            var, n = running_var(np.exp(self.lik_logvar), self.lik_n, mse, B)
            self.lik_logvar, self.lik_n = dt.type(np.log(var)), n
        dx = xt - xs
        if not warm_up:
            st |= self.rls(phi, dx, np.exp(self.tr_logvar), 1.0)
        resid = dx - phi @ self.w_mean
        mse = (resid ** 2).mean()
        var, n = running_var(np.exp(self.tr_logvar), self.tr_n, mse, B, size_cap=500)
        self.tr_logvar, self.tr_n = dt.type(np.log(var)), n
        return st

    # ---- epoch loop ----
    def run(self, y, u=None, *, eps, warm_up=False, sgd=True, update=True, q0: Optional[Gaussian] = None):
        """One pass of the time loop of VJF.fit (vjf/model.py:252-261) at a fixed warm-up phase.
        y: (T,B,D); eps: (T,2,B,d).  Returns mu(T,B,d), logvar(T,B,d), losses(T,4)."""
        T = y.shape[0]
        q = q0
        mus, lvs, losses = [], [], []
        for t in range(T):
            ut = None if u is None else u[t]
            q, loss, a, b, c = self.filter(y[t], ut, q, eps=eps[t], sgd=sgd, update=update,
                                           verbose=True, warm_up=warm_up)
            mus.append(q.mean); lvs.append(q.logvar); losses.append([loss, a, b, c])
        return np.stack(mus), np.stack(lvs), np.asarray(losses, dtype=self.dtype)

    def initialize_transition(self, xt, xs, ut=None, *, centroid: np.ndarray):
        """RBFDS.initialize vjf/model.py:379-388 + LinearRegression.initialize vjf/module.py:144-150.
        The U(-r, r) centroid re-draw comes from the caller (``centroid`` in [-1,1) units is scaled
        by r) so that runs are reproducible against the reference."""
        dt = self.dtype
        xu = xs if (ut is None or self.udim == 0) else np.concatenate([xs, ut], -1)
        mse = ((xt - xs) ** 2).mean()
        r = float(np.sqrt((xu * xu).sum(1)).max())
        self.centroid = np.asarray(centroid, dt)
        self.logwidth = np.full(self.n_rbf, math.log(r), dt)
        phi = self.feature(xu)
        st = self.rls(phi, xt - xs, mse)
        d = phi @ self.w_mean
        mse = ((xt - xs - d) ** 2).mean()
        self.tr_logvar = dt.type(np.log(mse))
        return st, r

    def forecast(self, x0, u=None, n_step=1, *, noise=False, w_eps=None, x_eps=None):
        """VJF.forecast vjf/model.py:321-324 -> RBFDS.forecast :342-361 with
        LinearRegression.forward(sampling=True) vjf/module.py:70-73.  w_eps: (n_step, R, d) draws
        for the weight sample, x_eps: (n_step, B, d) state-noise draws."""
        dt = self.dtype
        x0 = np.atleast_2d(np.asarray(x0, dt))
        x = np.empty((n_step + 1,) + x0.shape, dt)
        x[0] = x0
        s = np.exp(0.5 * self.tr_logvar)
        for t in range(n_step):
            xu = x[t] if (u is None or self.udim == 0) else np.concatenate([x[t], np.atleast_2d(u[t])], -1)
            w = self.w_mean + self.w_chol @ np.asarray(w_eps[t], dt)
            x[t + 1] = x[t] + self.feature(xu) @ w
            if noise:
                x[t + 1] = x[t + 1] + np.asarray(x_eps[t], dt) * s
        return x, x @ self.dec_w.T + self.dec_b

    def weight_kalman(self, xu, target, v, diffusion=0.0):
        """LinearRegression.kalman vjf/module.py:114-142: weight-space Kalman update with A = I, Q = diffusion * I,
        H = features (sample, feature), R = v * I_sample, through kalman.predict (vjf/kalman.py:15-50) and
        kalman.joseph_update AS WRITTEN (:102-145, S^-1 applied twice).  S is (sample, sample) here, like the reference."""
        dt = self.dtype
        assert diffusion >= 0.0, "diffusion needs to be non-negative"
        eye = np.eye(self.n_rbf, dtype=dt)
        H = self.feature(np.asarray(xu, dt))
        R = np.eye(H.shape[0], dtype=dt) * dt.type(v)
        yhat, mhat, Lhat = kalman_predict(self.w_mean, self.w_chol, eye, dt.type(diffusion) * eye, H)
        self.w_mean, self.w_chol = kalman_joseph_update(np.asarray(target, dt), yhat, mhat, Lhat, H, R)

    def fit(self, y, u=None, *, eps, centroid_unit, max_iter=200, beta=0.1, rtol=1e-4):
        """VJF.fit vjf/model.py:223-307: epochs over the sequence; warm-up exit (:278-292: decoder freeze, transition
        re-initialisation from the filtered means), convergence break (:293-296), running loss (:298), per-epoch lr decay
        (:303).  eps: (max_iter, T, 2, B, d) noise tape; centroid_unit(r) -> (R, d+u) centroids drawn U(-r, r).
        Returns mu, logvar, epoch_loss, n_epochs run."""
        dt = self.dtype
        y = np.asarray(y, dt)
        if y.ndim == 2:
            y = y[:, None, :]
        T = y.shape[0]
        warm_up = True
        running = float("nan")
        epoch_loss = float("nan")
        isclose = lambda a, b: bool(np.isfinite(a) and np.isfinite(b) and abs(a - b) <= 1e-8 + rtol * abs(b))  # torch.isclose
        n_epochs = 0
        mu = lv = None
        for i in range(max_iter):
            mu, lv, losses = self.run(y, u, eps=eps[i], warm_up=warm_up)
            n_epochs += 1
            epoch_loss = dt.type(np.sum(losses[:, 0], dtype=dt) / dt.type(T))
            if warm_up:
                if isclose(epoch_loss, running):
                    warm_up = False
                    running = epoch_loss
                    self.decoder_frozen = True
                    u_init = None if (u is None or self.udim == 0) else np.asarray(u, dt)[1:].reshape(-1, self.udim)
                    xt_, xs_ = mu[1:].reshape(-1, self.xdim), mu[:-1].reshape(-1, self.xdim)
                    xu = xs_ if u_init is None else np.concatenate([xs_, u_init], -1)
                    r = float(np.sqrt((xu * xu).sum(1)).max())
                    self.initialize_transition(xt_, xs_, u_init, centroid=centroid_unit(r))
            else:
                if isclose(epoch_loss, running):
                    break
            running = beta * running + (1 - beta) * epoch_loss if i > 0 else epoch_loss
            self.lr *= self.lr_decay
        return mu, lv, epoch_loss, n_epochs


# --------------------------------------------------------------------------------------------
# Kalman operator (vjf/kalman.py, vjf/numerical.py) -- single problem, (n,batch) state layout
# --------------------------------------------------------------------------------------------
def kalman_predict(x, L, A, Q, H):
    """kalman.predict vjf/kalman.py:15-50 with cholesky=True (V passed as its Cholesky factor)."""
    xhat = A @ x
    AL = A @ L
    Vhat = AL @ AL.T + Q
    yhat = H @ xhat
    return yhat, xhat, np.linalg.cholesky(Vhat)


def kalman_update(y, yhat, xhat, Lhat, H, R):
    """kalman.update vjf/kalman.py:53-99 (cholesky=True)."""
    e = y - yhat
    Vhat = Lhat @ Lhat.T
    HL = H @ Lhat
    S = HL @ HL.T + R
    L = np.linalg.cholesky(S)
    G = sla.solve_triangular(L, H @ Vhat, lower=True).T
    x = xhat + G @ sla.solve_triangular(L, e, lower=True)
    V = Vhat - G @ G.T
    return x, np.linalg.cholesky(V)


def kalman_joseph_update(y, yhat, xhat, Lhat, H, R):
    """kalman.joseph_update vjf/kalman.py:102-145 AS WRITTEN: G = (S^-1 H Vhat)^T already is the full
    gain, and S^-1 is applied a second time to e, H and sqrt(R) (:136-140).  sqrt(R) is elementwise."""
    e = y - yhat
    Vhat = Lhat @ Lhat.T
    HL = H @ Lhat
    S = HL @ HL.T + R
    L = np.linalg.cholesky(S)
    cs = lambda b: sla.cho_solve((L, True), b)
    G = cs(H @ Vhat).T
    x = xhat + G @ cs(e)
    ImKH = np.eye(Vhat.shape[0], dtype=Vhat.dtype) - G @ cs(H)
    IL = ImKH @ Lhat
    KR = G @ cs(np.sqrt(R))
    V = IL @ IL.T + KR @ KR.T
    return x, np.linalg.cholesky(V)


def positivize(a, eps=1e-3):
    """numerical.positivize vjf/numerical.py:8-14."""
    w, v = np.linalg.eigh(a)
    s = np.sqrt(np.maximum(w, eps))
    sq = v * s[None, :]
    return sq @ sq.T


def symmetrize(a):
    """numerical.symmetrize vjf/numerical.py:17-19 (upper triangle mirrored down)."""
    return np.triu(a) + np.swapaxes(np.triu(a, 1), -1, -2)
