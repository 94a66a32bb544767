"""Shared helpers for the parity tests (oracle <-> golden fixtures <-> CUDA path)."""
import os

import numpy as np

from oracle.vjf_oracle import Gaussian, OracleVJF

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def oracle_from_golden(g, state_prefix="init.", dtype=None, lr=None):
    ydim, xdim, udim, n_rbf, B, T = [int(v) for v in g["cfg"]]
    dtype = dtype or g["y"].dtype
    m = OracleVJF(ydim, xdim, udim, n_rbf, [int(h) for h in g["hidden"]], str(g["lik"]),
                  lr=float(g["lr"]) if lr is None and "lr" in g else (lr or 1e-4), dtype=dtype)
    m.set_state(sub(g, state_prefix))
    return m


def run_phases(model, g, filter_fn=None):
    """Replay a golden run_case on ``model`` (anything with the oracle's filter signature)."""
    y, eps = g["y"], g["eps"]
    u = g.get("u")
    mu, lv, losses = [], [], []
    q, t = None, 0
    for n, sgd, upd, warm, freeze in g["phases"]:
        if freeze:
            model.decoder_frozen = True
        for _ in range(int(n)):
            q, loss, a, b, c = model.filter(y[t], None if u is None else u[t], q, eps=eps[t], sgd=bool(sgd),
                                            update=bool(upd), verbose=True, warm_up=bool(warm))
            mu.append(np.asarray(q.mean)); lv.append(np.asarray(q.logvar)); losses.append([loss, a, b, c])
            t += 1
    return np.stack(mu), np.stack(lv), np.asarray(losses, dtype=np.float64)


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: max violation at {i}: got {a[i]!r} want {b[i]!r} "
                             f"(|err|={err[i]:.3e}, tol={tol[i]:.3e}); max|err|={err.max():.3e}")


def compare_state(got: dict, want: dict, rtol, atol, skip=()):
    for k, v in want.items():
        if k in skip or k not in got:
            continue
        if k == "w_chol":
            # w_chol is only defined through w_chol w_chol^T = P^-1 (vjf/module.py:102), compare that
            a = np.asarray(got[k], np.float64); b = np.asarray(v, np.float64)
            assert_close(a @ a.T, b @ b.T, rtol * 10, atol * 10, "w_chol w_chol^T")
            continue
        assert_close(got[k], v, rtol, atol, k)
