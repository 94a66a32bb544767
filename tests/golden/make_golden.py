"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (the reference is mounted read-only at /root/reference and does
not exist on the GPU box):

    python tests/golden/make_golden.py

The reference draws its reparametrisation noise from torch's global RNG (vjf/util.py:11-13).  To
make runs reproducible across implementations the generator patches the *name* ``reparametrize``
inside ``vjf.model`` (it is imported by name at vjf/model.py:18) so that it reads a pre-generated
noise tape; nothing else in the reference is touched.  Parameters are drawn once by the
reference's own constructors and stored, so the implementations under test copy them instead of
re-drawing.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.environ.get("VJF_REFERENCE", "/root/reference"))
import vjf.model as ref_model  # noqa: E402
from vjf import kalman as ref_kalman  # noqa: E402
from vjf import numerical as ref_numerical  # noqa: E402
from vjf.distribution import Gaussian  # noqa: E402
from vjf.model import VJF  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class Tape:
    def __init__(self):
        self.eps = None
        self.i = 0

    def load(self, eps):  # eps: (2, B, d) for one step
        self.eps, self.i = eps, 0

    def __call__(self, q):
        mean, logvar = q
        e = self.eps[self.i]
        self.i += 1
        return mean + e * torch.exp(.5 * logvar)


TAPE = Tape()
ref_model.reparametrize = TAPE


def get_state(m):
    s = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
    lr = m.transition.velocity
    s["w_mean"] = lr.w_mean.detach().clone().numpy()
    s["w_chol"] = lr.w_chol.detach().clone().numpy()
    s["w_precision"] = lr.w_precision.detach().clone().numpy()
    s["likelihood.n_sample"] = np.asarray(getattr(m.likelihood, "n_sample", 0))
    s["transition.n_sample"] = np.asarray(m.transition.n_sample)
    return s


def run_case(name, *, ydim, xdim, udim, n_rbf, hidden, lik, B, T, dtype, seed, phases, ygen, lr=1e-4):
    """phases: list of (n_steps, dict(sgd, update, warm_up, freeze_decoder))"""
    torch.set_default_dtype(dtype)
    torch.manual_seed(seed)
    m = VJF.make_model(ydim, xdim, udim, n_rbf, hidden, lik, lr=lr)
    init = get_state(m)
    g = torch.Generator().manual_seed(seed + 1)
    y = ygen(g, T, B, ydim).to(dtype)
    u = torch.randn(T, B, udim, generator=g).to(dtype) if udim > 0 else None
    eps = torch.randn(T, 2, B, xdim, generator=g).to(dtype)
    mu, lv, losses = [], [], []
    q = None
    t = 0
    phase_of_step = []
    for pi, (n, ph) in enumerate(phases):
        if ph.get("freeze_decoder"):
            m.decoder.requires_grad_(False)
        for _ in range(n):
            TAPE.load(eps[t])
            q, loss, a, b, c = m.filter(y[t], None if u is None else u[t], q, sgd=ph["sgd"],
                                        update=ph["update"], verbose=True, warm_up=ph["warm_up"])
            mu.append(q.mean.detach().numpy().copy())
            lv.append(q.logvar.detach().numpy().copy())
            losses.append([float(loss), float(a), float(b), float(c)])
            phase_of_step.append(pi)
            t += 1
    out = {"init." + k: v for k, v in init.items()}
    out.update({"final." + k: v for k, v in get_state(m).items()})
    out.update(dict(y=y.numpy(), eps=eps.numpy(), mu=np.stack(mu), logvar=np.stack(lv),
                    losses=np.asarray(losses, dtype=np.float64), phase_of_step=np.asarray(phase_of_step),
                    cfg=np.asarray([ydim, xdim, udim, n_rbf, B, T]), hidden=np.asarray(hidden),
                    lik=np.asarray(lik), lr=np.asarray(lr),
                    phases=np.asarray([[n, ph["sgd"], ph["update"], ph["warm_up"], ph.get("freeze_decoder", False)]
                                       for n, ph in phases], dtype=np.int64)))
    if u is not None:
        out["u"] = u.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", out["losses"][-1])


def grads_case(name, *, ydim, xdim, udim, n_rbf, hidden, lik, B, dtype, seed, warm_up, pre_steps=3):
    """Gradients of the loss as autograd computes them (read before the clip, vjf/model.py:209-210),
    after ``pre_steps`` ordinary steps so that no tensor sits at its initial value."""
    torch.set_default_dtype(dtype)
    torch.manual_seed(seed)
    m = VJF.make_model(ydim, xdim, udim, n_rbf, hidden, lik, lr=1e-2)
    g = torch.Generator().manual_seed(seed + 1)
    T = pre_steps + 1
    y = (torch.poisson(torch.rand(T, B, ydim, generator=g) * 3, generator=g) if lik == "poisson"
         else torch.randn(T, B, ydim, generator=g)).to(dtype)
    u = torch.randn(T, B, udim, generator=g).to(dtype) if udim > 0 else None
    eps = torch.randn(T, 2, B, xdim, generator=g).to(dtype)
    q = None
    for t in range(pre_steps):
        TAPE.load(eps[t])
        q, _ = m.filter(y[t], None if u is None else u[t], q, warm_up=False)
    state = get_state(m)
    t = pre_steps
    TAPE.load(eps[t])
    xs, pt, qt, xt, py = m.forward(y[t], q, None if u is None else u[t])
    loss, a, b, c = m.loss(y[t], xs, pt, qt, xt, py, components=True, warm_up=warm_up)
    m.zero_grad()
    loss.backward()
    out = {"state." + k: v for k, v in state.items()}
    for k, p in m.named_parameters():
        if p.grad is not None and k not in ("mean", "logvar"):
            out["grad." + k] = p.grad.detach().numpy().copy()
    out.update(dict(y=y[t].numpy(), eps=eps[t].numpy(), q_mean=q.mean.detach().numpy(), q_logvar=q.logvar.detach().numpy(),
                    qt_mean=qt.mean.detach().numpy(), qt_logvar=qt.logvar.detach().numpy(),
                    pt_mean=pt.mean.detach().numpy(), pt_logvar=pt.logvar.detach().numpy(),
                    losses=np.asarray([float(loss), float(a), float(b), float(c)]),
                    cfg=np.asarray([ydim, xdim, udim, n_rbf, B, 1]), hidden=np.asarray(hidden),
                    lik=np.asarray(lik), warm_up=np.asarray(warm_up)))
    if u is not None:
        out["u"] = u[t].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


def kalman_case(name, n, m_, batch, dtype, seed):
    torch.set_default_dtype(dtype)
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g).to(dtype)
    x = rn(n, batch)
    L0 = torch.linalg.cholesky(rn(n, n) @ rn(n, n).t() * 0.1 + torch.eye(n))
    A = torch.eye(n) + 0.1 * rn(n, n)
    Qh = rn(n, n) * 0.1
    Q = Qh @ Qh.t() + 0.01 * torch.eye(n)
    H = rn(m_, n)
    R = torch.diag(torch.rand(m_, generator=g).to(dtype) + 0.5)
    y = rn(m_, batch)
    yhat, xhat, Lhat = ref_kalman.predict(x, L0, A, Q, H, R)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        xu, Lu = ref_kalman.update(y, yhat, xhat, Lhat, H, R)
    xj, Lj = ref_kalman.joseph_update(y, yhat, xhat, Lhat, H, R)
    S = rn(n, n)
    sym_in = S
    pos_in = S + S.t()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), L0=L0.numpy(), A=A.numpy(), Q=Q.numpy(),
                        H=H.numpy(), R=R.numpy(), y=y.numpy(), yhat=yhat.numpy(), xhat=xhat.numpy(), Lhat=Lhat.numpy(),
                        x_upd=xu.numpy(), L_upd=Lu.numpy(), x_jos=xj.numpy(), L_jos=Lj.numpy(),
                        sym_in=sym_in.numpy(), sym_out=ref_numerical.symmetrize(sym_in).numpy(),
                        pos_in=pos_in.numpy(), pos_out=ref_numerical.positivize(pos_in).numpy())
    print(name, "ok")


def gauss_y(g, T, B, D):
    # script/example.py:17-33-style noisy limit cycle through a random linear map
    t = torch.arange(T, dtype=torch.float64)[:, None] * 0.0314 + torch.rand(1, B, generator=g, dtype=torch.float64) * 6.28
    x = torch.stack((torch.sin(t), torch.cos(t)), -1) + 0.1 * torch.randn(T, B, 2, generator=g, dtype=torch.float64)
    C = torch.randn(2, D, generator=g, dtype=torch.float64)
    b = torch.randn(D, generator=g, dtype=torch.float64)
    return x @ C + b + 0.1 * torch.randn(T, B, D, generator=g, dtype=torch.float64)


def pois_y(g, T, B, D):
    rate = torch.exp(torch.randn(T, B, D, generator=g, dtype=torch.float64) * 0.5 - 0.5)
    return torch.poisson(rate, generator=g)


if __name__ == "__main__":
    full = [(1, dict(sgd=True, update=True, warm_up=False))]
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        # BASELINE config 1 shapes (script/example.py:41-42): R=100, hidden [20], Gaussian, 1 trial
        run_case(f"c1_gauss_{tag}", ydim=20, xdim=2, udim=0, n_rbf=100, hidden=[20], lik="gaussian", B=1, T=40,
                 dtype=dt, seed=10, ygen=gauss_y, phases=[(40, dict(sgd=True, update=True, warm_up=False))])
        # test/test_model.py:32-44 shapes: Poisson, udim=1, two hidden layers; batch of trials
        run_case(f"pois_u_{tag}", ydim=10, xdim=3, udim=1, n_rbf=10, hidden=[5, 5], lik="poisson", B=16, T=25,
                 dtype=dt, seed=11, ygen=pois_y, lr=1e-2, phases=[(25, dict(sgd=True, update=True, warm_up=False))])
        # warm-up phase, then dynamics phase with the decoder frozen (vjf/model.py:278-283)
        run_case(f"gauss_phases_{tag}", ydim=12, xdim=2, udim=1, n_rbf=8, hidden=[6], lik="gaussian", B=8, T=30,
                 dtype=dt, seed=12, ygen=gauss_y, lr=1e-2,
                 phases=[(10, dict(sgd=True, update=True, warm_up=True)),
                         (12, dict(sgd=True, update=True, warm_up=False, freeze_decoder=True)),
                         (4, dict(sgd=False, update=True, warm_up=False)),
                         (4, dict(sgd=True, update=False, warm_up=False))])
        # reduced BASELINE config 2 (Lorenz shapes: xdim 3, ydim 200, 50 RBFs, hidden [64], Poisson)
        run_case(f"c2_small_{tag}", ydim=200, xdim=3, udim=0, n_rbf=50, hidden=[64], lik="poisson", B=64, T=8,
                 dtype=dt, seed=13, ygen=pois_y, phases=[(8, dict(sgd=True, update=True, warm_up=False))])
        grads_case(f"grads_pois_{tag}", ydim=10, xdim=3, udim=1, n_rbf=10, hidden=[5, 5], lik="poisson", B=16,
                   dtype=dt, seed=21, warm_up=False)
        grads_case(f"grads_gauss_{tag}", ydim=12, xdim=2, udim=0, n_rbf=8, hidden=[6], lik="gaussian", B=8,
                   dtype=dt, seed=22, warm_up=False)
        grads_case(f"grads_gauss_warm_{tag}", ydim=12, xdim=2, udim=0, n_rbf=8, hidden=[6], lik="gaussian", B=8,
                   dtype=dt, seed=23, warm_up=True)
        kalman_case(f"kalman_{tag}", n=4, m_=3, batch=5, dtype=dt, seed=31)
    torch.set_default_dtype(torch.float32)
