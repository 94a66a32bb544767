"""Adapters that let the golden replays in tests/helpers.py drive the CUDA model."""
import numpy as np
import torch

from tests.helpers import sub
from vjf_b200.model import VJF, Gaussian


def model_from_golden(g, prefix="init.", lr=None, max_trials=None):
    ydim, xdim, udim, n_rbf, B, T = [int(v) for v in g["cfg"]]
    lr = float(g["lr"]) if (lr is None and "lr" in g) else (lr or 1e-4)
    m = VJF.make_model(ydim, xdim, udim, n_rbf, [int(h) for h in g["hidden"]], str(g["lik"]), lr=lr,
                       max_trials=max_trials or max(B, 1))
    m.load_full_state(sub(g, prefix))
    return m


def state_np(m):
    return {k: v.detach().cpu().numpy() for k, v in m.full_state().items()}


class CudaAsOracle:
    """Gives vjf_b200.VJF the oracle's filter signature (numpy in/out, eps tape mandatory)."""

    def __init__(self, model):
        self.m = model

    @property
    def decoder_frozen(self):
        return not self.m.decoder.decode.weight.requires_grad

    @decoder_frozen.setter
    def decoder_frozen(self, v):
        self.m.decoder.requires_grad_(not v)

    def filter(self, y, u, q, *, eps, sgd, update, verbose, warm_up):
        qs = None if q is None else Gaussian(torch.as_tensor(q.mean), torch.as_tensor(q.logvar))
        out = self.m.filter(torch.as_tensor(y), None if u is None else torch.as_tensor(u), qs, sgd=sgd, update=update,
                            verbose=True, warm_up=warm_up, eps=torch.as_tensor(eps))
        qt = Gaussian(out[0].mean.cpu().numpy(), out[0].logvar.cpu().numpy())
        return (qt,) + tuple(float(v) for v in out[1:])
