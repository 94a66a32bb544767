"""Host-side mirror of the reference's public API for the filter/learning path.

``VJF.make_model / filter / fit / forecast`` keep the reference's names, argument meaning, return
conventions and error behaviour (vjf/model.py:309-319, :179-221, :223-307, :321-324); underneath,
every step is the hand-written sm_100a CUDA in ``csrc/`` reached through the C ABI of
``include/vjf_b200.h``.  PyTorch is used for device memory, streams and the state_dict plumbing only.
There is no CPU fallback: constructing a model without a CUDA device or without the compiled library
raises.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import warnings
from collections import namedtuple
from typing import Optional, Sequence

import torch
from torch import Tensor, nn

from . import _lib

Gaussian = namedtuple("Gaussian", ["mean", "logvar"])  # vjf/distribution.py:3


def _ptr(t: Optional[Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


class _Holder(nn.Module):
    """Plain container so that attribute paths and state_dict keys match the reference."""


class _Velocity(_Holder):
    """Stands in for LinearRegression(RBF) (vjf/module.py:37-150): ``velocity(x)`` returns a sample of
    the velocity field like the reference's default ``sampling=True`` (script/example.py:69)."""

    def __init__(self, owner):
        super().__init__()
        object.__setattr__(self, "_owner", owner)

    def forward(self, x, sampling=True):
        return self._owner._velocity(x, sampling)

    def kalman(self, x, target, v, diffusion: float = 0.):
        """LinearRegression.kalman (vjf/module.py:114-142): weight-space Kalman update; x: (sample, xdim + udim) inputs
        of the features, target: (sample, xdim), v: observation-noise variance, Q = diffusion * I."""
        assert diffusion >= 0., 'diffusion needs to be non-negative'  # module.py:127
        return self._owner._weight_kalman(x, target, v, diffusion)


class _Scheduler:
    """ExponentialLR stand-in (vjf/model.py:78, :303)."""

    def __init__(self, model, gamma):
        self.model, self.gamma = model, gamma

    def step(self):
        for g in self.model.optimizer.param_groups:
            g["lr"] *= self.gamma


class _Optimizer:
    def __init__(self, lr):
        self.param_groups = [{"lr": lr}]


class VJF(nn.Module):
    def __init__(self, ydim: int, xdim: int, udim: int, n_rbf: int, hidden_sizes: Sequence[int], likelihood: str = "poisson",
                 *, lr: float = 1e-4, lr_decay: float = .9, device=None, max_trials: int = 65536, seed: int = 0,
                 rls_precision: int = 32):
        """Use VJF.make_model (same advice as the reference, vjf/model.py:53-54)."""
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("vjf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._lib = _lib.load()
        lk = likelihood.lower() if isinstance(likelihood, str) else likelihood
        if lk not in _lib.LIK:
            raise NotImplementedError(f"likelihood {likelihood!r}: only 'poisson' and 'gaussian' exist (vjf/model.py:312-315)")
        self.ydim, self.xdim, self.udim, self.n_rbf = int(ydim), int(xdim), int(udim), int(n_rbf)
        self.hidden_sizes = [int(h) for h in hidden_sizes]
        self.likelihood_name = lk
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.max_trials = int(max_trials)
        self.seed = int(seed)
        self._step_index = 0
        self._cfg = _lib.make_config(self.ydim, self.xdim, self.udim, self.n_rbf, self.hidden_sizes, lk, self.max_trials)
        self._lay = _lib.get_layout(self._cfg)
        with torch.cuda.device(self.device):
            self._flat = torch.zeros(int(self._lay.total), dtype=torch.float32, device=self.device)
            h = C.c_void_p()
            _lib.check(self._lib.vjf_create(C.byref(self._cfg), _ptr(self._flat), C.byref(h)))
            self._h = h
            _lib.check(self._lib.vjf_init_state(self._h, self._stream()))
            if int(rls_precision) != 32:  # 64: double-precision RLS for long horizons (what the reference gets from float64)
                _lib.check(self._lib.vjf_set_rls_precision(self._h, int(rls_precision)))
        self._loss_buf = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._build_modules()
        self._init_random()
        self.optimizer = _Optimizer(float(lr))
        self.scheduler = _Scheduler(self, float(lr_decay))

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _view(self, off, *shape, transpose=False, grad=False):
        n = math.prod(shape) if shape else 1
        v = self._flat[off:off + n]
        if transpose:  # stored input-major [in][out]; expose torch's [out][in]
            v = v.view(shape[1], shape[0]).t()
        else:
            v = v.view(*shape) if shape else v.view(())
        return nn.Parameter(v, requires_grad=grad)

    def _build_modules(self):
        L, d, D, u, R = self._lay, self.xdim, self.ydim, self.udim, self.n_rbf
        # VJF.mean / VJF.logvar: prior, never stepped (vjf/model.py:66-67, :69-77)
        self.mean = self._view(L.prior_mean, d)
        self.logvar = self._view(L.prior_logvar, d)
        lik = _Holder()
        if self.likelihood_name == "gaussian":
            lik.logvar = self._view(L.lik_logvar, grad=True)
        self.likelihood = lik
        tr = _Holder()
        tr.logvar = self._view(L.tr_logvar)
        vel = _Velocity(self)
        feat = _Holder()
        feat.centroid = self._view(L.centroid, R, d + u)
        feat.logwidth = self._view(L.logwidth, R)
        vel.feature = feat
        tr.velocity = vel
        self.transition = tr
        rec = _Holder()
        mlp = nn.ModuleDict()
        n_in = D + u + 2 * d
        for i, hsz in enumerate(self.hidden_sizes):
            lin = _Holder()
            lin.weight = self._view(L.mlp_w[i], hsz, n_in, transpose=True, grad=True)
            lin.bias = self._view(L.mlp_b[i], hsz, grad=True)
            mlp[str(2 * i)] = lin  # Sequential indices of the Linear layers (vjf/recognition.py:20-25)
            n_in = hsz
        rec.mlp = mlp
        rec.mean = _Holder()
        rec.mean.weight = self._view(L.head_m_w, d, n_in, transpose=True, grad=True)
        rec.logvar = _Holder()
        rec.logvar.weight = self._view(L.head_v_w, d, n_in, transpose=True, grad=True)
        rec.logvar.bias = self._view(L.head_v_b, d, grad=True)
        self.recognition = rec
        dec = _Holder()
        dec.decode = _Holder()
        dec.decode.weight = self._view(L.dec_w, D, d, transpose=True, grad=True)
        dec.decode.bias = self._view(L.dec_b, D, grad=True)
        self.decoder = dec

    # RLS state lives beside the Parameters as plain tensors in the reference (vjf/module.py:50-54);
    # here they are views of the same flat buffer, so they checkpoint with ``full_state()``.
    @property
    def w_mean(self):
        return self._flat[self._lay.w_mean:self._lay.w_mean + self.n_rbf * self.xdim].view(self.n_rbf, self.xdim)

    def _rr(self, off):
        return self._flat[off:off + self.n_rbf ** 2].view(self.n_rbf, self.n_rbf)

    @property
    def w_chol(self):
        return self._rr(self._lay.w_chol)

    @property
    def w_precision(self):
        return self._rr(self._lay.w_precision)

    @property
    def w_pchol(self):
        return self._rr(self._lay.w_pchol)

    @torch.no_grad()
    def _init_random(self):
        """Random initial values with the reference's distributions: RBF centroids U(-2,2)
        (vjf/module.py:20); nn.Linear default init for recognition and decoder."""
        def linear_(w, b, fan_in):
            k = 1.0 / math.sqrt(fan_in)
            w.uniform_(-k, k)
            if b is not None:
                b.uniform_(-k, k)
        self.transition.velocity.feature.centroid.uniform_(-2.0, 2.0)
        n_in = self.ydim + self.udim + 2 * self.xdim
        for i, hsz in enumerate(self.hidden_sizes):
            lin = self.recognition.mlp[str(2 * i)]
            linear_(lin.weight, lin.bias, n_in)
            n_in = hsz
        linear_(self.recognition.mean.weight, None, n_in)
        linear_(self.recognition.logvar.weight, self.recognition.logvar.bias, n_in)
        linear_(self.decoder.decode.weight, self.decoder.decode.bias, self.xdim)

    # ------------------------------------------------------------------ state exchange
    def full_state(self) -> dict:
        """state_dict() plus what the reference keeps outside it (w_mean, w_chol, w_precision, counters)."""
        s = {k: v.detach().clone() for k, v in self.state_dict().items()}
        s["w_mean"], s["w_chol"], s["w_precision"] = self.w_mean.clone(), self.w_chol.clone(), self.w_precision.clone()
        s["w_pchol"] = self.w_pchol.clone()
        s["likelihood.n_sample"] = torch.tensor(int(self._flat[self._lay.lik_n].item()))
        s["transition.n_sample"] = torch.tensor(int(self._flat[self._lay.tr_n].item()))
        return s

    @torch.no_grad()
    def load_full_state(self, s: dict):
        own = dict(self.named_parameters())
        for k, v in s.items():
            v = torch.as_tensor(v)
            if k in own:
                own[k].copy_(v.to(self.device, torch.float32))
            elif k in ("w_mean", "w_chol", "w_precision", "w_pchol"):
                getattr(self, k).copy_(v.to(self.device, torch.float32))
            elif k == "likelihood.n_sample":
                self._flat[self._lay.lik_n] = float(v)
            elif k == "transition.n_sample":
                self._flat[self._lay.tr_n] = float(v)
        if "w_pchol" not in s and "w_precision" in s:
            self.w_pchol.copy_(torch.linalg.cholesky(self.w_precision.double()).float())

    def save(self, path):
        """On-disk checkpoint: everything a resumed fit needs -- parameters, the RLS state the reference keeps outside
        state_dict() (vjf/module.py:50-54), the sample counters, the learning rate after the decays so far (vjf/model.py:78, :303),
        the decoder freeze (:283) and the position in the Philox noise stream."""
        torch.save({"format": "vjf_b200/1",
                    "config": dict(ydim=self.ydim, xdim=self.xdim, udim=self.udim, n_rbf=self.n_rbf, hidden_sizes=self.hidden_sizes,
                                   likelihood=self.likelihood_name),
                    "state": {k: v.cpu() for k, v in self.full_state().items()},
                    "lr": self.lr, "lr_decay": self.scheduler.gamma, "step_index": self._step_index, "seed": self.seed,
                    "decoder_frozen": not self.decoder.decode.weight.requires_grad}, path)

    def load(self, path):
        """Restore a checkpoint written by save() into this model (same configuration required)."""
        ck = torch.load(path, map_location="cpu", weights_only=False)
        if ck.get("format") != "vjf_b200/1":
            raise RuntimeError(f"{path}: not a vjf_b200 checkpoint")
        mine = dict(ydim=self.ydim, xdim=self.xdim, udim=self.udim, n_rbf=self.n_rbf, hidden_sizes=self.hidden_sizes,
                    likelihood=self.likelihood_name)
        if ck["config"] != mine:
            raise RuntimeError(f"{path}: checkpoint of {ck['config']}, this model is {mine}")
        self.load_full_state(ck["state"])
        self.optimizer.param_groups[0]["lr"] = float(ck["lr"])
        self.scheduler.gamma = float(ck["lr_decay"])
        self._step_index, self.seed = int(ck["step_index"]), int(ck["seed"])
        self.decoder.requires_grad_(not ck["decoder_frozen"])
        return self

    @classmethod
    def from_checkpoint(cls, path, **kwargs):
        ck = torch.load(path, map_location="cpu", weights_only=False)
        c = ck["config"]
        m = cls(c["ydim"], c["xdim"], c["udim"], c["n_rbf"], c["hidden_sizes"], c["likelihood"], lr=ck["lr"], lr_decay=ck["lr_decay"],
                **kwargs)
        return m.load(path)

    @property
    def n_sample(self):
        return int(self._flat[self._lay.tr_n].item())

    def status(self, clear=True) -> int:
        """Device status word (VJF_ST_* bits); the reference reports the same events through logging /
        warnings / asserts (vjf/model.py:138-145, vjf/module.py:112, vjf/functional.py:60)."""
        out = C.c_uint32(0)
        _lib.check(self._lib.vjf_get_status(self._h, self._stream(), C.byref(out), 1 if clear else 0))
        st = out.value
        if st & _lib.ST_COMM_TIMEOUT:
            # a peer's contribution to the in-kernel all-reduce (sharded run) or the side-stream RLS launch of the wide-observation
            # path did not arrive within the time-out: nothing was applied after that point, the replicas / the state are not valid
            raise RuntimeError("vjf_b200: exchange time-out (VJF_ST_COMM_TIMEOUT): the run is invalid")
        if st & _lib.ST_CHOL_FAILED:
            warnings.warn("RLS failed.")  # vjf/module.py:112
        if st & _lib.ST_MSE_NONFINITE:
            logging.warning("non-finite squared error in gaussian_loss (the reference asserts, functional.py:60)")
        return st

    # ------------------------------------------------------------------ the hot path
    def _flags(self, sgd, update, warm_up, prior):
        f = 0
        if sgd:
            f |= _lib.FLAG_SGD
        if update:
            f |= _lib.FLAG_UPDATE
        if warm_up:
            f |= _lib.FLAG_WARMUP
        if not self.decoder.decode.weight.requires_grad:  # decoder.requires_grad_(False), vjf/model.py:283
            f |= _lib.FLAG_DECODER_FROZEN
        if prior:
            f |= _lib.FLAG_PRIOR_Q0
        return f

    def _coerce(self, a, name):
        if a is None:
            return None
        if isinstance(a, Gaussian):
            raise NotImplementedError  # vjf/model.py:42 / likelihood.py:58-59
        t = torch.as_tensor(a)
        t = t.to(device=self.device, dtype=torch.float32)
        return torch.atleast_2d(t).contiguous()

    @property
    def lr(self):
        return self.optimizer.param_groups[0]["lr"]

    @torch.no_grad()
    def filter(self, y, u=None, qs: Gaussian = None, *, sgd: bool = True, update: bool = True, verbose: bool = False,
               warm_up: bool = False, eps: Optional[Tensor] = None):
        """One filtering + learning step; same contract as the reference's VJF.filter
        (vjf/model.py:179-221).  ``eps`` (extension): a (2, batch, xdim) tape of N(0,1) draws to use instead
        of the in-kernel Philox stream (xs draw first, then xt, as vjf/model.py:112,119)."""
        y = self._coerce(y, "y")  # (batch, dim), model.py:194-195
        if y.shape[-1] != self.ydim:
            raise RuntimeError(f"y has {y.shape[-1]} columns, model ydim is {self.ydim}")
        B = y.shape[0]
        u = self._coerce(u, "u") if (u is not None and self.udim > 0) else None
        if self.udim > 0 and u is None:
            raise RuntimeError("model has udim > 0 but u is None")
        if qs is not None:
            qm = qs.mean.detach().to(self.device, torch.float32).reshape(B, self.xdim).contiguous()
            ql = qs.logvar.detach().to(self.device, torch.float32).reshape(B, self.xdim).contiguous()
        else:
            qm = ql = None
        if eps is not None:
            eps = torch.as_tensor(eps).to(self.device, torch.float32).reshape(2, B, self.xdim).contiguous()
        mean = torch.empty(B, self.xdim, dtype=torch.float32, device=self.device)
        logvar = torch.empty_like(mean)
        loss = torch.empty(4, dtype=torch.float32, device=self.device)
        flags = self._flags(sgd, update, warm_up, qs is None)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vjf_step(self._h, B, _ptr(y), _ptr(u), _ptr(qm), _ptr(ql), _ptr(eps), self.seed,
                                          self._step_index, flags, self.lr, _ptr(mean), _ptr(logvar), _ptr(loss),
                                          self._stream()))
        self._step_index += 1
        qt = Gaussian(mean, logvar)
        if verbose:
            return qt, loss[0], loss[1], loss[2], loss[3]
        return qt, loss[0]

    @torch.no_grad()
    def run(self, y, u=None, q0: Gaussian = None, *, sgd=True, update=True, warm_up=False, eps=None):
        """T steps in one persistent launch (the time loop of fit, vjf/model.py:252-261).
        y: (T, B, ydim) float32 or uint8 device/host tensor.  Returns mu (T,B,d), logvar (T,B,d), losses (T,4)."""
        y = torch.as_tensor(y)
        ydt = _lib.Y_U8 if y.dtype == torch.uint8 else _lib.Y_F32
        y = y.to(self.device) if ydt == _lib.Y_U8 else y.to(self.device, torch.float32)
        if y.ndim == 2:
            y = y[:, None, :]
        y = y.contiguous()
        T, B, _ = y.shape
        if u is not None and self.udim > 0:
            u = torch.as_tensor(u).to(self.device, torch.float32).reshape(T, B, self.udim).contiguous()
        else:
            u = None
        if eps is not None:
            eps = torch.as_tensor(eps).to(self.device, torch.float32).reshape(T, 2, B, self.xdim).contiguous()
        mu = torch.empty(T, B, self.xdim, dtype=torch.float32, device=self.device)
        lv = torch.empty_like(mu)
        losses = torch.empty(T, 4, dtype=torch.float32, device=self.device)
        qm = ql = None
        if q0 is not None:
            qm = q0.mean.to(self.device, torch.float32).reshape(B, self.xdim).contiguous()
            ql = q0.logvar.to(self.device, torch.float32).reshape(B, self.xdim).contiguous()
        flags = self._flags(sgd, update, warm_up, q0 is None)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vjf_run(self._h, T, B, _ptr(y), ydt, _ptr(u), _ptr(qm), _ptr(ql), _ptr(eps), self.seed,
                                         self._step_index, flags, self.lr, _ptr(mu), _ptr(lv), _ptr(losses), self._stream()))
        self._step_index += T
        return mu, lv, losses

    @torch.no_grad()
    def loss_gradients(self, y, u=None, qs: Gaussian = None, *, warm_up: bool = False, eps: Optional[Tensor] = None):
        """Gradients of the step's loss w.r.t. every trainable tensor, BEFORE the clip (what autograd leaves in ``.grad``
        after vjf/model.py:209), keyed like state_dict().  Runs the trial-parallel phase of the split step
        (vjf_step_phase_a: forward + hand-derived backward + slot reduction) and reads the reduced sums; no parameter
        is changed.  Returns (grads, Gaussian q_t)."""
        y = self._coerce(y, "y")
        B = y.shape[0]
        u = self._coerce(u, "u") if (u is not None and self.udim > 0) else None
        qm = ql = None
        if qs is not None:
            qm = qs.mean.detach().to(self.device, torch.float32).reshape(B, self.xdim).contiguous()
            ql = qs.logvar.detach().to(self.device, torch.float32).reshape(B, self.xdim).contiguous()
        if eps is not None:
            eps = torch.as_tensor(eps).to(self.device, torch.float32).reshape(2, B, self.xdim).contiguous()
        mean = torch.empty(B, self.xdim, dtype=torch.float32, device=self.device)
        logvar = torch.empty_like(mean)
        flags = self._flags(True, False, warm_up, qs is None)
        lib = self._lib
        with torch.cuda.device(self.device):
            _lib.check(lib.vjf_step_phase_a(self._h, B, B, _ptr(y), _lib.Y_F32, _ptr(u), _ptr(qm), _ptr(ql), _ptr(eps), self.seed,
                                            self._step_index, 0, flags, _ptr(mean), _ptr(logvar), self._stream()))
            n = int(lib.vjf_reduce_size(self._h))
            red = torch.as_tensor(_lib.DevBuf(lib.vjf_reduce_buffer(self._h), n), device=self.device).clone()
        red = red / B  # the slots hold sums over trials; the loss terms are batch means
        L, d, D = self._lay, self.xdim, self.ydim
        g = {}
        n_in = D + self.udim + 2 * d
        for i, hsz in enumerate(self.hidden_sizes):
            g[f"recognition.mlp.{2 * i}.weight"] = red[L.mlp_w[i]:L.mlp_w[i] + n_in * hsz].view(n_in, hsz).t()
            g[f"recognition.mlp.{2 * i}.bias"] = red[L.mlp_b[i]:L.mlp_b[i] + hsz]
            n_in = hsz
        g["recognition.mean.weight"] = red[L.head_m_w:L.head_m_w + n_in * d].view(n_in, d).t()
        g["recognition.logvar.weight"] = red[L.head_v_w:L.head_v_w + n_in * d].view(n_in, d).t()
        g["recognition.logvar.bias"] = red[L.head_v_b:L.head_v_b + d]
        g["decoder.decode.weight"] = red[L.dec_w:L.dec_w + d * D].view(d, D).t()
        g["decoder.decode.bias"] = red[L.dec_b:L.dec_b + D]
        if self.likelihood_name == "gaussian":
            g["likelihood.logvar"] = red[L.lik_logvar]
        return g, Gaussian(mean, logvar)

    @torch.no_grad()
    def fit(self, y, u=None, *, max_iter: int = 200, beta: float = 0.1, verbose: bool = False, rtol: float = 1e-4,
            progress: bool = True, eps=None, init_centroid=None):
        """Same contract as the reference's VJF.fit (vjf/model.py:223-307): epochs over the sequence with a
        warm-up phase, decoder freeze + RLS re-initialisation when the warm-up loss settles, per-epoch lr
        decay, convergence test on the running loss.  Returns (mu, logvar, epoch_loss).
        Extensions for reproducible tests: ``eps`` (max_iter, T, 2, B, xdim) noise tape instead of the in-kernel Philox stream,
        ``init_centroid`` the centroids RBFDS.initialize would re-draw (vjf/module.py:147)."""
        y = torch.as_tensor(y)
        y = y.to(self.device) if y.dtype == torch.uint8 else y.to(self.device, torch.float32)
        y = torch.atleast_2d(y)
        if y.ndim == 2:  # (T, D): iterating gives (D,) -> (1, D) steps (model.py:253, :195)
            y = y[:, None, :]
        T, B, _ = y.shape
        u_ = None
        if u is not None:
            u_ = torch.atleast_2d(torch.as_tensor(u).to(self.device, torch.float32))
            if u_.ndim == 2:
                u_ = u_[:, None, :]
        warm_up = True
        epoch_loss = torch.tensor(float("nan"))
        running_loss = torch.tensor(float("nan"))
        it = range(max_iter)
        bar = None
        if progress:
            try:
                from tqdm import trange
                bar = trange(max_iter)
                it = bar
            except ImportError:
                pass
        mu = lv = None
        for i in it:
            mu, lv, losses = self.run(y, u_, None, sgd=True, update=True, warm_up=warm_up, eps=None if eps is None else eps[i])
            epoch_loss = losses[:, 0].mean().cpu()
            self.status()
            if warm_up:
                if torch.isclose(epoch_loss, running_loss, rtol=rtol):
                    warm_up = False
                    running_loss = epoch_loss
                    print("\nWarm up stopped.\n")
                    self.decoder.requires_grad_(False)  # freeze decoder after warm up (model.py:283)
                    u_init = u_[1:].reshape(-1, u_.shape[-1]) if (u_ is not None and u_.shape[-1] > 0) else None
                    self.initialize_transition(mu[1:].reshape(-1, self.xdim), mu[:-1].reshape(-1, self.xdim), u_init,
                                               centroid=init_centroid)
            else:
                if torch.isclose(epoch_loss, running_loss, rtol=rtol):
                    print("\nConverged.\n")
                    break
            running_loss = beta * running_loss + (1 - beta) * epoch_loss if i > 0 else epoch_loss
            if bar is not None:
                post = {"Loss": running_loss.item()}
                if verbose:
                    last = losses[-1].cpu()
                    post.update({"Recon": last[1].item(), "Dynamics": last[2].item(), "Entropy": last[3].item()})
                bar.set_postfix(post)
            self.scheduler.step()
        if bar is not None:
            bar.close()
        return mu, lv, epoch_loss

    @classmethod
    def make_model(cls, ydim: int, xdim: int, udim: int, n_rbf: int, hidden_sizes: Sequence[int],
                   likelihood: str = "poisson", *args, **kwargs):
        """vjf/model.py:309-319."""
        return cls(ydim, xdim, udim, n_rbf, hidden_sizes, likelihood, *args, **kwargs)

    # ------------------------------------------------------------------ next-tier rows
    @torch.no_grad()
    def initialize_transition(self, xt, xs, ut=None, *, centroid: Optional[Tensor] = None):
        """RBFDS.initialize + LinearRegression.initialize (vjf/model.py:379-388, vjf/module.py:144-150):
        re-draw the centroids U(-r, r), width r, one RLS pass over the whole trajectory, state noise from
        the residual.  ``centroid`` (extension) injects the re-drawn centroids for reproducible tests."""
        xs = torch.atleast_2d(xs).to(self.device, torch.float32).contiguous()
        xt = torch.atleast_2d(xt).to(self.device, torch.float32).contiguous()
        if ut is not None and self.udim > 0:
            ut = torch.atleast_2d(ut).to(self.device, torch.float32).contiguous()
            xu = torch.cat((xs, ut), -1)
        else:
            ut, xu = None, xs
        r = xu.norm(dim=1).max().item()  # module.py:146
        feat = self.transition.velocity.feature
        if centroid is None:
            feat.centroid.uniform_(-r, r)  # module.py:147
        else:
            feat.centroid.copy_(torch.as_tensor(centroid).to(self.device, torch.float32))
        feat.logwidth.fill_(math.log(r))  # module.py:148
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vjf_rls_initialize(self._h, xs.shape[0], _ptr(xs), _ptr(xt), _ptr(ut), self._stream()))
        return r

    @torch.no_grad()
    def forecast(self, x0, u=None, n_step: int = 1, *, noise: bool = False, w_eps=None, x_eps=None):
        """vjf/model.py:321-324 -> RBFDS.forecast (:342-361): sampled-weight rollout and decoded
        observations.  Returns x (n_step+1, B, xdim), y (n_step+1, B, ydim).  ``w_eps``/``x_eps``
        (extension) inject the N(0,1) draws (module.py:71, model.py:359)."""
        x0 = torch.atleast_2d(torch.as_tensor(x0).to(self.device, torch.float32))
        B = x0.shape[0]
        x = torch.empty(n_step + 1, B, self.xdim, dtype=torch.float32, device=self.device)
        x[0] = x0
        yhat = torch.empty(n_step + 1, B, self.ydim, dtype=torch.float32, device=self.device)
        if u is not None and self.udim > 0:
            u = torch.atleast_2d(torch.as_tensor(u).to(self.device, torch.float32))
            assert u.shape[0] == n_step, "u must have length of n_step if present"  # model.py:354
            u = u.reshape(n_step, -1, self.udim).expand(n_step, B, self.udim).contiguous()
        else:
            u = None
        if w_eps is None:
            w_eps = torch.randn(n_step, self.n_rbf, self.xdim, device=self.device)
        w_eps = torch.as_tensor(w_eps).to(self.device, torch.float32).contiguous()
        if noise:
            if x_eps is None:
                x_eps = torch.randn(n_step, B, self.xdim, device=self.device)
            x_eps = torch.as_tensor(x_eps).to(self.device, torch.float32).contiguous()
        else:
            x_eps = None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vjf_forecast(self._h, n_step, B, _ptr(x), _ptr(yhat), _ptr(u), _ptr(w_eps), _ptr(x_eps),
                                              self._stream()))
        return x, yhat

    @torch.no_grad()
    def _weight_kalman(self, x, target, v, diffusion):
        x = torch.atleast_2d(torch.as_tensor(x).to(self.device, torch.float32))
        target = torch.atleast_2d(torch.as_tensor(target).to(self.device, torch.float32)).contiguous()
        xs = x[:, :self.xdim].contiguous()
        u = x[:, self.xdim:].contiguous() if self.udim > 0 else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vjf_weight_kalman(self._h, xs.shape[0], _ptr(xs), _ptr(target), _ptr(u), float(v), float(diffusion),
                                                   self._stream()))

    @torch.no_grad()
    def _velocity(self, x, sampling=True):
        """transition.velocity(x): one-step velocity (sampled weights by default, module.py:70-73)."""
        x = torch.atleast_2d(torch.as_tensor(x).to(self.device, torch.float32))
        B = x.shape[0]
        xd = x[:, :self.xdim].contiguous()
        u = x[:, self.xdim:].reshape(1, B, self.udim).contiguous() if self.udim > 0 else None
        w_eps = torch.randn(1, self.n_rbf, self.xdim, device=self.device) if sampling else \
            torch.zeros(1, self.n_rbf, self.xdim, device=self.device)
        xs, _ = self.forecast(xd, u, 1, w_eps=w_eps)
        return xs[1] - xs[0]

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.vjf_destroy(self._h)
                self._h = None
        except Exception:
            pass
