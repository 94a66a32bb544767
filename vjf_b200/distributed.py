"""Trials sharded over the GPUs of one node: one process per GPU, parameters and RLS state replicated.

Per time step every rank runs phase A on its own trials (vjf_step_phase_a), the packed vector of local
sums -- SGD gradients, RLS statistics phi^T phi / phi^T dx, loss sums -- is all-reduced once over
NCCL/NVLink (the only exchange of the step), and every rank applies the identical
phase B (clip + SGD, running variances, RLS factorisation) so the replicas stay in lock-step without a
broadcast.  The gradient clip is applied after the reduction, as the reference clips the batch-mean
gradient (vjf/model.py:210).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .model import VJF, Gaussian


_DevBuf = _lib.DevBuf


def shard_bounds(n_trials: int, world: int, rank: int):
    """Contiguous block partition of the trial axis (the last ranks get the smaller blocks)."""
    per = (n_trials + world - 1) // world
    lo = min(n_trials, rank * per)
    return lo, min(n_trials, lo + per)


def plan_step(t: int, sgd=True, update=True, warm_up=False, decoder_frozen=False):
    f = 0
    if sgd:
        f |= _lib.FLAG_SGD
    if update:
        f |= _lib.FLAG_UPDATE
    if warm_up:
        f |= _lib.FLAG_WARMUP
    if decoder_frozen:
        f |= _lib.FLAG_DECODER_FROZEN
    if t == 0:
        f |= _lib.FLAG_PRIOR_Q0
    return f


class ShardedVJF:
    def __init__(self, model: VJF, group=None):
        self.m = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        lib = model._lib
        n = int(lib.vjf_reduce_size(model._h))
        ptr = lib.vjf_reduce_buffer(model._h)
        self.reduce_buf = torch.as_tensor(_DevBuf(ptr, n), device=model.device)
        self.fused = False

    def connect(self):
        """Exchange the CUDA IPC handles of the per-rank exchange buffers so that the persistent kernel can do the
        per-step all-reduce itself over NVLink peer memory (vjf_run_sharded)."""
        if self.world == 1:
            return self
        m, lib = self.m, self.m._lib
        hbuf = C.create_string_buffer(64)
        with torch.cuda.device(m.device):
            _lib.check(lib.vjf_comm_local_handle(m._h, hbuf))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(hbuf.raw), group=self.group)
        blob = C.create_string_buffer(b"".join(handles), 64 * self.world)
        with torch.cuda.device(m.device):
            _lib.check(lib.vjf_comm_connect(m._h, self.rank, self.world, blob))
        dist.barrier(group=self.group)
        self.fused = True
        return self

    @torch.no_grad()
    def run(self, y, u=None, *, sgd=True, update=True, warm_up=False, eps=None, trial_offset=None):
        """y: (T, B_local, ydim) on this rank's device.  Returns mu, logvar (T, B_local, d), losses (T, 4)
        (losses are whole-job values, identical on every rank)."""
        m, lib = self.m, self.m._lib
        y = y.to(m.device) if y.dtype == torch.uint8 else y.to(m.device, torch.float32)
        ydt = _lib.Y_U8 if y.dtype == torch.uint8 else _lib.Y_F32
        T, B, _ = y.shape
        Bg, off0 = self._global_batch(B)
        off = off0 if trial_offset is None else trial_offset
        mu = torch.empty(T, B, m.xdim, device=m.device)
        lv = torch.empty_like(mu)
        losses = torch.empty(T, 4, device=m.device)
        frozen = not m.decoder.decode.weight.requires_grad
        p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        s = m._stream()
        if self.fused and self.world > 1:
            f = plan_step(0, sgd, update, warm_up, frozen)
            with torch.cuda.device(m.device):
                _lib.check(lib.vjf_run_sharded(m._h, T, B, Bg, off, p(y), ydt, p(u), None, None, p(eps), m.seed, m._step_index, f,
                                               m.lr, p(mu), p(lv), p(losses), s))
            m._step_index += T
            self._check_exchange()
            return mu, lv, losses
        for t in range(T):
            f = plan_step(t, sgd, update, warm_up, frozen)
            _lib.check(lib.vjf_step_phase_a(m._h, B, Bg, p(y[t]), ydt, p(None if u is None else u[t]),
                                            p(None if t == 0 else mu[t - 1]), p(None if t == 0 else lv[t - 1]),
                                            p(None if eps is None else eps[t]), m.seed, m._step_index + t, off, f,
                                            p(mu[t]), p(lv[t]), s))
            if self.world > 1:
                dist.all_reduce(self.reduce_buf, group=self.group)
            _lib.check(lib.vjf_step_phase_b(m._h, Bg, f, m.lr, p(losses[t]), s))
        m._step_index += T
        return mu, lv, losses

    def _check_exchange(self):
        """A lost peer turns into VJF_ST_COMM_TIMEOUT (the kernel stops applying updates): raise instead of returning a
        trajectory whose replicas have silently diverged.  Reads the status word without clearing the other bits."""
        st = C.c_uint32(0)
        _lib.check(self.m._lib.vjf_get_status(self.m._h, self.m._stream(), C.byref(st), 0))
        if st.value & _lib.ST_COMM_TIMEOUT:
            raise RuntimeError("vjf_b200: a peer's contribution did not arrive within the time-out (VJF_ST_COMM_TIMEOUT)")

    def _global_batch(self, B):
        """(total trials over all ranks, first global trial index of this rank): an exclusive prefix sum over the ranks'
        block sizes, so that uneven blocks (shard_bounds gives the last ranks smaller ones) keep disjoint Philox trial ids."""
        if self.world == 1:
            return B, 0
        sizes = torch.zeros(self.world, dtype=torch.int64, device=self.m.device)
        sizes[self.rank] = B
        dist.all_reduce(sizes, group=self.group)
        sizes = sizes.tolist()
        return int(sum(sizes)), int(sum(sizes[:self.rank]))

    @torch.no_grad()
    def run_host(self, y_host, mu_host, lv_host, losses_host, *, chunk_steps=16, sgd=True, update=True, warm_up=False):
        """End-to-end sharded run from (pinned) host buffers through the C ABI entry vjf_run_sharded_host: chunked,
        double-buffered H2D of this rank's observations overlapped with the fused compute + exchange kernel."""
        m, lib = self.m, self.m._lib
        assert self.fused and self.world > 1, "connect() first (peer-memory exchange)"
        ydt = _lib.Y_U8 if y_host.dtype == torch.uint8 else _lib.Y_F32
        T, B, _ = y_host.shape
        Bg, off = self._global_batch(B)
        frozen = not m.decoder.decode.weight.requires_grad
        f = plan_step(0, sgd, update, warm_up, frozen)
        p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        torch.cuda.current_stream(m.device).synchronize()
        with torch.cuda.device(m.device):
            _lib.check(lib.vjf_run_sharded_host(m._h, T, B, Bg, off, p(y_host), ydt, None, None, m.seed, m._step_index, f, m.lr,
                                                p(mu_host), p(lv_host), p(losses_host), chunk_steps))
        m._step_index += T
