"""Host-side logic of the trial-sharded path, exercised with gloo on CPU (world_size 2).

The kernels need a GPU; what runs here is everything around them: the partition of the trial axis, the
flags of a step, and the exchange pattern -- local sums of every rank are all-reduced (sum) once per step
and every rank must end up with the identical vector (which is what keeps the replicas in lock-step)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vjf_b200 import _lib
from vjf_b200.distributed import plan_step, shard_bounds


def test_shard_bounds_cover_the_trial_axis():
    for n, w in [(4096, 8), (65536, 8), (10, 4), (7, 8), (1, 2)]:
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
            assert b == c and a <= b
        assert sum(b - a for a, b in spans) == n


def test_plan_step_flags():
    f0 = plan_step(0)
    assert f0 & _lib.FLAG_PRIOR_Q0 and f0 & _lib.FLAG_SGD and f0 & _lib.FLAG_UPDATE and not f0 & _lib.FLAG_WARMUP
    f1 = plan_step(3, sgd=False, update=True, warm_up=True, decoder_frozen=True)
    assert not f1 & _lib.FLAG_PRIOR_Q0 and not f1 & _lib.FLAG_SGD and f1 & _lib.FLAG_WARMUP and f1 & _lib.FLAG_DECODER_FROZEN


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_trials, ps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_trials, world, rank)
    rng = np.random.default_rng(0)
    per_trial = torch.as_tensor(rng.normal(size=(n_trials, ps)))          # what each trial contributes to the sums
    local = per_trial[lo:hi].sum(0)                                       # phase A of this rank
    nb = torch.tensor([hi - lo]); dist.all_reduce(nb)                     # global batch (ShardedVJF.run does the same)
    dist.all_reduce(local)                                                # the single exchange of the step
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    q.put((rank, int(nb.item()), local.numpy(), all(torch.equal(g, gathered[0]) for g in gathered)))
    dist.destroy_process_group()


def test_allreduce_of_local_sums_equals_global_sum():
    world, n_trials, ps = 2, 37, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_trials, ps, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    want = np.random.default_rng(0).normal(size=(n_trials, ps)).sum(0)
    for rank, nb, vec, identical in res:
        assert nb == n_trials and identical
        np.testing.assert_allclose(vec, want, rtol=1e-12, atol=1e-12)
