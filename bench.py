#!/usr/bin/env python
"""Benchmark of the VJF filter + learning step (BASELINE.json metric: trial-steps/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C2"): Lorenz attractor latent, xdim=3, ydim=200 Poisson, 50 RBF
centers, hidden [64], 4096 parallel trials per GPU, T=256 time steps per bench step.  One bench "step"
= one pass of the time loop of VJF.fit (one epoch: T filter+learning steps, sgd=True, update=True,
warm_up=False) over the synthetic batch = B*T trial-steps.

  value     device-resident inputs, CUDA-event timed, max over ranks
  e2e       same work through the C ABI entry vjf_run_host with PINNED HOST buffers: host->device
            copies of the observations and device->host copies of trajectory + losses inside the timed region
  roofline  algorithmic HBM bytes of the persistent step kernel / its measured duration vs the measured
            copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the numpy oracle (a port of the reference's algorithm) on the host cores, bounded sample

--impl reference times the CPU implementation of the same path (the oracle port; the reference is pure
Python and is not present on the GPU box) on the same workload, bounded sample per step.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(name="C2-lorenz-poisson", ydim=200, xdim=3, udim=0, n_rbf=50, hidden=[64], likelihood="poisson",
          trials_per_gpu=4096, T=256)
# BASELINE.json configs[3]: neural-population scale, 65536 trials sharded over the GPUs (strong scaling), SURVEY.md 8d row C4
C4 = dict(name="C4-population-poisson", ydim=2000, xdim=8, udim=0, n_rbf=64, hidden=[128], likelihood="poisson",
          global_trials=65536, T=16)
# BASELINE.json configs[2]: the tensor-pipe configuration: 1024 RBFs, 16 384 trials, Gaussian observations
C3 = dict(name="C3-rbf1024-gaussian", ydim=500, xdim=10, udim=0, n_rbf=1024, hidden=[128], likelihood="gaussian",
          trials_per_gpu=16384, T=16)
# BASELINE.json configs[4]: long-horizon latency path: 1024 trials, T = 100 000, Gaussian observations
C5 = dict(name="C5-long-horizon-gaussian", ydim=50, xdim=4, udim=0, n_rbf=32, hidden=[32], likelihood="gaussian",
          trials_per_gpu=1024, T=100000)
# the same C2 model in the throughput regime: enough trials per step that the serial part of a step is amortised
C2_THROUGHPUT = dict(trials_per_gpu=65536, T=16)
DATA_DESC = "synthetic (Lorenz-driven Poisson counts, random-init parameters; CPU-seeded, identical for both arms)"
ALGO_BYTES_PER_TRIAL_STEP = lambda c, y_bytes=4: y_bytes * c["ydim"] + 4 * (c["udim"] + 4 * c["xdim"])


# ------------------------------------------------------------------------------------------------
def lorenz_poisson(T, B, D, seed):
    """Synthetic data of SURVEY.md section 8d row C2: Lorenz (sigma=10, rho=28, beta=8/3, RK4 dt=0.01),
    per-trial random initial state, z-scored with the attractor's own moments; C ~ N(0,1)/sqrt(3), b=-1;
    y ~ Poisson(exp(xC+b)).  Generated on the CPU from a seeded generator ONE TIME STEP AT A TIME, so that the first
    T' < T steps of a longer run are the same tensors: the GPU arm (T = 256) and the CPU reference arm (a bounded
    T-prefix) see identical observations."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, 3, generator=g, dtype=torch.float64) * 5 + torch.tensor([0., 0., 25.], dtype=torch.float64)
    C = (torch.randn(3, D, generator=g, dtype=torch.float64) / 3 ** 0.5)

    def f(s):
        return torch.stack((10 * (s[:, 1] - s[:, 0]), s[:, 0] * (28 - s[:, 2]) - s[:, 1], s[:, 0] * s[:, 1] - 8 / 3 * s[:, 2]), -1)

    # fixed z-scoring constants (mean / std of the Lorenz attractor's coordinates) instead of statistics of the generated
    # run, which would depend on T
    mean = torch.tensor([0.0, 0.0, 23.55], dtype=torch.float64)
    std = torch.tensor([7.92, 9.01, 8.62], dtype=torch.float64)
    dt, burn = 0.01, 200
    gy = torch.Generator(device="cpu").manual_seed(seed + 1)
    y = torch.empty(T, B, D, dtype=torch.float32)
    for t in range(burn + T):
        k1 = f(x); k2 = f(x + 0.5 * dt * k1); k3 = f(x + 0.5 * dt * k2); k4 = f(x + dt * k3)
        x = x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        if t >= burn:
            rate = torch.exp(torch.clamp(((x - mean) / std) @ C - 1.0, max=4.0)).float()
            y[t - burn] = torch.poisson(rate, generator=gy)
    return y  # float32 counts (T,B,D), CPU


def bench_state(cfg, seed=1234):
    """Initial parameters of the benchmark model in the reference's state_dict layout, drawn on the CPU with the
    reference's distributions (nn.Linear default init, centroids U(-2,2), vjf/module.py:20) -- loaded into both arms."""
    import math
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    D, d, u, R = cfg["ydim"], cfg["xdim"], cfg["udim"], cfg["n_rbf"]

    def lin(out, inp, bias=True):
        k = 1.0 / math.sqrt(inp)
        w = (torch.rand(out, inp, generator=g) * 2 - 1) * k
        b = (torch.rand(out, generator=g) * 2 - 1) * k if bias else None
        return w, b
    s = {}
    n_in = D + u + 2 * d
    for i, h in enumerate(cfg["hidden"]):
        s[f"recognition.mlp.{2 * i}.weight"], s[f"recognition.mlp.{2 * i}.bias"] = lin(h, n_in)
        n_in = h
    s["recognition.mean.weight"], _ = lin(d, n_in, bias=False)
    s["recognition.logvar.weight"], s["recognition.logvar.bias"] = lin(d, n_in)
    s["decoder.decode.weight"], s["decoder.decode.bias"] = lin(D, d)
    s["transition.velocity.feature.centroid"] = torch.rand(R, d + u, generator=g) * 4 - 2
    return s


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.max_mhz = index, period, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_traffic_per_time_step(kernel):
    """DRAM bytes per time step of the time-loop kernel from the committed ncu --set full capture (C2 shapes, 4096 trials):
    profiles/traffic_r02.json for the tile pipeline, traffic_r01.json for the persistent kernel."""
    for name in ("traffic_r02.json", "traffic_r01.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            if t.get("kernel", "vjf_persistent_kernel") == kernel:
                return float(t["dram_bytes_read_per_time_step"]) + float(t["dram_bytes_write_per_time_step"]), "profiles/" + name
        except Exception:
            pass
    return None, None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_port_rate(cfg, B, T_sample, seed=0, y=None):
    """trial-steps/s of the numpy oracle (reference algorithm port) on a bounded sample (fallback when the reference
    package itself is not available)."""
    from oracle.vjf_oracle import OracleVJF
    rng = np.random.default_rng(seed)
    o = OracleVJF(cfg["ydim"], cfg["xdim"], cfg["udim"], cfg["n_rbf"], cfg["hidden"], cfg["likelihood"], dtype=np.float32)
    if y is None:
        y = rng.poisson(0.5, (T_sample, B, cfg["ydim"])).astype(np.float32)
    eps = rng.normal(size=(T_sample + 1, 2, B, cfg["xdim"])).astype(np.float32)
    o.run(y[:1], None, eps=eps[:1])  # warm the BLAS threads / allocator
    t0 = time.perf_counter()
    o.run(y[:T_sample], None, eps=eps[1:])
    dt = time.perf_counter() - t0
    return B * T_sample / dt, dt


class _Tape:
    """Stands in for the NAME vjf.model.reparametrize (imported by name at vjf/model.py:18): reads N(0,1) draws from a
    seeded tape instead of torch's global RNG; nothing else of the reference is touched."""

    def __init__(self):
        self.eps, self.i = None, 0

    def load(self, eps):
        self.eps, self.i = eps, 0

    def __call__(self, q):
        import torch
        mean, logvar = q
        e = self.eps[self.i]
        self.i += 1
        return mean + e * torch.exp(.5 * logvar)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ReferenceRunner:
    """The UNMODIFIED reference (oracle/_ref/vjf, see oracle/make_ref.py): VJF.filter looped over a T-prefix of the
    benchmark's own observations, from the benchmark's own initial parameters, fp32, all host threads."""

    def __init__(self, cfg, B, T_s, ranks):
        import torch
        from oracle import make_ref
        self.ref_model = make_ref.import_reference()
        self.tape = _Tape()
        self.ref_model.reparametrize = self.tape
        torch.set_num_threads(os.cpu_count())
        torch.set_default_dtype(torch.float32)
        self.cfg, self.B, self.T_s = cfg, B, T_s
        per = B // ranks
        self.y = torch.cat([lorenz_poisson(T_s, per, cfg["ydim"], seed=1000 + r) for r in range(ranks)], 1)
        g = torch.Generator().manual_seed(77)
        self.eps = torch.randn(T_s, 2, B, cfg["xdim"], generator=g)
        self.state = bench_state(cfg)

    def fresh(self):
        import torch
        c = self.cfg
        torch.manual_seed(0)
        m = self.ref_model.VJF.make_model(c["ydim"], c["xdim"], c["udim"], c["n_rbf"], c["hidden"], c["likelihood"])
        missing = m.load_state_dict(self.state, strict=False)
        assert not missing.unexpected_keys, missing
        return m

    def epoch(self):
        """T_s filter+learning steps from the initial state; returns seconds."""
        m = self.fresh()
        q = None
        t0 = time.perf_counter()
        for t in range(self.T_s):
            self.tape.load(self.eps[t])
            q, loss = m.filter(self.y[t], None, q, sgd=True, update=True, warm_up=False)
        dt = time.perf_counter() - t0
        self.last_loss = float(loss.detach())
        return dt

    def describe(self, dt):
        span = (f"first {self.T_s} of the {self.cfg['T']} time steps" if self.T_s <= self.cfg["T"] else
                f"{self.T_s} time steps (the {self.cfg['T']}-step workload continued with the same generator)")
        return (f"{span} x {self.B} trials ({dt:.1f} s), UNMODIFIED reference "
                f"vjf.model.VJF.filter (oracle/_ref), fp32, torch {self._tv()} with {os.cpu_count()} threads on {cpu_model()}; "
                f"same observations and initial parameters as the GPU arm, N(0,1) noise from a seeded tape")

    @staticmethod
    def _tv():
        import torch
        return torch.__version__


def reference_available():
    try:
        from oracle import make_ref
        make_ref.make()
        return make_ref.available() or os.path.isdir(os.path.join(make_ref.REF_SRC, "vjf"))
    except Exception:
        return False


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = cfg["trials_per_gpu"] * args.gpus  # whole-job batch on the host cores
    cores = os.cpu_count()
    # the reference's (B,B) temporaries (vjf/module.py:76) make a step O(B^2): keep the sample bounded as the job grows
    T_s = max(1, args.ref_steps_per_step // (args.gpus * args.gpus))
    if reference_available():
        rr = ReferenceRunner(cfg, B, T_s, args.gpus)
        rr.T_s = 1
        for _ in range(args.warmup):  # warm-up epochs of one time step each (thread pools, allocator)
            rr.epoch()
        rr.T_s = T_s
        t0 = time.perf_counter()
        dts = [rr.epoch() for _ in range(args.steps)]
        wall = time.perf_counter() - t0
        val = B * T_s * args.steps / sum(dts)
        kind, sample = "reference", rr.describe(sum(dts))
    else:
        rng = np.random.default_rng(0)
        y = rng.poisson(0.5, (T_s, B, cfg["ydim"])).astype(np.float32)
        for _ in range(args.warmup):
            cpu_port_rate(cfg, B, 1, y=y)
        t0 = time.perf_counter()
        rates = [cpu_port_rate(cfg, B, T_s, seed=i, y=y)[0] for i in range(args.steps)]
        wall = time.perf_counter() - t0
        val = float(np.mean(rates))
        kind = "port"
        sample = f"oracle/_ref ABSENT -> numpy/OpenBLAS oracle port of vjf/model.py:179-221; {T_s} time steps x {B} trials per bench step"
    out = {"impl": "reference", "metric": "trial-steps/sec", "value": val, "unit": "trial-steps/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": DATA_DESC,
           "config": workload_config(cfg, args.gpus),
           "cpu_baseline": {"value": val, "unit": "trial-steps/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": "trial-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(cfg, n_gpus):
    """Identical for both arms (the reference arm times a bounded T-prefix of this workload and says so in cpu_baseline.sample)."""
    return {"workload": f"{cfg['name']}: xdim={cfg['xdim']} ydim={cfg['ydim']} {cfg['likelihood']} n_rbf={cfg['n_rbf']} "
                        f"hidden={cfg['hidden']} trials={cfg['trials_per_gpu']}/GPU x {n_gpus} GPU, "
                        f"T={cfg['T']} time steps per bench step (sgd=True, update=True, warm_up=False)",
            "trials_per_gpu": cfg["trials_per_gpu"], "global_trials": cfg["trials_per_gpu"] * n_gpus,
            "time_steps_per_step": cfg["T"], "parallelism": f"trial-sharded x{n_gpus}",
            "l2_policy": "inputs larger than L2 (observations of one bench step = %.0f MB/GPU)" %
                         (cfg["trials_per_gpu"] * cfg["T"] * cfg["ydim"] * 4 / 1e6)}



def synthetic_counts_gpu(T, B, D, d, dev, seed):
    """Poisson counts driven by a smooth d-dimensional latent (coupled oscillators), generated on the device: used where the
    CPU Lorenz recipe would take minutes (65536 trials, ydim 2000)."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    t = torch.arange(T, device=dev, dtype=torch.float32)[:, None, None] * 0.05
    ph = torch.rand(1, B, d, device=dev, generator=g) * 6.283185
    x = torch.sin(t * (1 + torch.arange(d, device=dev, dtype=torch.float32)) + ph)
    Cm = torch.randn(d, D, device=dev, generator=g) / d ** 0.5
    return torch.poisson(torch.exp(torch.clamp(x @ Cm - 1.0, max=3.0)), generator=g)


def limit_cycle_gaussian(t0, T, B, D, dev, seed):
    """a 4-D latent made of two limit cycles (periods 63 and 41 steps, per-trial phases, slow amplitude
    modulation), linear-Gaussian observations y = x C + b + 0.1 N(0,1).  Closed form in the time index, so any window
    [t0, t0 + T) of the 100 000-step sequence can be generated on the device without holding the rest."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    ph = torch.rand(1, B, 2, device=dev, generator=g) * 6.283185
    Cm = torch.randn(4, D, device=dev, generator=g) * 0.5
    b = torch.randn(D, device=dev, generator=g)
    t = (t0 + torch.arange(T, device=dev, dtype=torch.float64))[:, None, None]
    w = torch.tensor([6.283185307 / 63.0, 6.283185307 / 41.0], device=dev, dtype=torch.float64)
    th = (t * w + ph.double())
    amp = 1.0 + 0.2 * torch.sin(t * (6.283185307 / 977.0) + ph.double()[..., :1])
    x = torch.cat((amp * torch.cos(th), amp * torch.sin(th)), -1).float()  # (T, B, 4)
    gn = torch.Generator(device=dev).manual_seed(seed + 1 + t0)
    return x @ Cm + b + 0.1 * torch.randn(T, B, D, device=dev, generator=gn)


def rotation_gaussian(T, B, D, d, dev, seed):
    """latent = stable random rotation (pairs of planes, radius pulled towards 1) + 0.1 noise, roughly
    unit variance; y = x C + b + 0.1 N(0,1).  Generated on the device."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    th = torch.rand(d // 2, device=dev, generator=g) * 0.3 + 0.05
    x = torch.randn(B, d, device=dev, generator=g)
    Cm = torch.randn(d, D, device=dev, generator=g) / d ** 0.5
    b = torch.randn(D, device=dev, generator=g) * 0.5
    y = torch.empty(T, B, D, device=dev)
    c, s_ = torch.cos(th), torch.sin(th)
    for t in range(T):
        xe, xo = x[:, 0::2].clone(), x[:, 1::2].clone()
        x[:, 0::2] = c * xe - s_ * xo
        x[:, 1::2] = s_ * xe + c * xo
        r = x.pow(2).mean(-1, keepdim=True).sqrt()
        x = x * (1 + 0.1 * (1 - r)) + 0.1 * torch.randn(B, d, device=dev, generator=g)
        y[t] = x @ Cm + b + 0.1 * torch.randn(B, D, device=dev, generator=g)
    return y


def c3_tensor_flops_per_step(B, R):
    """tcgen05 FLOPs the large-n_rbf path issues per time step: GEMM 1 (phi w_chol, the K chunks past a tile's last column
    skipped: 5/8 of B R^2 MACs at n_rbf = 1024) and GEMM 2 (phi^T phi, 20 of 32 tiles), three tf32 products per MAC."""
    nt = (R + 255) // 256
    k1 = sum(min(R, 256 * (j + 1)) for j in range(nt)) * 256           # K columns x N columns summed over the N tiles
    mt = (R + 127) // 128
    t2 = sum(1 for i in range(mt) for j in range(nt) if i >= 2 * j)
    return 3 * 2 * (B * k1 + t2 * 128 * 256 * B)


def time_runs(fn, reps):
    """Mean CUDA-event time (ms) of fn() over reps calls, after one warm-up call."""
    import torch
    fn(); torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def c4_tensor_flops_per_trial_step(cfg):
    """Tensor-core FLOPs the wide-observation path issues per trial-step: the forward and the weight-gradient GEMM over the
    observation columns, two tf32 products each (spike counts are exact in tf32; the weight / g_pre operand is a hi/lo pair)."""
    return 2 * 2 * 2 * cfg["ydim"] * cfg["hidden"][0]


def tensor_flops_per_trial_step(cfg):
    """Tensor-core FLOPs the tile pipeline issues per trial-step (the C2 path is tensor-bound before it is
    HBM-bound once fp32-grade accuracy is bought with hi/lo operand pairs): layer-1 forward and weight gradient over
    K1 + 1 columns, quadratic form and Gram matrix over the padded RBF width; two products per contraction on exact count
    columns, three elsewhere (counted as three: an upper bound of the work, a lower bound of the implied peak time)."""
    D, d, R, H = cfg["ydim"], cfg["xdim"], cfg["n_rbf"], cfg["hidden"][0]
    K1 = D + cfg["udim"] + 2 * d + 1
    Rk = (R + 7) // 8 * 8
    NQ = (Rk + d + 15) // 16 * 16
    return 3 * 2 * (2 * K1 * H + Rk * NQ + NQ * NQ)


def sharded_parity_check(dev, world, rank, seed=5, cfg=None, Bl=192, T=4):
    """Short sharded run (C2 shapes by default, 192 trials per rank, 4 steps, noise tape) against the fp64 oracle of the WHOLE
    batch on rank 0, and bitwise identity of the replicas: SCALE lines carry the evidence that the multi-GPU path computes the
    same thing (the driver's GPU tests run on one GPU only).  With cfg = C4 the wide-observation path and its push all-reduce."""
    import torch
    import torch.distributed as dist
    from vjf_b200.model import VJF
    from vjf_b200.distributed import ShardedVJF
    cfg = cfg or C2
    D, d = cfg["ydim"], cfg["xdim"]
    Bg = Bl * world
    rng = np.random.default_rng(seed)
    t = np.arange(T)[:, None, None] * 0.05
    ph = rng.uniform(0, 2 * np.pi, (1, Bg, d))
    Cm = rng.normal(size=(d, D)) / np.sqrt(d)
    y = rng.poisson(np.exp(np.clip(np.sin(t * (1 + np.arange(d)) + ph) @ Cm - 1.0, None, 3.0))).astype(np.float32)
    eps = rng.normal(size=(T, 2, Bg, d)).astype(np.float32)
    torch.manual_seed(seed)
    m = VJF.make_model(D, d, 0, cfg["n_rbf"], cfg["hidden"], cfg["likelihood"], lr=1e-3, max_trials=Bl, seed=3, device=dev)
    m.load_full_state(bench_state(cfg, seed=4321))
    st0 = {k: v.detach().cpu().numpy() for k, v in m.full_state().items()}
    runner = ShardedVJF(m).connect()
    lo = rank * Bl
    mu, lv, losses = runner.run(torch.as_tensor(y[:, lo:lo + Bl]).to(dev), eps=torch.as_tensor(np.ascontiguousarray(eps[:, :, lo:lo + Bl])).to(dev))
    torch.cuda.synchronize()
    status = m.status()
    flat = m._flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([int(torch.equal(ref, flat))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    mus = [torch.empty_like(mu) for _ in range(world)]
    dist.all_gather(mus, mu.contiguous())
    out = {"trials_per_rank": Bl, "time_steps": T, "replicas_identical": bool(same.item()), "status_word": int(status),
           "kernel_kind": int(m._lib.vjf_last_launch_kind())}
    if rank == 0:
        from oracle.vjf_oracle import OracleVJF
        o = OracleVJF(D, d, 0, cfg["n_rbf"], cfg["hidden"], cfg["likelihood"], lr=1e-3, dtype=np.float64)
        o.set_state(st0)
        omu, olv, ol = o.run(y.astype(np.float64), None, eps=eps.astype(np.float64))
        got = torch.cat(mus, 1).cpu().numpy()
        want = o.get_state()
        have = {k: v.detach().cpu().numpy() for k, v in m.full_state().items()}
        perr = max(float(np.abs(np.asarray(have[k], np.float64) - want[k]).max() / max(1.0, np.abs(want[k]).max()))
                   for k in want if k in have and k not in ("w_chol", "w_pchol"))
        out.update({"max_err_mu": float(np.abs(got - omu).max()), "max_rel_err_losses": float(np.abs(losses.cpu().numpy() - ol).max() / max(1.0, np.abs(ol).max())),
                    "max_rel_err_parameters": perr, "against": "fp64 oracle of the whole batch (oracle/vjf_oracle.py)"})
    del runner, m
    import gc
    gc.collect()  # the handle (and its peer mappings) goes away before the next model connects
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from vjf_b200 import _lib
    from vjf_b200.model import VJF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, D, d = cfg["trials_per_gpu"], cfg["T"], cfg["ydim"], cfg["xdim"]
    lib = _lib.load()
    parity = sharded_parity_check(dev, world, rank) if world > 1 else None

    torch.manual_seed(1234)
    model = VJF.make_model(D, d, cfg["udim"], cfg["n_rbf"], cfg["hidden"], cfg["likelihood"], max_trials=B, seed=99, device=dev)
    model.load_full_state(bench_state(cfg))  # identical parameters on every rank and in the reference arm
    y_host = lorenz_poisson(T, B, D, seed=1000 + rank).contiguous().pin_memory()
    y_dev = y_host.to(dev)
    mu_h = torch.empty(T, B, d).pin_memory(); lv_h = torch.empty(T, B, d).pin_memory(); ls_h = torch.empty(T, 4).pin_memory()

    # Every bench step is "the first epoch of a fit": it starts from the same initial parameters (a 70 KB
    # device-to-device copy inside the timed region).  The information-form RLS accumulates phi^T phi / v without
    # forgetting (shrink = 1, vjf/model.py:371); in fp32 -- the reference's dtype too -- it degrades after a few
    # 1e6 samples (~1500 time steps at 4096 trials), so an unbounded number of back-to-back epochs would time a
    # different, degenerate regime.
    state0 = model._flat.clone()
    if world > 1:
        from vjf_b200.distributed import ShardedVJF
        runner = ShardedVJF(model).connect()  # per-step all-reduce inside the persistent kernel (NVLink peer memory)

        def step_dev():
            model._flat.copy_(state0)
            return runner.run(y_dev)

        def step_e2e():
            model._flat.copy_(state0)
            runner.run_host(y_host, mu_h, lv_h, ls_h, chunk_steps=args.chunk)
    else:
        def step_dev():
            model._flat.copy_(state0)
            return model.run(y_dev)
        flags = _lib.FLAG_SGD | _lib.FLAG_UPDATE | _lib.FLAG_PRIOR_Q0
        p = lambda t: C.c_void_p(t.data_ptr())

        def step_e2e():
            model._flat.copy_(state0)
            torch.cuda.current_stream().synchronize()
            _lib.check(lib.vjf_run_host(model._h, T, B, p(y_host), _lib.Y_F32, None, None, model.seed, model._step_index, flags,
                                        model.lr, p(mu_h), p(lv_h), p(ls_h), args.chunk))
            model._step_index += T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also brings the clocks up from idle) ----
    for _ in range(max(3, args.warmup)):
        step_dev()
    torch.cuda.synchronize()
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < args.spinup:
        step_dev()
        torch.cuda.synchronize()
    model.status()

    # ---- timed region: device-resident inputs ----
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.vjf_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for a, b in evs:
        a.record(); step_dev(); b.record()
    t1.record()
    barrier()
    launches = lib.vjf_launch_count() - launches0
    total_ms = t0.elapsed_time(t1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    clocks = sampler.stop()
    status = model.status()
    kind = int(lib.vjf_last_launch_kind())

    # ---- timed region: end to end through the C ABI with pinned host buffers ----
    step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - w0

    # ---- the same end-to-end call with the spike counts handed over as uint8 (numpy counts are a valid `fit` input of the
    #      reference too, vjf/model.py:243-245 coerces them); 4x fewer PCIe bytes.  Reported beside the fp32 number. ----
    e2e_u8 = None
    if world == 1:
        y_u8 = y_host.to(torch.uint8)
        if bool((y_u8.to(torch.float32) == y_host).all()):
            y_u8 = y_u8.pin_memory()

            def step_e2e_u8():
                model._flat.copy_(state0)
                torch.cuda.current_stream().synchronize()
                _lib.check(lib.vjf_run_host(model._h, T, B, p(y_u8), _lib.Y_U8, None, None, model.seed, model._step_index, flags,
                                            model.lr, p(mu_h), p(lv_h), p(ls_h), 2 * args.chunk))
                model._step_index += T
            step_e2e_u8()
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            for _ in range(args.steps):
                step_e2e_u8()
            torch.cuda.synchronize()
            e2e_u8 = {"value": B * T * args.steps / (time.perf_counter() - w0), "unit": "trial-steps/s", "h2d_bytes_per_step": int(y_u8.numel()),
                      "d2h_bytes_per_step": int((mu_h.numel() + lv_h.numel() + ls_h.numel()) * 4),
                      "api": "vjf_run_host with y_dtype = VJF_Y_U8 (counts as bytes)", "status_word": int(model.status())}


    # ---- further regimes (reported beside the headline; same kernels, CUDA-event timed, device-resident inputs) ----
    extras = {}
    peak, peak_src = measured_peak_gbs()
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        tf32_peak = float(peaks["bf16_tflops_sustained"]) / 2.0
    except Exception:
        tf32_peak = 1400.0 / 2.0
    if world == 1 and not args.no_extras:
        # (a) throughput regime of the C2 model: 65536 trials per step
        Bt, Tt = C2_THROUGHPUT["trials_per_gpu"], C2_THROUGHPUT["T"]
        mt = VJF.make_model(D, d, cfg["udim"], cfg["n_rbf"], cfg["hidden"], cfg["likelihood"], max_trials=Bt, seed=99, device=dev)
        mt.load_full_state(bench_state(cfg))
        yt = lorenz_poisson(Tt, Bt, D, seed=2000).to(dev)
        st_t = mt._flat.clone()

        def step_t():
            mt._flat.copy_(st_t)
            mt.run(yt)
        ms = time_runs(step_t, 5)
        tps = Bt * Tt / (ms * 1e-3)
        ach = ALGO_BYTES_PER_TRIAL_STEP(cfg) * tps / 1e9
        fl = tensor_flops_per_trial_step(cfg)
        extras["throughput_regime"] = {
            "workload": f"{cfg['name']} shapes, {Bt} trials x {Tt} time steps per launch (Lorenz-driven counts), in-kernel Philox",
            "value": tps, "unit": "trial-steps/s", "us_per_time_step": ms / Tt * 1e3, "kernel_kind": int(lib.vjf_last_launch_kind()),
            "status_word": int(mt.status()),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak},
            "tensor_bound": {"tensor_flops_per_trial_step": fl, "achieved_tflops": fl * tps / 1e12, "peak_tflops": tf32_peak,
                             "frac": fl * tps / 1e12 / tf32_peak,
                             "note": "hi/lo tf32 operand pairs (three products per contraction) make this path tensor-bound before it is "
                                     "HBM-bound: peak = MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32 runs at half the bf16 rate)"}}
        del mt, yt
        # (b) one GPU's share of the C4 split over 8 GPUs (8192 trials, ydim 2000): the per-GPU work of the sharded benchmark
        c4 = C4
        B4, T4 = c4["global_trials"] // 8, c4["T"]
        m4 = VJF.make_model(c4["ydim"], c4["xdim"], 0, c4["n_rbf"], c4["hidden"], c4["likelihood"], max_trials=B4, seed=99, device=dev)
        m4.load_full_state(bench_state(c4))
        y4 = synthetic_counts_gpu(T4, B4, c4["ydim"], c4["xdim"], dev, 31)
        st4 = m4._flat.clone()

        def step_4():
            m4._flat.copy_(st4)
            m4.run(y4)
        ms = time_runs(step_4, 3)
        tps = B4 * T4 / (ms * 1e-3)
        ach = ALGO_BYTES_PER_TRIAL_STEP(c4) * tps / 1e9
        fl4 = c4_tensor_flops_per_trial_step(c4)
        extras["c4_one_gpu_share"] = {"workload": f"{c4['name']}: ydim 2000 xdim 8 n_rbf 64 hidden [128], {B4} trials (1/8 of 65536) x {T4} steps",
                                      "value": tps, "unit": "trial-steps/s", "us_per_time_step": ms / T4 * 1e3,
                                      "kernel_kind": int(lib.vjf_last_launch_kind()), "status_word": int(m4.status()),
                                      "kernel_kind_meaning": "3 = wide-observation launch sequence (csrc/wide.cu: tcgen05 GEMMs over all trials), 0 = general persistent kernel",
                                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                                   "note": "algorithmic bytes (SURVEY 8d); the launch sequence reads the observations three times "
                                                           "(forward GEMM, likelihood stage, weight-gradient GEMM)"},
                                      "tensor": {"tensor_flops_per_trial_step": fl4, "achieved_tflops": fl4 * tps / 1e12, "peak_tflops": tf32_peak,
                                                 "frac": fl4 * tps / 1e12 / tf32_peak}}
        del m4, y4
        # (b2) the whole C4 batch on ONE GPU: 65536 trials per step
        B4f, T4f = c4["global_trials"], 8
        m4 = VJF.make_model(c4["ydim"], c4["xdim"], 0, c4["n_rbf"], c4["hidden"], c4["likelihood"], max_trials=B4f, seed=99, device=dev)
        m4.load_full_state(bench_state(c4))
        y4 = synthetic_counts_gpu(T4f, B4f, c4["ydim"], c4["xdim"], dev, 31)
        st4 = m4._flat.clone()
        ms = time_runs(step_4, 3)
        tps = B4f * T4f / (ms * 1e-3)
        ach = ALGO_BYTES_PER_TRIAL_STEP(c4) * tps / 1e9
        extras["c4_full_batch_one_gpu"] = {"workload": f"{c4['name']}: ydim 2000 xdim 8 n_rbf 64 hidden [128], {B4f} trials x {T4f} steps on one GPU",
                                           "value": tps, "unit": "trial-steps/s", "us_per_time_step": ms / T4f * 1e3,
                                           "kernel_kind": int(lib.vjf_last_launch_kind()), "status_word": int(m4.status()),
                                           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak},
                                           "tensor": {"tensor_flops_per_trial_step": fl4, "achieved_tflops": fl4 * tps / 1e12, "peak_tflops": tf32_peak,
                                                      "frac": fl4 * tps / 1e12 / tf32_peak}}
        del m4, y4
    if world == 1 and not args.no_extras:
        # (c) C5, the long-horizon latency path: 1024 trials, Gaussian observations, consecutive time steps with NO state reset,
        #     double-precision RLS (vjf_set_rls_precision(64)); observations generated window by window on the device
        c5 = C5
        B5, T5, W5 = c5["trials_per_gpu"], args.c5_steps, 10000
        for bits in (64, 32):
            m5 = VJF.make_model(c5["ydim"], c5["xdim"], 0, c5["n_rbf"], c5["hidden"], c5["likelihood"], max_trials=B5, seed=99, device=dev,
                                rls_precision=bits)
            m5.load_full_state(bench_state(c5))
            q5, ms5, st5 = None, 0.0, 0
            for c0 in range(0, T5, W5):
                y5 = limit_cycle_gaussian(c0, min(W5, T5 - c0), B5, c5["ydim"], dev, seed=7)
                a5 = torch.cuda.Event(enable_timing=True); b5 = torch.cuda.Event(enable_timing=True)
                a5.record()
                mu5, lv5, ls5 = m5.run(y5, None, q5)
                b5.record(); torch.cuda.synchronize()
                ms5 += a5.elapsed_time(b5)
                q5 = (mu5[-1].clone(), lv5[-1].clone())
                from vjf_b200.model import Gaussian as _G
                q5 = _G(*q5)
                st5 |= int(m5.status())
            extras["c5_long_horizon" if bits == 64 else "c5_long_horizon_fp32_rls"] = {
                "workload": f"{c5['name']}: xdim 4 ydim 50 gaussian n_rbf 32 hidden [32], {B5} trials, {T5} consecutive time steps, no state reset, "
                            f"in-kernel Philox, RLS in fp{bits}", "us_per_time_step": ms5 / T5 * 1e3, "time_steps_per_s": T5 / (ms5 * 1e-3),
                "value": B5 * T5 / (ms5 * 1e-3), "unit": "trial-steps/s", "status_word": st5, "final_loss": float(ls5[:, 0].mean().item()),
                "kernel_kind": int(lib.vjf_last_launch_kind()), "note": "latency-bound path (SURVEY 8d): report us/step"}
            del m5
    if world == 1 and not args.no_extras:
        # (d) C3, the tensor-pipe configuration: n_rbf = 1024, 16 384 trials -- the large-n_rbf launch sequence (csrc/bigr.cu)
        c3 = C3
        B3, T3 = c3["trials_per_gpu"], c3["T"]
        m3 = VJF.make_model(c3["ydim"], c3["xdim"], 0, c3["n_rbf"], c3["hidden"], c3["likelihood"], max_trials=B3, seed=99, device=dev)
        m3.load_full_state(bench_state(c3))
        y3 = rotation_gaussian(T3, B3, c3["ydim"], c3["xdim"], dev, 17)
        st3 = m3._flat.clone()

        def step_3():
            m3._flat.copy_(st3)
            m3.run(y3)
        ms = time_runs(step_3, 3)
        tps = B3 * T3 / (ms * 1e-3)
        fl = c3_tensor_flops_per_step(B3, c3["n_rbf"])
        extras["c3_tensor_pipe"] = {
            "workload": f"{c3['name']}: xdim 10 ydim 500 gaussian n_rbf 1024 hidden [128], {B3} trials x {T3} time steps, in-kernel Philox",
            "value": tps, "unit": "trial-steps/s", "us_per_time_step": ms / T3 * 1e3, "kernel_kind": int(lib.vjf_last_launch_kind()),
            "status_word": int(m3.status()),
            "tensor_pipe": {"tensor_flops_per_time_step": fl, "achieved_tflops": fl / (ms / T3 * 1e-3) / 1e12, "peak_tflops": tf32_peak,
                            "frac": fl / (ms / T3 * 1e-3) / 1e12 / tf32_peak,
                            "useful_flops_per_time_step": 2 * 2 * B3 * c3["n_rbf"] ** 2 // 2,
                            "note": "whole-step utilisation (GEMMs + everything else of the step) against MEASURED_PEAKS.json bf16_tflops_sustained / 2; "
                                    "three tf32 products per multiply-add (hi/lo operand split) for fp32-grade results"},
            "roofline": {"bound": "tensor", "achieved": fl / (ms / T3 * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": fl / (ms / T3 * 1e-3) / 1e12 / tf32_peak}}
        del m3, y3
    if world > 1 and not args.no_extras:
        # C4 strong scaling: 65536 trials in all, sharded over the ranks (BASELINE.json configs[3])
        from vjf_b200.distributed import ShardedVJF as _S
        c4 = C4
        B4, T4 = c4["global_trials"] // world, c4["T"]
        m4 = VJF.make_model(c4["ydim"], c4["xdim"], 0, c4["n_rbf"], c4["hidden"], c4["likelihood"], max_trials=B4, seed=99, device=dev)
        m4.load_full_state(bench_state(c4))
        y4 = synthetic_counts_gpu(T4, B4, c4["ydim"], c4["xdim"], dev, 31 + rank)
        st4 = m4._flat.clone()
        r4 = _S(m4).connect()

        def step_4():
            m4._flat.copy_(st4)
            r4.run(y4)
        dist.barrier(); torch.cuda.synchronize()
        ms = time_runs(step_4, 3)
        t4 = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        ms = float(t4.item())
        tps = c4["global_trials"] * T4 / (ms * 1e-3)
        ach = ALGO_BYTES_PER_TRIAL_STEP(c4) * tps / world / 1e9
        extras["c4_strong_scaling"] = {"workload": f"{c4['name']}: ydim 2000 xdim 8 n_rbf 64 hidden [128], {c4['global_trials']} trials over {world} GPUs "
                                                   f"({B4} per GPU) x {T4} steps, " + ("wide-observation launch sequence, push all-reduce over NVLink peer memory" if int(lib.vjf_last_launch_kind()) == 3 else "in-kernel NVLink all-reduce"), "scaling": "strong",
                                       "value": tps, "unit": "trial-steps/s", "us_per_time_step": ms / T4 * 1e3,
                                       "kernel_kind": int(lib.vjf_last_launch_kind()), "status_word": int(m4.status()),
                                       "roofline_per_gpu": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}}
        del r4, m4, y4
        extras["c4_strong_scaling"]["parity_check"] = sharded_parity_check(dev, world, rank, seed=6, cfg=c4, Bl=96, T=3)

    if world > 1:
        tt = torch.tensor([total_ms, e2e_s, kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, kern_ms = tt.tolist()

    if rank == 0:
        units = B * world * T * args.steps
        value = units / (total_ms * 1e-3)
        algo_bytes = ALGO_BYTES_PER_TRIAL_STEP(cfg) * B * T  # per launch of the time-loop kernel, per GPU
        kname = "vjf_tile_kernel" if kind == 1 else "vjf_persistent_kernel"
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
        tps, tsrc = measured_traffic_per_time_step(kname) if (cfg["trials_per_gpu"] == C2["trials_per_gpu"] and world == 1) else (None, None)
        out = {"metric": "trial-steps/sec", "value": value, "unit": "trial-steps/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": DATA_DESC,
               "config": workload_config(cfg, world),
               "us_per_time_step": total_ms / args.steps / T * 1e3,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": (tps * T if tps else None), "traffic_source": (f"{tsrc} (ncu dram__bytes_read+write per time step) x T" if tsrc else None),
                            "algorithmic_bytes_per_launch": algo_bytes, "kernel": kname,
                            "algorithmic_bytes_per_trial_step": ALGO_BYTES_PER_TRIAL_STEP(cfg), "peak_source": peak_src,
                            "note": "B=4096 trials/step is latency-bound by the per-step serial chain (grid barriers + RLS factorisation), see DESIGN.md"},
               "e2e": {"value": units / e2e_s, "unit": "trial-steps/s", "h2d_bytes_per_step": int(y_host.numel() * 4),
                       "d2h_bytes_per_step": int((mu_h.numel() + lv_h.numel() + ls_h.numel()) * 4),
                       "api": ("vjf_run_sharded_host" if world > 1 else "vjf_run_host") + " (C ABI, pinned host buffers, chunked H2D overlapped with compute)"},
               "gpu_launches": int(launches), "clocks": clocks, "status_word": int(status),
               "status_word_e2e": int(model.status())}
        if e2e_u8:
            out["e2e_u8"] = e2e_u8
        out["kernel_kind"] = {"value": kind, "meaning": "1 = tile pipeline (tcgen05 on every contraction, TMA tensor loads), 0 = general persistent kernel"}
        if parity is not None:
            out["parity_check"] = parity
        out.update(extras)
        if world == 1 and not args.no_cpu:
            if reference_available():
                rr = ReferenceRunner(cfg, B, args.cpu_steps, 1)
                rr.T_s = 8
                rr.epoch()  # warm-up (thread pools, allocator)
                rr.T_s = args.cpu_steps
                dt = rr.epoch()
                out["cpu_baseline"] = {"value": B * args.cpu_steps / dt, "unit": "trial-steps/s", "cores": os.cpu_count(),
                                       "kind": "reference", "sample": rr.describe(dt)}
            else:
                val, dt = cpu_port_rate(cfg, B, args.cpu_steps)
                out["cpu_baseline"] = {"value": val, "unit": "trial-steps/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"oracle/_ref ABSENT -> {args.cpu_steps} time steps x {B} trials ({dt:.1f} s), numpy/OpenBLAS oracle port of vjf/model.py:179-221"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trials", type=int, default=None, help="trials per GPU (default: the C2 workload, 4096)")
    ap.add_argument("--T", type=int, default=None, help="time steps per bench step (default 256)")
    ap.add_argument("--chunk", type=int, default=16, help="time steps per H2D chunk in the e2e path (the uint8 variant uses 2x)")
    ap.add_argument("--spinup", type=float, default=1.0, help="seconds of extra untimed load so clocks leave idle")
    ap.add_argument("--cpu-steps", type=int, default=768, help="time steps of the CPU baseline sample (~10-20 s of the reference: the T-step workload repeated)")
    ap.add_argument("--ref-steps-per-step", type=int, default=128, help="--impl reference: time steps (a T-prefix of the workload) per bench step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--c5-steps", type=int, default=100000, help="time steps of the C5 long-horizon entry")
    ap.add_argument("--no-extras", action="store_true", help="skip the further regimes (throughput regime, C4 share / strong scaling)")
    args = ap.parse_args()
    cfg = dict(C2)
    if args.trials:
        cfg["trials_per_gpu"] = args.trials
    if args.T:
        cfg["T"] = args.T
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
