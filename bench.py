#!/usr/bin/env python
"""bench.py — headline benchmark of the dmip-b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

Default workload (BASELINE.json configs[4], the config the metric "score-net evals/sec at 1/2/4/8 B200 + % tensor peak"
is quoted on): synthetic CDE score net, xdim 100, ydim 27 (SURVEY.md §8a †), hidden [512,512,512], weights
torch.manual_seed(0) default init, one observation y ~ N(0,I) (seed 1), --particles (default 1,048,576) per GPU,
--sde-steps (default 1000) Euler–Maruyama steps, Philox seed 1234.  One bench "step" = one full posterior-sampling
call  model(y, num_samples=particles, num_steps=sde_steps)  = particles x sde_steps score-net evaluations = ONE launch
of the persistent kernel k_tc_mlp.  Weak scaling: every rank integrates its own `particles` (global particle index
offset = rank * particles; no communication on the hot loop).

`value`    — evals/s with y and the samples resident in HBM (CUDA events around the K calls, max over ranks).
`e2e`      — the same metric through the reference-facing call with HOST buffers: y is a pinned host tensor, the result
             is the numpy array the reference returns (device->host copy inside the timed region).
`roofline` — tensor bound: achieved = evals/s x F (F = algorithmic FLOP per evaluation, unpadded dims, tanh / SDE
             update not counted) vs the measured SUSTAINED bf16 cuBLAS peak (the kernel runs > 1 s).
`cpu_baseline` / `--impl reference` — the CPU oracle port of the reference sampler (oracle/sampler.py, which follows
             models/diffusion.py:27-46 op for op) on all host cores, on a bounded sample of the same workload.

Other workloads (not the headline line; same JSON shape): cdiffe_scat = configs[2] (scatterometry CDiffE, 1M x 1000),
dps_scat = configs[3] per-GPU share (32 observations x 65,536 particles), pinn_linear = configs[1] (linear CDE,
PINNLoss fwd+bwd+Adam at batch 65,536; metric samples/s).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN = [512, 512, 512]
WORKLOADS = {
    # name: (model class, xdim, ydim, n_obs per GPU, default particles per observation, default SDE steps)
    "synthetic": ("CDE", 100, 27, 1, 1 << 20, 1000),
    "cdiffe_scat": ("CDiffE", 3, 23, 1, 1 << 20, 1000),
    "dps_scat": ("Posterior", 3, 23, 32, 1 << 16, 1000),
}


def flop_per_eval(kind, xdim, ydim):
    """Algorithmic FLOPs of one score evaluation (2 (in 512 + 512 512 + 512 512 + 512 out))."""
    def net(i, o):
        return 2 * (i * 512 + 512 * 512 + 512 * 512 + 512 * o)
    if kind == "CDE":
        return net(xdim + ydim + 1, xdim)
    if kind == "CDiffE":
        return net(xdim + ydim + 1, xdim + ydim)
    return net(xdim + 1, xdim) + net(xdim + ydim + 1, xdim)      # DPS: prior + likelihood net


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=0, help="particles per observation per GPU (0 = workload default)")
    ap.add_argument("--sde-steps", type=int, default=0)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="synthetic", choices=list(WORKLOADS) + ["pinn_linear", "mcmc_scat"],
                    help="synthetic = BASELINE configs[4] (the headline line); the others are configs[2], [3], [1]")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- CPU baseline
def cpu_params(in_dim, out_dim, seed=0):
    """nn.Linear default init under torch.manual_seed(seed), built on CPU (the GPU arm's synthetic weights)."""
    import torch
    torch.manual_seed(seed)
    dims = [in_dim] + HIDDEN + [out_dim]
    layers = [torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])]
    return [(l.weight.detach().clone(), l.bias.detach().clone()) for l in layers]


def cpu_sampler_rate(n, s, xdim=100, ydim=27):
    """Time the CPU oracle port of BaseClassDiffusionModel.forward on n particles x s steps (all host cores)."""
    import torch
    from oracle import sampler as osamp            # the one place bench.py may execute oracle/
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = cpu_params(xdim + ydim + 1, xdim)
    g = torch.Generator().manual_seed(1)
    y = torch.randn(ydim, generator=g)
    x0 = torch.randn(n, xdim, generator=g)
    noise = torch.randn(1, n, xdim, generator=g).expand(s, n, xdim)   # one draw reused: timing is unaffected
    with torch.no_grad():
        osamp.em_sampler_cde(params, y, x0[:256], noise[:2, :256], 2)       # warm-up
        t0 = time.perf_counter()
        osamp.em_sampler_cde(params, y, x0, noise, s)
        dt = time.perf_counter() - t0
    return n * s / dt, cores, dt


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.3 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------------- reference arm
CFG_NAMES = {"synthetic": "configs[4]: synthetic CDE", "cdiffe_scat": "configs[2]: scatterometry CDiffE",
             "dps_scat": "configs[3] (per-GPU share): scatterometry DPS"}


def workload_name(args):
    """`config.workload` of the sampler workloads — the same string on our arm and on the reference arm"""
    kind, xdim, ydim, n_obs, n_def, s_def = WORKLOADS[args.workload]
    N, S = args.particles or n_def, args.sde_steps or s_def
    return (f"{CFG_NAMES[args.workload]} xdim={xdim} ydim={ydim} hidden=512x3, {n_obs} observation(s) x "
            f"{N} particles per GPU x {S} SDE steps per bench step, Philox noise in-kernel")


def run_reference(args, rank):
    if rank != 0:
        return
    # bounded sample of the same workload: same net, same SDE; ~10-30 s of CPU work per bench step
    n, s = 65536, 100
    for _ in range(min(args.warmup, 1)):
        cpu_sampler_rate(2048, 4)
    secs = []
    cores = os.cpu_count()
    for _ in range(args.steps):
        _, cores, dt = cpu_sampler_rate(n, s)
        secs.append(dt)
    value = n * s * len(secs) / sum(secs)
    line = {
        "impl": "reference", "metric": "score-net evals/sec (posterior sampler)", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args) if args.workload in WORKLOADS else args.workload,
                   "sample": f"CPU step = {n} particles x {s} SDE steps of the synthetic CDE net (xdim 100, ydim 27); the "
                             "rate is linear in particles x steps (SURVEY.md App. B)"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{n} particles x {s} steps, torch CPU fp32, {cores} threads"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def setup_dist(world):
    import torch
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return local, dist


def run_sampler(args, rank, world):
    import torch
    import dmip
    from dmip.models import diffusion as dm

    local, dist = setup_dist(world)
    if not dmip.is_available():
        raise RuntimeError("libdmip_sm100.so missing or device is not sm_100 — no fallback")
    kind, xdim, ydim, n_obs, n_def, s_def = WORKLOADS[args.workload]
    N, S = args.particles or n_def, args.sde_steps or s_def
    F = flop_per_eval(kind, xdim, ydim)

    torch.manual_seed(0)
    cls = {"CDE": dm.CDE, "CDiffE": dm.CDiffE, "Posterior": dm.PosteriorDiffusionEstimator}[kind]
    model = cls(xdim, ydim, HIDDEN)          # default nn.Linear init under seed 0 (random-init weights, synthetic)
    model.sde.eval()
    y_host = torch.randn(n_obs, ydim, generator=torch.Generator().manual_seed(1 + rank))
    y_host = (y_host[0] if n_obs == 1 else y_host).contiguous().pin_memory()
    y_dev = y_host.cuda()
    per_rank = n_obs * N                      # particles integrated by one rank per bench step
    kw = dict(num_samples=N, num_steps=S, precision=args.precision, seed=1234, gidx_base=rank * per_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        flush.zero_()            # also warms the fill kernel: its first launch loads a torch module (up to 0.6 s, host side)
        model(y_dev, return_tensor=True, **kw)
    barrier()

    # ---- device-resident timing (CUDA events on the stream the kernel is launched on = torch's current stream)
    clocks = ClockSampler(local)
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    t_wall0 = time.time()
    barrier()
    ev0.record()
    for i in range(args.steps):
        flush.zero_()                                                     # evict L2 between timed iterations
        kev[i][0].record()
        out = model(y_dev, return_tensor=True, **kw)
        kev[i][1].record()
        launches += model.last_launch_count
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps      # k_tc_mlp (+ the weight re-pack check): per launch
    gaps_ms = [ev0.elapsed_time(kev[0][0])] + [kev[i][1].elapsed_time(kev[i + 1][0]) for i in range(args.steps - 1)] + \
        [kev[-1][1].elapsed_time(ev1)]
    clk = clocks.stop(t_wall0, t_wall1)
    finite = bool(torch.isfinite(out).all().item())

    # ---- end-to-end through the reference-facing call, host buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        x_np = model(y_host, **kw)                                        # H2D of y, D2H of the samples (numpy)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert x_np.shape[-1] == xdim

    if dist is not None:
        tmax = torch.tensor([ms, e2e_s, kernel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms, e2e_s, kernel_ms = tmax.tolist()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    evals = float(per_rank) * S * args.steps * world
    value = evals / (ms * 1e-3)
    peaks = load_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained")
    peak_src = "measured sustained bf16 cuBLAS (MEASURED_PEAKS.json): the kernel runs inside a > 1 s step"
    if not peak_tf:
        peak_tf, peak_src = 1400.0, "fallback sustained bf16 (B200_PROFILING.md)"
    ach_tf = (float(per_rank) * S / (kernel_ms * 1e-3)) * F / 1e12      # per GPU, from the kernel's own launch duration
    traffic = None
    try:
        # measured for the default launch of a workload only (ncu capture of that exact command); null otherwise
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_tc_mlp_dram_bytes_per_launch", {})
        traffic = tr.get(args.workload) if (not args.particles and not args.sde_steps) else None
    except Exception:
        pass
    line = {
        "metric": "score-net evals/sec (posterior sampler)", "value": value, "unit": "evals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args),
                   "samples_per_sec": value / S, "l2": "256 MB buffer zeroed between timed iterations",
                   "finite": finite, "gaps_ms": [round(g, 2) for g in gaps_ms]},
        "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                     "traffic": traffic, "peak_source": peak_src, "flop_per_eval": F,
                     "frac_of_burst_peak": (ach_tf / peaks["bf16_tflops"]) if peaks.get("bf16_tflops") else None,
                     "kernel": "k_tc_mlp (one launch per bench step)", "kernel_ms": kernel_ms},
        "e2e": {"value": evals / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": n_obs * ydim * 4,
                "d2h_bytes_per_step": per_rank * xdim * 4},
        "gpu_launches": launches,
        "clocks": clk,
    }
    if not args.no_cpu_baseline and world == 1:
        n, s = 65536, 100
        r, cores, dt = cpu_sampler_rate(n, s)
        line["cpu_baseline"] = {"value": r, "unit": "evals/s", "cores": cores, "kind": "port",
                                "sample": f"{n} particles x {s} steps of the synthetic CDE net, torch CPU fp32, {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_pinn(args, rank, world):
    """BASELINE configs[1]: linear CDE + PINNLoss (Score-FPE, exact divergence) training, batch 65,536 per GPU,
    data parallel (gradient all-reduce).  One bench step = one optimisation step."""
    import torch
    import dmip
    from dmip import distributed as dd, losses as dl
    from dmip.models.diffusion import CDE

    local, dist = setup_dist(world)
    if not dmip.is_available():
        raise RuntimeError("libdmip_sm100.so missing or device is not sm_100 — no fallback")
    B = args.particles or 65536
    torch.manual_seed(0)
    model = CDE(2, 2, HIDDEN)
    opt = torch.optim.Adam(model.sde.a.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(7 + rank)
    x = torch.randn(B, 2, generator=g)
    A = torch.tensor([[1.0, 0.5], [0.0, 1.0]])
    y = x @ A.T + 0.3 * torch.randn(B, 2, generator=g)
    xh, yh = x.pin_memory(), y.pin_memory()
    xd, yd = x.cuda(), y.cuda()
    t = (torch.rand(B, 1, generator=g) * 0.998 + 1e-3).cuda()
    loss_fn = dl.PINNLoss(lambda xx, yy: -xx, lam=0.001, lam2=0.1, pde_loss="FPE", ic_metric="L2", pde_metric="L1")

    def step(xa, ya):
        loss, info = dd.train_step_data_parallel(model, opt, loss_fn, xa, ya, t)
        return loss

    for _ in range(max(args.warmup, 0)):
        step(xd, yd)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step(xd, yd)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = getattr(loss_fn, "last_launch_count", 0) * args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step(xh.cuda(non_blocking=True), yh.cuda(non_blocking=True))
        lv = loss.item()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        tmax = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms, e2e_s = tmax.tolist()
        dist.destroy_process_group()
    if rank != 0:
        return
    F = flop_per_eval("CDE", 2, 2)
    value = B * world * args.steps / (ms * 1e-3)
    ach = value / world * 14 * F / 1e12
    print(json.dumps({
        "metric": "PINNLoss training samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: linear CDE + PINNLoss(FPE exact divergence, L1, ic L2), batch {B} per GPU, Adam",
                   "loss": lv},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": 72.0, "unit": "TFLOP/s", "frac": ach / 72.0, "traffic": None,
                     "peak_source": "nominal fp32 FFMA rate (148 SM x 128 FMA x 2 x 1.9 GHz): the loss kernels are fp32 FFMA",
                     "flop_per_sample": 14 * F, "kernel": "k_jets_fwd + k_jets_bwd + k_wgrad"},
        "e2e": {"value": B * world * args.steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": B * 16,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
    }), flush=True)


def run_mcmc(args, rank, world):
    """Ground-truth generation for the scatterometry KL (generate_scatterometry_ground_truth.py): random-walk Metropolis
    chains on the surrogate posterior, 8 observations x 30,000 chains per GPU (config_scatterometry.yml: n_samples_x),
    METR_STEPS = 1000 per bench step, one kernel launch.  Observations shard over ranks, no collective."""
    import torch
    import dmip
    from dmip import mcmc
    local, dist = setup_dist(world)
    if not dmip.is_available():
        raise RuntimeError("libdmip_sm100.so missing or device is not sm_100 — no fallback")
    torch.manual_seed(0)
    fm = torch.nn.Sequential(torch.nn.Linear(3, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                             torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 23)).cuda()
    for q in fm.parameters():
        q.requires_grad = False
    n_obs, n_per, S = 8, args.particles or 30000, args.sde_steps or 1000
    g = torch.Generator().manual_seed(13 + rank)
    xt = (torch.rand(n_obs, 3, generator=g) * 2 - 1).cuda()
    with torch.no_grad():
        ys = fm(xt)
    ys_h = ys.cpu().pin_memory()
    x0 = (torch.rand(n_obs * n_per, 3, generator=g) * 2 - 1).cuda()

    def step(yv, to_host=False):
        x, _ = mcmc.anneal_to_energy(x0, fm, 0.2, 0.01, yv, 1000.0, S, 0.5, seed=1, gidx_base=rank * n_obs * n_per)
        return x.cpu() if to_host else x

    for _ in range(max(args.warmup, 0)):
        step(ys)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step(ys)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(ys_h.cuda(non_blocking=True), to_host=True)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        tmax = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms, e2e_s = tmax.tolist()
        dist.destroy_process_group()
    if rank != 0:
        return
    F = 2 * (3 * 256 + 256 * 256 * 2 + 256 * 23)
    work = n_obs * n_per * S * world * args.steps
    value = work / (ms * 1e-3)
    ach = value / world * F / 1e12
    print(json.dumps({
        "metric": "Metropolis chain-steps/sec (scatterometry ground truth)", "value": value, "unit": "chain-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"scatterometry Metropolis ground truth: {n_obs} observations x {n_per} chains per GPU x {S} "
                               "steps per bench step, random-init surrogate 3-256-256-256-23, Philox in-kernel"},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": 72.0, "unit": "TFLOP/s", "frac": ach / 72.0, "traffic": None,
                     "peak_source": "nominal fp32 FFMA rate (148 SM x 128 FMA x 2 x 1.9 GHz): k_metropolis is fp32 FFMA",
                     "flop_per_chain_step": F, "kernel": "k_metropolis (one launch per bench step)"},
        "e2e": {"value": work / e2e_s, "unit": "chain-steps/s", "h2d_bytes_per_step": n_obs * 23 * 4,
                "d2h_bytes_per_step": n_obs * n_per * 12},
        "gpu_launches": getattr(mcmc.anneal_to_energy, "last_launch_count", 0) * args.steps,
    }), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    elif args.workload == "mcmc_scat":
        run_mcmc(args, rank, world)
    elif args.workload == "pinn_linear":
        run_pinn(args, rank, world)
    else:
        run_sampler(args, rank, world)


if __name__ == "__main__":
    main()
