#!/usr/bin/env python
"""bench.py — headline benchmark of the dmip-b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[4], the config the metric "score-net evals/sec at 1/2/4/8 B200 + % tensor peak" is
quoted on): synthetic CDE score net, xdim 100, ydim 27 (SURVEY.md §8a †), hidden [512,512,512], weights
torch.manual_seed(0) default init, one observation y ~ N(0,I) (seed 1), --particles (default 1,048,576) per GPU,
--sde-steps (default 1000) Euler–Maruyama steps, Philox seed 1234.  One bench "step" = one full posterior-sampling
call  model(y, num_samples=particles, num_steps=sde_steps)  = particles x sde_steps score-net evaluations.
Weak scaling: every rank integrates its own `particles` (global particle index offset = rank * particles; no
communication on the hot loop).

`value`  — evals/s with y and the samples resident in HBM (CUDA events around the K calls, max over ranks).
`e2e`    — the same metric through the reference-facing call with HOST buffers: y is a pinned host tensor, the
            result is the numpy array the reference returns (device->host copy inside the timed region).
`roofline` — tensor bound: achieved = evals/s x F (F = 1,282,048 algorithmic FLOP per evaluation, unpadded dims,
            tanh/SDE update not counted) vs the measured sustained bf16 cuBLAS peak.
`cpu_baseline` / `--impl reference` — the CPU oracle port of the reference sampler (oracle/sampler.py, which
            follows models/diffusion.py:27-46 op for op) on all host cores, on a bounded sample of the same workload.
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

XDIM, YDIM, HIDDEN = 100, 27, [512, 512, 512]
F_EVAL = 2 * ((XDIM + YDIM + 1) * 512 + 512 * 512 + 512 * 512 + 512 * XDIM)  # 1,282,048


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=1 << 20)
    ap.add_argument("--sde-steps", type=int, default=1000)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- CPU baseline
def synth_params_torch():
    """Same weights as the GPU arm: nn.Linear default init under torch.manual_seed(0), built on CPU."""
    import torch
    torch.manual_seed(0)
    dims = [XDIM + YDIM + 1] + HIDDEN + [XDIM]
    layers = [torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])]
    return [(l.weight.detach().clone(), l.bias.detach().clone()) for l in layers]


def cpu_sampler_rate(n, s, repeats=1):
    """Time the CPU oracle port of BaseClassDiffusionModel.forward on n particles x s steps (all host cores)."""
    import torch
    from oracle import sampler as osamp            # the one place bench.py may execute oracle/
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = synth_params_torch()
    g = torch.Generator().manual_seed(1)
    y = torch.randn(YDIM, generator=g)
    x0 = torch.randn(n, XDIM, generator=g)
    noise = torch.randn(1, n, XDIM, generator=g).expand(s, n, XDIM)   # one draw reused: timing is unaffected
    with torch.no_grad():
        osamp.em_sampler_cde(params, y, x0[:256], noise[:2, :256], 2)       # warm-up
        best = float("inf")
        for _ in range(repeats):
            t0 = time.perf_counter()
            osamp.em_sampler_cde(params, y, x0, noise, s)
            best = min(best, time.perf_counter() - t0)
    return n * s / best, cores, best


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


# ----------------------------------------------------------------------------------------------- arms
def run_reference(args, rank):
    if rank != 0:
        return
    # bounded sample of the same workload: same net, same SDE; ~10-30 s of CPU work per bench step
    n, s = 65536, 100
    for _ in range(min(args.warmup, 1)):
        cpu_sampler_rate(2048, 4)
    rates, secs = [], []
    for _ in range(args.steps):
        r, cores, dt = cpu_sampler_rate(n, s)
        rates.append(r)
        secs.append(dt)
    value = n * s * len(rates) / sum(secs)
    line = {
        "impl": "reference", "metric": "score-net evals/sec (posterior sampler)", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[4]: synthetic CDE xdim={XDIM} ydim={YDIM} hidden=512x3; CPU sample "
                               f"{n} particles x {s} SDE steps per bench step (rate is linear in N*S, SURVEY.md App. B)"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{n} particles x {s} steps, torch CPU fp32, {os.cpu_count()} threads"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world):
    import torch
    import dmip
    from dmip.models.diffusion import CDE

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not dmip.is_available():
        raise RuntimeError("libdmip_sm100.so missing or device is not sm_100 — no fallback")

    torch.manual_seed(0)
    model = CDE(XDIM, YDIM, HIDDEN)          # default nn.Linear init under seed 0 (random-init weights, synthetic)
    model.sde.eval()
    y_host = torch.randn(YDIM, generator=torch.Generator().manual_seed(1)).pin_memory()
    y_dev = y_host.cuda()
    N, S = args.particles, args.sde_steps
    kw = dict(num_samples=N, num_steps=S, precision=args.precision, seed=1234, gidx_base=rank * N)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        model(y_dev, return_tensor=True, **kw)
    barrier()

    # ---- device-resident timing
    clocks = ClockSampler(local)
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    t_wall0 = time.time()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        flush.zero_()                                                     # evict L2 between timed iterations
        out = model(y_dev, return_tensor=True, **kw)
        launches += model.last_launch_count
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop(t_wall0, t_wall1)
    finite = bool(torch.isfinite(out).all().item())

    # ---- end-to-end through the reference-facing call, host buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        x_np = model(y_host, **kw)                                        # H2D of y, D2H of the samples (numpy)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    if dist is not None:
        tmax = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms, e2e_s = tmax.tolist()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    evals = float(N) * S * args.steps * world
    value = evals / (ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained")
    peak_src = "measured sustained bf16 (MEASURED_PEAKS.json)"
    if not peak_tf:
        peak_tf, peak_src = 1400.0, "fallback sustained bf16 (B200_PROFILING.md)"
    ach_tf = value / world * F_EVAL / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_tc_mlp_dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": "score-net evals/sec (posterior sampler)", "value": value, "unit": "evals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"configs[4]: synthetic CDE xdim={XDIM} ydim={YDIM} hidden=512x3, {N} particles/GPU x "
                               f"{S} SDE steps per bench step, Philox noise in-kernel",
                   "samples_per_sec": value / S, "l2": "256 MB buffer zeroed between timed iterations",
                   "finite": finite},
        "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                     "traffic": traffic, "peak_source": peak_src, "flop_per_eval": F_EVAL,
                     "kernel": "k_tc_mlp (one launch per bench step)"},
        "e2e": {"value": evals / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": YDIM * 4,
                "d2h_bytes_per_step": N * XDIM * 4},
        "gpu_launches": launches,
        "clocks": clk,
    }
    if not args.no_cpu_baseline and world == 1:
        n, s = 65536, 100
        r, cores, dt = cpu_sampler_rate(n, s)
        line["cpu_baseline"] = {"value": r, "unit": "evals/s", "cores": cores, "kind": "port",
                                "sample": f"{n} particles x {s} steps of the same net, torch CPU fp32, {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
