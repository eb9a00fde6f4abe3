"""Layer-0 operand modes of the tcgen05 sampler: accuracy against the fp32 kernel / the oracle and throughput.
l0_split 1: bf16 x | 2: bf16 hi + lo (default) | 3: + W0 split | 4: one f16 part."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import gpu_cases as gc
from util import load_golden

fx = load_golden("sampler_trained_cde_linear")
m = gc.trained_model()
N, S, seed = 65536, 200, 2024
hi = m(fx["y"], num_samples=N, num_steps=S, precision="fp32", seed=seed)
for split in (1, 2, 3, 4):
    m.l0_split = split
    lo = m(fx["y"], num_samples=N, num_steps=S, precision="bf16", seed=seed)
    d = np.abs(lo - hi)
    print(f"trained linear CDE N={N} S={S}: l0_split={split}  max|dx| {d.max():.3e}  mean|dx| {d.mean():.3e}  "
          f"mean shift {np.abs(lo.mean(0) - hi.mean(0)).max():.2e}  std ratio-1 {np.abs(lo.std(0) / hi.std(0) - 1).max():.2e}", flush=True)
for name, kind in (("sampler_cde_synth", "CDE"), ("sampler_cde_scat", "CDE"), ("sampler_cdiffe_scat", "CDiffE"), ("sampler_dps_scat", "Posterior")):
    for split in (1, 2, 4):
        err, tol, _ = gc.case_sampler(name, kind, "bf16", split)
        print(f"{name}: l0_split={split} rel err {err / tol * 2e-3:.3e} (tol 2e-3)", flush=True)
# throughput on the headline workload
from dmip.models.diffusion import CDE
torch.manual_seed(0)
big = CDE(100, 27, [512, 512, 512])
y = torch.randn(27).cuda()
for split in (2, 4, 1, 2, 4):
    big.l0_split = split
    for _ in range(2):
        big(y, num_samples=1 << 20, num_steps=200, seed=1, return_tensor=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        big(y, num_samples=1 << 20, num_steps=200, seed=1, return_tensor=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"synthetic 1M x 200: l0_split={split}  {ms:.1f} ms  {(1 << 20) * 200 / ms / 1e6:.1f} M evals/s... = {(1 << 20) * 200 / (ms * 1e-3):.4e} evals/s", flush=True)
