"""Tiny driver for ncu captures of the sampler kernel:  python tests/prof_sampler.py [particles] [sde_steps] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmip.models.diffusion import CDE

N = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 20
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
torch.manual_seed(0)
m = CDE(100, 27, [512, 512, 512])
y = torch.randn(27, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(2):
    out = m(y, num_samples=N, num_steps=S, precision=prec, seed=1234, return_tensor=True)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
out = m(y, num_samples=N, num_steps=S, precision=prec, seed=1234, return_tensor=True)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)
print(f"N={N} S={S} {prec}: {ms:.3f} ms  {N * S / ms * 1e3:.4g} evals/s  finite={bool(torch.isfinite(out).all())}")
