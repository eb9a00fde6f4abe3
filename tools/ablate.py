"""Per-step cost of the tcgen05 sampler for compile-time ablation builds (csrc/build.sh with DMIP_EXP=<bits>,
DMIP_OUT=gpurun_scratch/lib_exp<bits>.so); fixed per-launch / per-tile costs are removed by differencing two step
counts.  python tests/ablate.py [bits ...]   (0 = the production library)"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CODE = """
import os, sys, torch
sys.path.insert(0, %r)
from dmip.models.diffusion import CDE
torch.manual_seed(0)
m = CDE(100, 27, [512, 512, 512])
y = torch.randn(27, generator=torch.Generator().manual_seed(1)).cuda()
N = 148 * 128 * 8
res = []
for S in (40, 120):
    for _ in range(2):
        m(y, num_samples=N, num_steps=S, seed=1, return_tensor=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m(y, num_samples=N, num_steps=S, seed=1, return_tensor=True)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1))
per_step_us = (res[1] - res[0]) / 80 / 8 * 1e3
fixed_us = (res[0] - 40 * (res[1] - res[0]) / 80) / 8 * 1e3
print(f"per tile-step {per_step_us:7.2f} us   fixed per tile {fixed_us:8.1f} us   ({res[0]:.2f} ms, {res[1]:.2f} ms)")
""" % ROOT
for d in (sys.argv[1:] or ["0"]):
    env = dict(os.environ)
    if d != "0":
        env["DMIP_LIB"] = os.path.join(ROOT, "gpurun_scratch", f"lib_exp{d}.so")
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, env=env, timeout=300)
    print(f"EXP {d:>3s}: {r.stdout.strip() or r.stderr.strip()[-300:]}")
