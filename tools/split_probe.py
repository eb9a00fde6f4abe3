import sys, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, numpy as np
from gpu_cases import trained_model, load_golden
from dmip.models.diffusion import CDE
fx = load_golden("sampler_trained_cde_linear")
m = trained_model()
N, S = 65536, 200
hi = m(fx["y"], num_samples=N, num_steps=S, precision="fp32", seed=2024)
for sp in (2, 1):
    m.l0_split = sp
    lo = m(fx["y"], num_samples=N, num_steps=S, precision="bf16", seed=2024)
    print("split", sp, "dmean", float(np.abs(lo.mean(0) - hi.mean(0)).max()), "rstd", float(np.abs(lo.std(0) / hi.std(0) - 1).max()),
          "maxabs", float(np.abs(lo - hi).max()), flush=True)
torch.manual_seed(0)
big = CDE(100, 27, [512, 512, 512]); big.sde.to("cuda")
y = torch.randn(27)
for sp in (2, 1, 2, 1):
    big.l0_split = sp
    big(y, num_samples=1 << 20, num_steps=100, seed=1, return_tensor=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); big(y, num_samples=1 << 20, num_steps=500, seed=1, return_tensor=True); e1.record(); torch.cuda.synchronize()
    print("split", sp, "evals/s %.4g" % ((1 << 20) * 500 / (e0.elapsed_time(e1) * 1e-3)), flush=True)
