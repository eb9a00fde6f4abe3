"""One K4 call at 1,048,576 rows (after a warm-up call) — the launch profiled under ncu (profiles/r02_ncu_k_surrogate_tc*.txt):
    ncu --set full --clock-control none --import-source on -k regex:k_surrogate -c 2 -o gpurun_out/r02_prof_k4 python tools/k4_prof.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dmip import utils_scatterometry as us
from gpu_cases import _surrogate_module
from util import load_golden

fm, _ = _surrogate_module()
fx = load_golden("scat_energy")
n = int(os.environ.get("K4_ROWS", 1048576))
x = torch.rand(n, 3, device="cuda") * 2 - 1
y = fx["y"][:1].cuda().expand(n, -1).contiguous()
for _ in range(2):
    us.surrogate_call(fm, x, y, 0.2, 0.01, 1000.0)
torch.cuda.synchronize()
