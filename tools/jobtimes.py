"""Per-job timing of the production code path (library built with DMIP_JOBMARKS=1 into gpurun_scratch/lib_jobmarks.so):
for one steady-state SDE step of CTA 0, when the issuer starts / finishes issuing each accumulator job and when row
warp 0 sees the accumulator / finishes its epilogue.   DMIP_LIB=gpurun_scratch/lib_jobmarks.so python tests/jobtimes.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmip import _lib
from dmip.models.diffusion import CDE

S = 60
L = _lib.require_gpu()
L.dmip_debug_set_timeline.argtypes = [C.c_void_p, C.c_int32]
torch.manual_seed(0)
m = CDE(100, 27, [512, 512, 512])
y = torch.randn(27, generator=torch.Generator().manual_seed(1)).cuda()
N = 148 * 128 * 2
m(y, num_samples=N, num_steps=S, seed=1, return_tensor=True)
cap = 4 * 16384
buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
L.dmip_debug_set_timeline(C.c_void_p(buf.data_ptr()), cap)
m(y, num_samples=N, num_steps=S, seed=1, return_tensor=True)
torch.cuda.synchronize()
L.dmip_debug_set_timeline(None, 0)
b = buf.cpu().numpy().reshape(4, -1)
roles = []
for r in range(4):
    n = int(b[r, 0])
    roles.append([((int(v) >> 16) & ((1 << 47) - 1), int(v) & 0xFFFF) for v in b[r, 1:n + 1]])
mma, row0, row12 = roles[1], roles[2], roles[3]
starts = [i for i, (t, c) in enumerate(mma) if c == 0xD00]
per = [mma[starts[i + 1]][0] - mma[starts[i]][0] for i in range(len(starts) - 1)]
print("cycles per pass:", sorted(per)[len(per) // 2], "(median of", len(per), ")")
k = 20
lo, hi = mma[starts[k]][0], mma[starts[k + 1]][0]
names = {0x100: "issue_start", 0x200: "issue_done", 0x300: "acc_seen", 0x400: "epi_done", 0x500: "a0_arrive", 0x600: "acc_wait", 0xD00: "pass"}
ev = [(t - lo, "mma ", c) for t, c in mma if lo <= t <= hi] + [(t - lo, "row0", c) for t, c in row0 if lo <= t <= hi] \
    + [(t - lo, "row12", c) for t, c in row12 if lo <= t <= hi]
for t, who, c in sorted(ev):
    print(f"{t:7d} {who:5s} {names.get(c & 0xF00, hex(c)):12s} job {c & 0xFF}")
