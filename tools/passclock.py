"""Cycles per pass (SDE step) of CTA 0 with a minimal timeline (pass boundaries only) next to the wall-clock per
tile-step: tells the SM clock the kernel really ran at.  DMIP_DBG must include 64.  python tests/passclock.py [S]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmip import _lib
from dmip.models.diffusion import CDE

S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
L = _lib.require_gpu()
L.dmip_debug_set_timeline.argtypes = [C.c_void_p, C.c_int32]
torch.manual_seed(0)
m = CDE(100, 27, [512, 512, 512])
y = torch.randn(27, generator=torch.Generator().manual_seed(1)).cuda()
N = 148 * 128 * 4
for _ in range(2):
    m(y, num_samples=N, num_steps=S, seed=1, return_tensor=True)
cap = 4 * 16384
buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
L.dmip_debug_set_timeline(C.c_void_p(buf.data_ptr()), cap)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m(y, num_samples=N, num_steps=S, seed=1, return_tensor=True)
e1.record()
torch.cuda.synchronize()
L.dmip_debug_set_timeline(None, 0)
ms = e0.elapsed_time(e1)
b = buf.cpu().numpy().reshape(4, -1)
n = int(b[1, 0])
t = [(int(v) >> 16) & ((1 << 47) - 1) for v in b[1, 1:n + 1] if (int(v) & 0xFFFF) == 0xD00]
d = [t[i + 1] - t[i] for i in range(len(t) - 1)]
d.sort()
med = d[len(d) // 2]
print(f"passes {len(d)}  cycles/pass median {med}  p10 {d[len(d)//10]} p90 {d[9*len(d)//10]}   wall {ms:.2f} ms -> "
      f"{ms * 1e3 / (4 * S):.2f} us per tile-step -> implied clock {med / (ms * 1e3 / (4 * S)) / 1e3:.2f} GHz")
ev = [((int(v) >> 16) & ((1 << 47) - 1), int(v) & 0xFFFF) for v in b[1, 1:n + 1]]
starts = [i for i, (tt, c) in enumerate(ev) if c == 0xD00]
if len(starts) > 12:
    seg = ev[starts[10]:starts[11] + 1]
    print("one pass, issuer: " + " ".join(f"{hex(c)}@{tt - seg[0][0]}" for tt, c in seg))
