"""Per-phase timeline of CTA 0 of the tcgen05 sampler (debug hook dmip_debug_set_timeline).
python tests/timeline.py [particles] [sde_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ctypes as C
from dmip import _lib
from dmip.models.diffusion import CDE

N = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
S = int(sys.argv[2]) if len(sys.argv) > 2 else 6
L = _lib.require_gpu()
L.dmip_debug_set_timeline.argtypes = [C.c_void_p, C.c_int32]
torch.manual_seed(0)
m = CDE(100, 27, [512, 512, 512])
y = torch.randn(27, generator=torch.Generator().manual_seed(1)).cuda()
m(y, num_samples=N, num_steps=S, seed=1, return_tensor=True)
cap = 8192
buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
L.dmip_debug_set_timeline(C.c_void_p(buf.data_ptr()), cap)
m(y, num_samples=N, num_steps=S, seed=1, return_tensor=True)
torch.cuda.synchronize()
L.dmip_debug_set_timeline(None, 0)
b = buf.cpu().numpy()
n = int(b[0])
ev = sorted(((int(v) >> 16) & ((1 << 47) - 1), int(v) & 0xFFFF) for v in b[1:n + 1])
t0 = ev[0][0]
names = {0xA00: "kb_top", 0xB00: "kb_waited", 0xC00: "kb_probed", 0xD00: "kb_issued", 0x100: "mma_start", 0x200: "mma_issued", 0x300: "epi_accfull", 0x400: "epi_done", 0x500: "a0_arrive"}
print("n events", n)
# print steps 2..3 in detail
a0 = [i for i, (t, c) in enumerate(ev) if c == 0x500]
lo, hi = a0[2], a0[4] if len(a0) > 4 else len(ev)
for t, c in ev[lo:hi]:
    nm = "preadd_done" if c == 0x510 else names.get(c & 0xF00, hex(c))
    print(f"{t - ev[lo][0]:8d}  {nm:12s} job {c & 0xFF}")
per_step = [(ev[a0[i + 1]][0] - ev[a0[i]][0]) for i in range(len(a0) - 1)]
print("cycles per step:", per_step)
fine = [(t, c) for t, c in ev if (c & 0xF00) >= 0xA00]
for t, c in fine:
    print(f"   fine {t - fine[0][0]:7d} {names[c & 0xF00]:10s} kb {c & 0xFF}")
