"""Timeline of CTA 0 of k_tcl_fwd (pair mode) for one steady-state tile pair of a PINN step at batch 65,536: when the
issuer starts / finishes each GEMM and gets each hready, when row warps 0 (local rows), 8 (remote rows) and 15 see the
accumulators, publish their items, finish their stash stores and pass the loss stage.  Needs the timing build:
    DMIP_JOBMARKS=1 DMIP_OUT=$PWD/gpurun_scratch/lib_jobmarks.so bash <pkg>/csrc/build.sh
    DMIP_LIB=$PWD/gpurun_scratch/lib_jobmarks.so python tools/tcl_timeline.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmip import _lib, losses as dl
from dmip.models.diffusion import CDE

L = _lib.require_gpu()
L.dmip_debug_set_timeline.argtypes = [C.c_void_p, C.c_int32]
torch.manual_seed(0)
m = CDE(2, 2, [512, 512, 512])
B = 65536
x, y = torch.randn(B, 2, device="cuda"), torch.randn(B, 2, device="cuda")
t = torch.rand(B, 1, device="cuda") * 0.98 + 0.01
eps = torch.randn(B, 2, device="cuda")
loss_fn = dl.PINNLoss(lambda xx, yy: -xx, lam=0.001, lam2=0.1, pde_loss="FPE", ic_metric="L2", pde_metric="L1")
loss_fn(m.sde, x, y, x, t, eps, None, None)
cap = 4 * 8192
buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
L.dmip_debug_set_timeline(C.c_void_p(buf.data_ptr()), cap)
loss_fn(m.sde, x, y, x, t, eps, None, None)
torch.cuda.synchronize()
L.dmip_debug_set_timeline(None, 0)
b = buf.cpu().numpy().reshape(4, -1)
roles = []
for r in range(4):
    n = int(b[r, 0])
    roles.append([((int(v) >> 16) & ((1 << 47) - 1), int(v) & 0xFFFF) for v in b[r, 1:n + 1]])
mma = roles[0]
starts = [i for i, (tt, c) in enumerate(mma) if c == 0x100]
per = [mma[starts[i + 1]][0] - mma[starts[i]][0] for i in range(len(starts) - 1)]
print("cycles per tile pair:", sorted(per)[len(per) // 2], "(median of", len(per), ")")
k = min(20, len(starts) - 2)
lo, hi = mma[starts[k]][0], mma[starts[k + 1]][0]
def name(c):
    h = c & 0xF00
    if h == 0x100: return f"GEMM {c & 15} start"
    if h == 0x200: return f"GEMM {(c >> 4) & 15} got hready[{c & 15}]"
    if h == 0x300: return f"GEMM {c & 15} issued"
    if h == 0x400: return f"acc_full seen, layer {c & 15}"
    if h == 0x500: return f"layer {(c >> 4) & 15} item {c & 15} published"
    if h == 0x600: return f"layer {(c >> 4) & 15} item {c & 15} stash done"
    return {0x700: "outputs staged", 0x710: "loss stage starts", 0x720: "loss stage done"}.get(c, hex(c))
ev = []
for who, r in zip(("issuer", "row0 ", "row8 ", "row15"), roles):
    ev += [(tt - lo, who, c) for tt, c in r if lo <= tt <= hi]
for tt, who, c in sorted(ev):
    print(f"{tt:7d} {who} {name(c)}")
