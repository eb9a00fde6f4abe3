"""Per-role timeline of CTA 0 of the tcgen05 sampler (debug hook dmip_debug_set_timeline; store-only probes).
python tests/timeline.py [particles] [sde_steps] [model: synth|scat|cdiffe|dps]

Roles: 0 producer (weight stages), 1 MMA issuer, 2 epilogue thread 0 (column half 0), 3 epilogue thread 128 (half 1).
Prints, for the MMA issuer, where every cycle of a steady-state SDE step goes: issuing, or waiting for
(a) a weight stage (ring underflow), (b) the previous layer's epilogue (hready), (c) an accumulator buffer,
(d) the layer-0 operand of the next step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import torch
from dmip import _lib
from dmip.models.diffusion import CDE, CDiffE, PosteriorDiffusionEstimator

N = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kind = sys.argv[3] if len(sys.argv) > 3 else "synth"
L = _lib.require_gpu()
L.dmip_debug_set_timeline.argtypes = [C.c_void_p, C.c_int32]
torch.manual_seed(0)
if kind == "synth":
    m, ydim = CDE(100, 27, [512, 512, 512]), 27
elif kind == "scat":
    m, ydim = CDE(3, 23, [512, 512, 512]), 23
elif kind == "cdiffe":
    m, ydim = CDiffE(3, 23, [512, 512, 512]), 23
else:
    m, ydim = PosteriorDiffusionEstimator(3, 23, [512, 512, 512]), 23
y = torch.randn(ydim, generator=torch.Generator().manual_seed(1)).cuda()
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
    print(f"role {r}: {n} events")
mma, epi = roles[1], roles[2]
t0 = min(ev[0][0] for ev in roles if ev)

# ---- MMA issuer: wait accounting per step (steps delimited by the a0 wait, code 0xD00)
starts = [i for i, (t, c) in enumerate(mma) if c == 0xD00]
pairs = {0x700: 0x800, 0x900: 0xA00, 0xB00: 0xC00, 0xD00: 0xE00}
label = {0x700: "weight-stage underflow", 0x900: "hready (prev-layer epilogue)", 0xB00: "acc buffer busy", 0xD00: "a0 (next step operand)"}
print("\nMMA issuer, per pass: cycles total | waits by cause")
for si in range(len(starts) - 1):
    seg = mma[starts[si]:starts[si + 1] + 1]
    total = seg[-1][0] - seg[0][0]
    waits = {k: 0 for k in pairs}
    open_ = {}
    for t, c in seg[:-1]:
        k = c & 0xF00
        if k in pairs:
            open_[pairs[k]] = (t, k)
        elif k in open_:
            ts, kk = open_.pop(k)
            waits[kk] += t - ts
    w = sum(waits.values())
    print(f"  pass {si}: {total:7d} | " + " | ".join(f"{label[k]} {v:6d}" for k, v in waits.items()) + f" | issuing {total - w:6d}")

# ---- detail of one steady-state pass: per job, MMA issue window and epilogue window
if len(starts) > 3:
    lo, hi = mma[starts[2]][0], mma[starts[3]][0]
    print("\npass 2 detail (cycles from its a0 wait):")
    evs = [(t, "mma", c) for t, c in mma if lo <= t <= hi] + [(t, "epi0", c) for t, c in epi if lo <= t <= hi] + \
          [(t, "epi1", c) for t, c in roles[3] if lo <= t <= hi]
    names = {0x100: "mma_start", 0x200: "mma_issued", 0x300: "accfull", 0x400: "epi_done", 0x500: "a0_arrive",
             0x700: "stage_wait", 0x800: "stage_ok", 0x900: "hready_wait", 0xA00: "hready_ok", 0xB00: "acc_wait",
             0xC00: "acc_ok", 0xD00: "a0_wait", 0xE00: "a0_ok", 0xF00: "issue_pt"}
    for t, who, c in sorted(evs):
        k = c & 0xF00
        arg = f"job {(c >> 4) & 0xF} kb {c & 0xF}" if k in (0x700, 0x800, 0x900, 0xA00) else f"job {c & 0xFF}"
        print(f"{t - lo:8d} {who:5s} {names.get(k, hex(c)):12s} {arg}")
# ---- producer: stage issue cadence
prod = roles[0]
if len(prod) > 200:
    d = [prod[i + 1][0] - prod[i][0] for i in range(100, len(prod) - 1)]
    d.sort()
    print(f"\nproducer: cycles between stage issues  median {d[len(d)//2]}  p10 {d[len(d)//10]}  p90 {d[9*len(d)//10]}")
# ---- producer vs issuer lag (stages issued by the producer before each job starts)
if len(starts) > 3 and prod:
    lo, hi = mma[starts[2]][0], mma[starts[3]][0]
    pj = [(t - lo, c & 0xFF) for t, c in prod if lo - 3000 <= t <= hi]
    print("\nproducer stage issues (cycle, stage index) around pass 2:")
    print("  " + " ".join(f"{t}:{st}" for t, st in pj[:120]))
