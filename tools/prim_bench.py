"""Cycles per iteration of the synchronisation primitives the sampler's producer / issuer loops are made of."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmip import _lib

_lib.require_gpu()
from tools.probe import probe
L = probe.lib()
names = ["try_wait (completed phase)", "elect + arrive", "arrive + wait for it", "clock64", "tcgen05.commit (idle)",
         "commit + wait for it", "tcgen05.fence::after", "3 constant-bank loads", "issuer skeleton (wait+fence+commit)",
         "producer skeleton (wait+expect_tx)"]
iters = 2000
out = torch.zeros(64, dtype=torch.int64, device="cuda")
for _ in range(2):
    _lib.check(L.dmip_debug_prim_bench(iters, out.data_ptr(), None))
    torch.cuda.synchronize()
for n, v in zip(names, out.cpu().tolist()):
    print(f"{n:40s} {v / iters:8.1f} cycles")
r = out.cpu().tolist()
for i, n in enumerate(["try_wait (hardware suspend)", "test_wait spin", "try_wait, 0 ns suspend hint"]):
    print(f"ping-pong round trip, {n:30s} {r[20 + i] / iters:8.1f} cycles")
print("wake-up latency (cycles) after sleeping for D cycles:   try_wait   test_wait spin   try_wait hint 0")
for k in range(6):
    print(f"  D = {250 << k:6d}: {r[24 + 4 * k]:10d} {r[24 + 4 * k + 1]:14d} {r[24 + 4 * k + 2]:16d}")
print("XU / FMA pipe rates with 16 warps (4 per scheduler), lanes per clock per SM:")
for i, n in enumerate(["MUFU.TANH", "MUFU.EX2", "MUFU.RCP", "F2FP.BF16 pack", "FFMA", "TANH + pack mix (1.5 instr)",
                       "mul.wide.u32 (IMAD.WIDE)", "mul.hi.u32", "mul.lo.u32", "cvt.rn.f32.u32 (I2F)",
                       "tanh.approx.f16x2 (2 per instr)", "tanh.approx.bf16x2 (2 per instr)", "cvt.rn.f16x2.f32 pack"]):
    cyc = r[40 + i]
    print(f"  {n:28s} {cyc / (iters * 8):7.2f} cycles per warp-instruction per scheduler x4 warps -> {512 * iters * 8 / cyc:6.1f} lanes/clk/SM")
print("mma.sync (legacy warp-level tensor path), 16 warps x 8 independent accumulator tiles:")
for i, (n, fl) in enumerate([("m16n8k8 tf32", 2 * 16 * 8 * 8), ("m16n8k16 bf16", 2 * 16 * 8 * 16)]):
    cyc = r[53 + i]
    print(f"  {n:16s} {cyc / (iters * 8 * 4):7.2f} cycles per mma per scheduler -> {16 * 8 * iters * fl / cyc:8.0f} FLOP/clk/SM")
