"""tcgen05.mma design probe: cycles per instruction for single-CTA (M=128) and CTA-pair (M=256) instructions, A from
shared / tensor memory, with and without: a concurrent bulk-TMA stream into the same shared memory, random (power-
drawing) operand data, per-K-block commits, concurrent tcgen05.ld/st readers (an epilogue stand-in)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmip import _lib

_lib.require_gpu()
from tools.probe import probe
L = probe.lib()
src = torch.zeros((8 << 20) + (64 << 10), dtype=torch.uint8, device="cuda")
grid = 148
iters, k = 256, 256


def run(cg, a_tmem, n, alt=0, stream=0, rnd=0, commit=0, readers=0, math=0, kbwait=0):
    buf = torch.zeros(4 * grid, dtype=torch.int64, device="cuda")
    mode = a_tmem | (alt << 1) | (rnd << 2) | (commit << 3) | (readers << 4) | (math << 6) | (kbwait << 8)
    _lib.check(L.dmip_debug_mma_bench2(cg, mode, n, k, iters, stream, grid, src.data_ptr(), buf.data_ptr(), None))
    torch.cuda.synchronize()
    c = buf.cpu().view(grid, 4).double()
    lead = c[::cg]
    n_mma = iters * k // 16
    cyc = lead[:, 1].mean().item() / n_mma
    flop = 2 * 128 * cg * n * 16 / cyc / cg
    return cyc, flop


print(f"{'cfg':64s} {'cyc/MMA':>8s} {'FLOP/clk/SM':>12s}")
full = len(sys.argv) > 1 and sys.argv[1] == "full"
for cg in ((1, 2) if full else (1,)):
    for a_tmem in (0, 1):
        for n in (128, 256):
            for (stream, rnd, commit, readers, math, kbwait) in [(0, 0, 0, 0, 0, 0), (0, 0, 1, 0, 0, 1), (0, 0, 1, 0, 0, 2),
                                                                 (0, 0, 1, 0, 0, 3), (0, 0, 1, 0, 0, 4), (16384, 1, 1, 2, 3, 1)]:
                cyc, flop = run(cg, a_tmem, n, 0, stream, rnd, commit, readers, math, kbwait)
                print(f"cg{cg} A-{'tmem' if a_tmem else 'smem'} N={n:3d} stream={stream:5d} rnd={rnd} commit={commit} readers={readers} "
                      f"math={math} kbwait={kbwait} {cyc:8.1f} {flop:12.0f}")
