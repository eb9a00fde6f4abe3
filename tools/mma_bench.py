"""tcgen05.mma throughput probe (M=128): cycles per instruction for N in {64,128,256}, A from smem / tmem."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmip import _lib

_lib.require_gpu()
from tools.probe import probe
L = probe.lib()
for grid in (1, 148):
    for mode in (0, 1):
        for n in (64, 128, 256):
            buf = torch.zeros(2 * grid, dtype=torch.int64, device="cuda")
            iters, k = 64, 256
            _lib.check(L.dmip_debug_mma_bench(mode, n, k, iters, grid, buf.data_ptr(), None))
            torch.cuda.synchronize()
            c = buf.cpu().view(grid, 2).double()
            n_mma = iters * k // 16
            print(f"grid {grid:3d} A-in-{'tmem' if mode else 'smem'} N={n:3d}: issue {c[:,0].mean()/n_mma:6.1f} cyc/MMA, "
                  f"complete {c[:,1].mean()/n_mma:6.1f} cyc/MMA  -> {2*128*n*16/(c[:,1].mean()/n_mma):7.0f} FLOP/clk/SM")
