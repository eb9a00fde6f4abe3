"""Probe of the MN-major shared-memory operand descriptors and of the bf16x3 split product on B200 (tools/probe).
Prints the max error of D = A B^T against fp64 for every (A major, B major, N, LBO/SBO assignment) and the cycles per
MMA at the shapes the fused loss kernels use.  Each variant runs in its own process (a bad descriptor may trap)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = """
import sys, torch
sys.path.insert(0, {root!r})
from tools.probe import probe
torch.manual_seed(0)
n, k, a_mn, b_mn, split, swap, iters = {n}, {k}, {a_mn}, {b_mn}, {split}, {swap}, {iters}
a = torch.randn(128, k, device='cuda'); b = torch.randn(n, k, device='cuda')
# dense MN-major images: 64-element MN groups are `lbo` apart, 8-k groups `sbo` apart; atoms tile MN first
a_lbo, a_sbo = 1024, 1024 * 2
b_lbo, b_sbo = 1024, 1024 * max(n // 64, 1)
img = (a_lbo, a_sbo, b_lbo, b_sbo)
fld = (a_sbo, a_lbo, b_sbo, b_lbo) if swap else img
d, cyc = probe.umma2(a, b, a_mn, b_mn, split, img, fld, iters=iters, want_cycles=True)
if split:
    ref = (a.double() @ b.double().T)
else:
    ref = (a.bfloat16().double() @ b.bfloat16().double().T)
err = ((d.double() - ref).abs().max() / ref.abs().max()).item() if iters == 1 else float('nan')
n_mma = iters * (k // 16) * (3 if split else 1)
print('RESULT n=%d k=%d a_mn=%d b_mn=%d split=%d swap=%d  rel.err %.3e  cyc/MMA %.1f' % (n, k, a_mn, b_mn, split, swap, err, cyc[1] / n_mma))
"""


def run(**kw):
    code = CODE.format(root=ROOT, **kw)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
        out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(out[0] if out else f"CRASH {kw} rc={r.returncode} {(r.stderr.strip().splitlines() or ['?'])[-1][:200]}", flush=True)
    except subprocess.TimeoutExpired:
        print("TIMEOUT", kw, flush=True)


if __name__ == "__main__":
    for a_mn, b_mn in ((0, 0), (0, 1), (1, 1), (1, 0)):
        for n in (64, 128, 256):
            for swap in (0, 1):
                if swap and not (a_mn or b_mn):
                    continue
                run(n=n, k=64, a_mn=a_mn, b_mn=b_mn, split=0, swap=swap, iters=1)
    for n in (16, 64):
        run(n=n, k=128, a_mn=0, b_mn=1, split=1, swap=0, iters=1)
    run(n=256, k=64, a_mn=1, b_mn=1, split=1, swap=0, iters=1)
    for n in (16, 64, 128, 256):          # timing: long accumulation chains
        run(n=n, k=128, a_mn=0, b_mn=1, split=0, swap=0, iters=64)
        run(n=n, k=128, a_mn=1, b_mn=1, split=0, swap=0, iters=64)
