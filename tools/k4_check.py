"""K4 A/B on one GPU: the tcgen05 surrogate kernel against the fp32 FFMA kernel (DMIP_SURROGATE_PATH=ffma) and the
reference fixture — errors in the units of the test tolerances, then the time of one call at 4,194,304 rows."""
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
DEV = "cuda"


def both(x, y, mode, want_fx=True):
    os.environ.pop("DMIP_SURROGATE_PATH", None)
    tc = us.surrogate_call(fm, x, y, 0.2, 0.01, 1000.0, mode=mode, want_fx=want_fx)
    n_tc = us.surrogate_call.last_launch_count
    os.environ["DMIP_SURROGATE_PATH"] = "ffma"
    ff = us.surrogate_call(fm, x, y, 0.2, 0.01, 1000.0, mode=mode, want_fx=want_fx)
    os.environ.pop("DMIP_SURROGATE_PATH", None)
    return tc, ff, n_tc


for n in (512, 77, 1, 128, 129, 100003):
    if n <= 512:
        x, y = fx["x"][:n].to(DEV), fx["y"][:n].to(DEV)
    else:
        g = torch.Generator().manual_seed(1)
        x = (torch.rand(n, 3, generator=g) * 2.4 - 1.2).to(DEV)
        y = fx["y"][torch.randint(0, 512, (n,), generator=g)].to(DEV)
    for mode in (us.SURR_ENERGY, us.SURR_LIK_VJP):
        (E, g_, f), (E2, g2, f2), nl = both(x, y, mode)
        torch.cuda.synchronize()
        e_f = (f - f2).abs().max().item() / 1e-5
        e_g = (g_ - g2).abs().max().item() / (2e-4 * g2.abs().max().item())
        e_E = ((E - E2).abs() / (2e-4 * E2.abs() + 1e-2)).max().item() if E is not None else 0.0
        print(f"n={n} mode={mode} launches={nl}  tc vs ffma: e_f={e_f:.3f} e_E={e_E:.4f} e_g={e_g:.4f}", flush=True)
E, g_, f = us.surrogate_call(fm, fx["x"].to(DEV), fx["y"].to(DEV), 0.2, 0.01, 1000.0, want_fx=True)
print("fixture: e_f", (f.cpu() - fx["fx"]).abs().max().item() / 1e-5, "e_E",
      ((E.cpu() - fx["E"]).abs() / (2e-4 * fx["E"].abs() + 1e-2)).max().item(), "e_g",
      (g_.cpu() - fx["grad"]).abs().max().item() / (2e-4 * fx["grad"].abs().max().item()))
n = 4194304
x = torch.rand(n, 3, device=DEV) * 2 - 1
y = fx["y"][:1].to(DEV).expand(n, -1).contiguous()
for path in ("tc", "ffma"):
    if path == "ffma":
        os.environ["DMIP_SURROGATE_PATH"] = "ffma"
    for _ in range(2):
        us.surrogate_call(fm, x, y, 0.2, 0.01, 1000.0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        us.surrogate_call(fm, x, y, 0.2, 0.01, 1000.0)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{path}: {ms:.3f} ms per {n} rows = {n / ms * 1e3:.3e} rows/s = {n * 550912 / ms * 1e-9:.1f} TFLOP/s algorithmic", flush=True)
