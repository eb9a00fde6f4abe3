"""tcgen05 loss path vs the fp32 FFMA path vs the reference fixture, one subprocess per case (a device trap poisons the
process).  Prints the four loss scalars of both paths and the relative error of every parameter gradient.
Usage: python tools/tcl_diag.py [filter]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = """
import os, sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/tests')
name = {name!r}
import gpu_cases as gc
from dmip import losses as dl
out = {{}}
for path in ('ffma', 'tc'):
    os.environ['DMIP_LOSS_PATH'] = path
    torch.manual_seed(0)
    if name.startswith('loss_posterior'):
        err, tol, ex = gc.case_posterior_loss(name)
    else:
        err, tol, ex = gc.case_loss(name)
    out[path] = (err, tol, ex)
    print('RESULT %-32s %-5s %s err=%.3e tol=%.1e loss=%.7g ref=%.7g' % (name, path, 'PASS' if err <= tol else 'FAIL', err, tol, ex['loss'], ex['ref']), flush=True)
"""


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    names = ["loss_dsm_cde_linear", "loss_dsmpde_cde_linear_cfpe", "loss_pinn_cde_linear_cfpe", "loss_pinn_cde_linear",
             "loss_dsm_cdiffe_linear", "loss_dsm_cde_scat", "loss_pinn_cde_linear_g3", "loss_pinn_cde_linear_l2l1",
             "loss_dsmpde_cde_linear", "loss_pinn_cde_scat", "loss_pinn_cdiffe_linear", "loss_pinn_cdiffe_scat",
             "loss_pinn_cde_scat_hutch", "loss_dsmpde_cdiffe_scat_hutch", "loss_posterior_scat", "loss_pinn_small"]
    for name in names:
        if flt and flt not in name:
            continue
        try:
            r = subprocess.run([sys.executable, "-c", CODE.format(root=ROOT, name=name)], capture_output=True, text=True, timeout=300)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT") or "grad mismatch" in l]
            print("\n".join(lines) if lines else f"{name}: no output", flush=True)
            if r.returncode != 0:
                print(f"  CRASH rc={r.returncode}", " | ".join((r.stderr.strip().splitlines() or ["?"])[-3:])[:600], flush=True)
        except subprocess.TimeoutExpired:
            print(f"{name}: TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
