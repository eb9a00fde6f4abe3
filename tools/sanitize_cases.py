"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck), one tool per call:
    compute-sanitizer --tool memcheck python tools/sanitize_cases.py
K1 (k_tc_mlp: CDE, CDiffE, DPS; 2 tiles x 4 steps, with and without corrector), K1' (k_f32_mlp), K2/K3 (k_tcl_fwd / bwd /
wgrad: DSM, PINN, PosteriorLoss), K4 (k_surrogate, k_metropolis), metrics."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import gpu_cases as gc

torch.manual_seed(0)
for kind in ("CDE", "CDiffE", "Posterior"):
    m = gc._model(kind, 3, 23, (512, 512, 512), 5)
    y = torch.randn(23)
    for prec in ("bf16", "fp32"):
        for nc in (0, 1):
            x = m(y, num_samples=200, num_steps=4, precision=prec, seed=3, n_corrector=nc)
            assert x.shape == (200, 3)
    print("sampler", kind, "ok", flush=True)
for name in ("loss_dsm_cde_linear", "loss_pinn_cde_linear", "loss_pinn_cde_scat", "loss_pinn_cdiffe_scat"):
    for path in ("tc", "ffma"):
        os.environ["DMIP_LOSS_PATH"] = path
        err, tol, _ = gc.case_loss(name)
        print("loss", name, path, "PASS" if err <= tol else "FAIL", flush=True)
os.environ.pop("DMIP_LOSS_PATH", None)
err, tol, _ = gc.case_posterior_loss("loss_posterior_scat")
print("posterior loss", "PASS" if err <= tol else "FAIL", flush=True)
err, tol, _ = gc.case_surrogate_energy()
print("surrogate", "PASS" if err <= tol else "FAIL", flush=True)
err, tol, _ = gc.case_metropolis("philox")
print("metropolis", "PASS" if err <= tol else "FAIL", flush=True)
err, tol, _ = gc.case_histogram_kl()
print("metrics", "PASS" if err <= tol else "FAIL", flush=True)
