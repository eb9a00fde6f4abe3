"""Staged GPU diagnostic: runs each parity case in its own subprocess (a device trap poisons the CUDA context of
the process that hit it) with a timeout, prints one line per case.  Usage:  python tools/gpu_diag.py [filter]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests")     # gpu_cases.py lives with the tests

CASES = [
    ("umma_ss_n128_k64", "case_umma(0, 128, 64)"),
    ("umma_ss_n128_k256", "case_umma(0, 128, 256)"),
    ("umma_ss_n16_k128", "case_umma(0, 16, 128)"),
    ("umma_ts_n128_k64", "case_umma(1, 128, 64)"),
    ("umma_ts_n128_k256", "case_umma(1, 128, 256)"),
    ("umma_ts_n48_k128", "case_umma(1, 48, 128)"),
    ("pack_split2", "case_pack(split=2)"),
    ("pack_split3", "case_pack(split=3)"),
    ("fwd_f32_linear", "case_forward('mlp_cde_linear', 'fp32')"),
    ("fwd_f32_small", "case_forward('mlp_small', 'fp32')"),
    ("fwd_f32_synth", "case_forward('mlp_synth', 'fp32')"),
    ("fwd_bf16_linear", "case_forward('mlp_cde_linear', 'bf16')"),
    ("fwd_bf16_cdiffe_scat", "case_forward('mlp_cdiffe_scat', 'bf16')"),
    ("fwd_bf16_synth", "case_forward('mlp_synth', 'bf16')"),
    ("smp_f32_cde_linear", "case_sampler('sampler_cde_linear', 'CDE', 'fp32')"),
    ("smp_f32_cde_small", "case_sampler('sampler_cde_small', 'CDE', 'fp32')"),
    ("smp_f32_cdiffe_scat", "case_sampler('sampler_cdiffe_scat', 'CDiffE', 'fp32')"),
    ("smp_f32_dps_scat", "case_sampler('sampler_dps_scat', 'Posterior', 'fp32')"),
    ("smp_bf16_cde_linear", "case_sampler('sampler_cde_linear', 'CDE', 'bf16')"),
    ("smp_bf16_cde_linear_split3", "case_sampler('sampler_cde_linear', 'CDE', 'bf16', 3)"),
    ("smp_bf16_cde_linear_split1", "case_sampler('sampler_cde_linear', 'CDE', 'bf16', 1)"),
    ("smp_bf16_cde_meanstd", "case_sampler('sampler_cde_linear_meanstd', 'CDE', 'bf16')"),
    ("smp_bf16_cde_scat", "case_sampler('sampler_cde_scat', 'CDE', 'bf16')"),
    ("smp_bf16_cde_synth", "case_sampler('sampler_cde_synth', 'CDE', 'bf16')"),
    ("smp_bf16_cdiffe_linear", "case_sampler('sampler_cdiffe_linear', 'CDiffE', 'bf16')"),
    ("smp_bf16_cdiffe_scat", "case_sampler('sampler_cdiffe_scat', 'CDiffE', 'bf16')"),
    ("smp_bf16_dps_scat", "case_sampler('sampler_dps_scat', 'Posterior', 'bf16')"),
    ("trained_f32_injected", "case_sampler_trained('fp32', 'injected')"),
    ("trained_f32_philox", "case_sampler_trained('fp32', 'philox')"),
    ("trained_bf16_injected", "case_sampler_trained('bf16', 'injected')"),
    ("trained_bf16_philox", "case_sampler_trained('bf16', 'philox')"),
] + [("loss_" + n[5:], f"case_loss('{n}')") for n in [
    "loss_dsm_small", "loss_dsm_cde_linear", "loss_dsm_cdiffe_linear", "loss_dsm_cde_scat",
    "loss_pinn_small", "loss_pinn_cde_linear", "loss_pinn_cde_linear_g3", "loss_pinn_cde_linear_l2l1",
    "loss_pinn_cde_linear_cfpe", "loss_pinn_cde_scat", "loss_pinn_cdiffe_linear",
    "loss_dsmpde_cde_linear", "loss_dsmpde_cde_linear_cfpe",
    "loss_pinn_cdiffe_scat", "loss_pinn_cde_scat_hutch", "loss_dsmpde_cdiffe_scat_hutch"]] + [
    ("loss_adj_" + n[5:], f"case_loss('{n}', 'exact_adjoint')") for n in [
        "loss_pinn_small", "loss_pinn_cde_linear_g3", "loss_pinn_cde_scat", "loss_pinn_cdiffe_linear",
        "loss_dsmpde_cde_linear"]] + [
    ("surr_energy", "case_surrogate_energy()"),
    ("surr_vjp", "case_surrogate_vjp()"),
    ("loss_posterior_scat", "case_posterior_loss('loss_posterior_scat')"),
    ("loss_posterior_small", "case_posterior_loss('loss_posterior_small')"),
    ("shards_bf16", "case_shards_are_bit_identical('bf16')"),
    ("shards_fp32", "case_shards_are_bit_identical('fp32')"),
    ("batched_obs", "case_batched_observations()"),
    ("posterior_stats", "case_posterior_statistics()"),
    ("edge_shapes", "case_edge_shapes()"),
    ("train_dsm", "case_train_epoch('DSM')"),
    ("train_pinn", "case_train_epoch('PINN')"),
    ("hist_kl", "case_histogram_kl()"),
    ("mcmc_injected", "case_metropolis('injected')"),
    ("mcmc_philox", "case_metropolis('philox')"),
    ("evaluate", "case_evaluate()"),
    ("host_alias", "case_host_results_do_not_alias()"),
]

TEMPLATE = """
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})
from gpu_cases import *
err, tol, extra = {expr}
info = ' '.join(f'{{k}}={{v:.3g}}' for k, v in extra.items() if isinstance(v, float))
print('RESULT', 'PASS' if err <= tol else 'FAIL', f'err={{err:.3e}} tol={{tol:.3e}}', info)
if err > tol and 'out' in extra:
    o, r = extra['out'], extra['ref']
    bad = ((o - r).abs() > tol).nonzero()
    print('  n_bad', len(bad), 'of', o.numel(), 'first', bad[:6].tolist(), 'nan', int(torch.isnan(o).sum()))
    print('  out', o.flatten()[:6].tolist()); print('  ref', r.flatten()[:6].tolist())
if err > tol and 'd' in extra:
    o, r = extra['d'], extra['ref']
    bad = ((o - r).abs() > tol)
    print('  n_bad', int(bad.sum()), 'rows', bad.any(1).nonzero().flatten()[:16].tolist(), 'cols', bad.any(0).nonzero().flatten()[:16].tolist(), 'nan', int(torch.isnan(o).sum()))
    print('  d[0,:8]', o[0,:8].tolist()); print('  r[0,:8]', r[0,:8].tolist())
    print('  d[1,:8]', o[1,:8].tolist()); print('  r[1,:8]', r[1,:8].tolist())
"""


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    n_fail = 0
    for name, expr in CASES:
        if flt and flt not in name:
            continue
        code = TEMPLATE.format(root=ROOT, here=HERE, expr=expr)
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT") or l.startswith("  ")]
            if r.returncode != 0 or not lines:
                n_fail += 1
                print(f"{name:32s} CRASH rc={r.returncode}", (r.stderr.strip().splitlines() or ["?"])[-1][:300], flush=True)
            else:
                n_fail += "FAIL" in lines[0]
                print(f"{name:32s}", "\n".join(lines), flush=True)
        except subprocess.TimeoutExpired:
            n_fail += 1
            print(f"{name:32s} TIMEOUT", flush=True)
    print("failures:", n_fail)
    return 1 if n_fail else 0


if __name__ == "__main__":
    sys.exit(main())
