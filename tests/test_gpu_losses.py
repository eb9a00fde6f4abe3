"""GPU parity tests of the fused score-training losses and the scatterometry surrogate kernel (pytest -m gpu)."""
import pytest

import gpu_cases as gc

pytestmark = pytest.mark.gpu


def _ok(res):
    err, tol, extra = res
    assert err <= tol, (err, tol, {k: v for k, v in extra.items() if isinstance(v, float)})


@pytest.fixture(params=["tc", "ffma"])
def loss_path(request, monkeypatch):
    """Both kernel families against the same reference fixtures: 'tc' = the tcgen05 kernels wherever they apply
    (csrc/dmip_tcl.cu; the default), 'ffma' = the fp32 FFMA kernels of csrc/dmip_loss.cu forced for every pass."""
    monkeypatch.setenv("DMIP_LOSS_PATH", request.param)
    return request.param


@pytest.mark.parametrize("name", sorted(gc.LOSS_CASES))
def test_fused_loss(name, loss_path):
    _ok(gc.case_loss(name))


@pytest.mark.parametrize("name", ["loss_pinn_small", "loss_pinn_cde_linear_g3", "loss_pinn_cde_scat",
                                  "loss_pinn_cdiffe_linear", "loss_dsmpde_cde_linear"])
def test_fused_loss_exact_divergence_adjoint_route(name, loss_path):
    """d <= 4 runs forward-only by default; the adjoint route (what d = 26 uses) must give the same reference values"""
    _ok(gc.case_loss(name, "exact_adjoint"))


@pytest.mark.parametrize("name", ["loss_posterior_scat", "loss_posterior_small"])
def test_posterior_loss(name, loss_path):
    _ok(gc.case_posterior_loss(name))


def test_tensor_core_loss_path_is_the_one_that_runs():
    """The [512,512,512] fixtures must take the tcgen05 kernels (5 launches: pack, forward, backward, 4 weight-gradient
    GEMMs = 7) — not silently the FFMA kernels (11 launches for the same pass)."""
    _ok(gc.case_loss_path_taken())


@pytest.mark.parametrize("path", ["tc", "ffma"])
def test_surrogate_energy_and_score(path):
    _ok(gc.case_surrogate_energy(path))


@pytest.mark.parametrize("path", ["tc", "ffma"])
def test_surrogate_likelihood_vjp(path):
    _ok(gc.case_surrogate_vjp(path))


@pytest.mark.parametrize("path", ["tc", "ffma"])
def test_surrogate_one_observation_per_block_of_rows(path):
    _ok(gc.case_surrogate_observation_blocks(path))


def test_surrogate_tensor_core_kernel_other_input_and_output_widths():
    _ok(gc.case_surrogate_other_shapes())


def test_surrogate_tensor_core_kernel_matches_the_fp32_kernel_on_100003_rows():
    _ok(gc.case_surrogate_tc_vs_ffma())


def test_histogram_kl_matches_numpy_bit_for_bit():
    _ok(gc.case_histogram_kl())


@pytest.mark.parametrize("mode", ["injected", "philox"])
def test_metropolis_chains_match_reference(mode):
    _ok(gc.case_metropolis(mode))


@pytest.mark.parametrize("problem", ["linear", "scat"])
def test_evaluation_tables_match_the_reference_evaluate_loops(problem):
    _ok(gc.case_evaluate_parity(problem))


def test_evaluate_loops_run_on_gpu():
    _ok(gc.case_evaluate())


@pytest.mark.parametrize("kind", ["DSM", "PINN"])
def test_fused_loss_at_baseline_batch_65536(kind, loss_path):
    """BASELINE configs[1] size against the chunked fp64 oracle (sum-of-means identity)"""
    _ok(gc.case_loss_at_baseline_batch(kind))


@pytest.mark.parametrize("kind", ["DSM", "PINN"])
def test_fused_loss_repeats_bit_for_bit_up_to_atomics(kind):
    """the same step twelve times at a ragged batch: the cross-CTA protocol of the pair kernels has no timing dependence"""
    _ok(gc.case_loss_repeats(kind))


def test_kernels_stay_inside_their_buffers():
    """canary-guarded outputs and workspaces over ragged sizes (the stand-in for compute-sanitizer, closed on this pool)"""
    _ok(gc.case_guarded_buffers())


def test_documented_boundary_deviations():
    _ok(gc.case_boundary_deviations())


@pytest.mark.parametrize("which", ["cde_linear", "cdiffe_scat", "mlp2"])
def test_autograd_through_the_score_net_runs_on_the_library_kernels(which):
    _ok(gc.case_mlp_autograd(which))
