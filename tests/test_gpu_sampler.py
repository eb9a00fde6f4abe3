"""GPU parity tests of the sampler / forward hot path (run on the B200 box: pytest -m gpu)."""
import pytest

import gpu_cases as gc

pytestmark = pytest.mark.gpu


def _ok(res):
    err, tol, extra = res
    assert err <= tol, (err, tol, {k: v for k, v in extra.items() if isinstance(v, float)})


@pytest.mark.parametrize("mode,n,k", [(0, 128, 64), (0, 128, 256), (0, 16, 128), (1, 128, 64), (1, 128, 256), (1, 48, 128)])
def test_umma_building_block(mode, n, k):
    _ok(gc.case_umma(mode, n, k))


@pytest.mark.parametrize("split", [1, 2, 3, 4])
def test_pack_image(split):
    _ok(gc.case_pack(split=split))


@pytest.mark.parametrize("name", ["mlp_cde_linear", "mlp_cdiffe_scat", "mlp_synth", "mlp_small"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward(name, precision):
    _ok(gc.case_forward(name, precision))     # mlp_small + bf16 dispatches to the fp32 kernels (widths != 512)


SAMPLERS = [("sampler_cde_linear", "CDE"), ("sampler_cde_linear_meanstd", "CDE"), ("sampler_cde_scat", "CDE"),
            ("sampler_cde_synth", "CDE"), ("sampler_cde_small", "CDE"), ("sampler_cdiffe_linear", "CDiffE"),
            ("sampler_cdiffe_scat", "CDiffE"), ("sampler_dps_scat", "Posterior")]


@pytest.mark.parametrize("name,kind", SAMPLERS)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sampler_injected_noise(name, kind, precision):
    _ok(gc.case_sampler(name, kind, precision))


@pytest.mark.parametrize("split", [1, 2, 3])
def test_sampler_layer0_split_modes(split):
    _ok(gc.case_sampler("sampler_cde_linear", "CDE", "bf16", split))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("mode", ["injected", "philox"])
def test_sampler_trained_reference_default_steps(precision, mode):
    _ok(gc.case_sampler_trained(precision, mode))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_shards_reproduce_the_single_gpu_run_bit_for_bit(precision):
    _ok(gc.case_shards_are_bit_identical(precision))


def test_batched_observations_match_per_observation_calls():
    _ok(gc.case_batched_observations())


def test_posterior_statistics_preserved_at_64k_particles():
    _ok(gc.case_posterior_statistics())


def test_edge_shapes_and_invalid_arguments():
    _ok(gc.case_edge_shapes())


def test_returned_host_arrays_are_never_overwritten():
    _ok(gc.case_host_results_do_not_alias())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sampler_trained_1000_steps(precision):
    _ok(gc.case_sampler_trained_1000_steps(precision))


@pytest.mark.parametrize("kind", ["CDiffE", "Posterior"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_philox_mode_equals_injected_mirror(kind, precision):
    _ok(gc.case_philox_equals_injected(kind, precision))


@pytest.mark.parametrize("kind", ["CDiffE", "Posterior"])
def test_batched_observations_cdiffe_dps(kind):
    _ok(gc.case_batched_observations_variant(kind))


@pytest.mark.parametrize("kind", ["CDiffE", "Posterior"])
def test_scatterometry_posterior_statistics_bf16_vs_fp32(kind):
    _ok(gc.case_scat_statistics(kind))


@pytest.mark.parametrize("kind", ["CDE", "CDiffE", "Posterior"])
@pytest.mark.parametrize("sde_kind,n_corr", [("VE", 0), ("VE", 1), ("VP", 2)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ve_sde_and_predictor_corrector_modes(kind, sde_kind, n_corr, precision):
    _ok(gc.case_pc_sampler(kind, sde_kind, n_corr, precision))


def test_ve_sampler_bf16_statistics_match_fp32():
    _ok(gc.case_ve_trained_posterior())
