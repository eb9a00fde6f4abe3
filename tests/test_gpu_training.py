"""End-to-end training through the drop-in `train_epoch` (pytest -m gpu): fused loss kernels + autograd glue + Adam."""
import pytest

import gpu_cases as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("loss_name", ["DSM", "PINN"])
def test_train_epoch_reduces_the_loss_and_learns_the_posterior(loss_name):
    err, tol, extra = gc.case_train_epoch(loss_name)
    assert err <= tol, extra


@pytest.mark.parametrize("kind", ["DSM", "PINN"])
def test_data_parallel_shards_sum_to_the_single_process_gradient(kind):
    """per-rank gradients with means over the GLOBAL batch, summed, equal the one-process step."""
    err, tol, extra = gc.case_data_parallel_emulated(kind)
    assert err <= tol, extra


@pytest.mark.parametrize("kind", ["DSM", "PINN"])
def test_data_parallel_two_ranks_nccl(kind):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    err, tol, extra = gc.case_data_parallel_nccl(kind)
    assert err <= tol, extra


def test_sample_t_on_the_device_matches_the_host_mirror():
    """models/diffusion.py:48-58 + sdes.py:51-57 as one kernel (no CPU draw, no host-to-device copy per batch)"""
    err, tol, extra = gc.case_sample_t_device()
    assert err <= tol, extra


@pytest.mark.parametrize("loss_name", ["DSM", "PINN"])
def test_train_epoch_fused_optimizer_step_equals_the_literal_triple(loss_name):
    err, tol, extra = gc.case_train_epoch_paths_agree(loss_name)
    assert err <= tol, extra


@pytest.mark.parametrize("loss_name,capturable", [("DSM", False), ("DSM", True), ("PINN", False)])
def test_train_epoch_as_cuda_graph_replays(loss_name, capturable):
    """graph=True: every batch is one replay of the captured step (eager or captured optimizer); same learning criteria"""
    err, tol, extra = gc.case_train_epoch(loss_name, graph=True, capturable=capturable)
    assert err <= tol, extra


def test_graphed_steps_of_successive_models_in_one_process():
    err, tol, extra = gc.case_graphed_steps_of_successive_models()
    assert err <= tol, extra
