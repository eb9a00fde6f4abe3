"""End-to-end training through the drop-in `train_epoch` (pytest -m gpu): fused loss kernels + autograd glue + Adam."""
import pytest

import gpu_cases as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("loss_name", ["DSM", "PINN"])
def test_train_epoch_reduces_the_loss_and_learns_the_posterior(loss_name):
    err, tol, extra = gc.case_train_epoch(loss_name)
    assert err <= tol, extra
