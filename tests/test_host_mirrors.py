"""CPU tests of the host-side mirrors that carry no kernel: the linear forward problem (linear_problem.py:5-65) and the
VP closed forms (sdes.py:21-49) of the `dmip` package against values produced by the unmodified reference
(fixtures vp_closed_forms / sampler_trained_cde_linear, see oracle/make_golden.py)."""
import torch

from util import load_golden


def test_linear_problem_matches_reference_closed_forms():
    from dmip.linear_problem import LinearForwardProblem
    lin = LinearForwardProblem()
    fx = load_golden("vp_closed_forms")
    x, y = fx["lin_x"], fx["lin_y"]
    assert torch.allclose(lin.score_posterior(x, y), fx["lin_score"], rtol=1e-5, atol=1e-5)
    # forward model: y = A x + b (datasets add 0.3-scaled noise on top; here the noiseless map)
    assert torch.allclose(lin(x), x @ torch.tensor([[1.0, 0.5], [0.0, 1.0]]).T + torch.tensor([0.3, 0.5]))
    # posterior of one observation vs the reference's get_posterior stored next to the trained sampler fixture
    fs = load_golden("sampler_trained_cde_linear")
    post = lin.get_posterior(fs["y"], device="cpu")
    assert torch.allclose(post.mean, fs["post_mean"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(post.covariance_matrix, fs["post_cov"], rtol=1e-5, atol=1e-6)
    # batched mean used by dmip.evaluation == per-observation posteriors
    ys = torch.stack([fs["y"], fs["y"] + 0.5])
    assert torch.allclose(lin.posterior_mean(ys)[0], post.mean, atol=1e-6)
    # the posterior score is the gradient of the posterior log-density
    xs = x[:8].clone().requires_grad_(True)
    lp = lin.get_posterior(fs["y"], device="cpu").log_prob(xs).sum()
    (g,) = torch.autograd.grad(lp, xs)
    assert torch.allclose(g, lin.score_posterior(xs.detach(), fs["y"].expand(8, 2)), rtol=2e-4, atol=2e-4)


def test_vp_sde_mirror_matches_reference_closed_forms():
    from dmip import sdes
    fx = load_golden("vp_closed_forms")
    sde = sdes.VariancePreservingSDE()
    t = fx["t"]
    assert torch.allclose(sde.beta(t), fx["beta"], rtol=1e-6)
    assert torch.allclose(sde.mean_weight(t), fx["alpha"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(sde.var(t), fx["var"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(sde.f(t, fx["y0"]), fx["f"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(sde.g(t, fx["y0"]), fx["g"], rtol=1e-6)
