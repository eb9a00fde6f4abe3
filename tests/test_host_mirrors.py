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


def test_linear_log_posterior_matches_the_reference_expression():
    """linear_problem.py:48-58, including its quirky batch mean  y_res @ (A^T Sigma_y^-1)  (ADVICE round 1)."""
    from dmip.linear_problem import LinearForwardProblem
    lin = LinearForwardProblem()
    fx = load_golden("linear_log_posterior")
    assert torch.allclose(lin.log_posterior(fx["x"], fx["y"]), fx["log_posterior"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(lin.score_posterior(fx["x"], fx["y"]), fx["score"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(lin(fx["x"]), fx["fwd"], rtol=1e-6, atol=1e-6)
    post = lin.get_posterior(fx["y"][0], device="cpu")
    assert torch.allclose(post.mean, fx["post_mean0"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(post.covariance_matrix, fx["post_cov"], rtol=1e-5, atol=1e-6)


def test_debiased_t_sampler_follows_its_analytic_distribution():
    """sdes.py:51-57 draws t ~ q(t) ∝ beta(t) / var(t), flat below t_epsilon (sdeflow-light's truncated VP sampler; the
    module is not vendored upstream — parity unpinned).  What CAN be pinned: the inverse CDF must
    invert the analytic CDF (round trip), and draws must pass a Kolmogorov–Smirnov test against it."""
    import math

    import numpy as np
    from scipy import stats

    from dmip import sdes
    sde = sdes.VariancePreservingSDE()
    bmin, bmax, te, T = sde.beta_min, sde.beta_max, sde.t_epsilon, sde.T
    db = bmax - bmin

    def big_b(t):
        return 0.5 * t * t * db + t * bmin

    def antider(t):                       # integral of beta / var = log(1 - exp(-B)) + B
        return np.log1p(-np.exp(-big_b(t))) + big_b(t)

    r_eps = (bmin + db * te) / (1.0 - math.exp(-big_b(te)))
    Z = r_eps * te + antider(T) - antider(te)

    def cdf(t):
        t = np.asarray(t, dtype=np.float64)
        return np.where(t <= te, r_eps * t, r_eps * te + antider(np.maximum(t, te)) - antider(te)) / Z

    u = torch.linspace(1e-4, 1 - 1e-4, 4001, dtype=torch.float64)
    t = sdes.vp_truncated_inverse_cdf(u, bmin, bmax, te, T)
    assert float(t.min()) > 0 and float(t.max()) <= T
    assert np.allclose(cdf(t.numpy()), u.numpy(), atol=2e-6)
    assert bool((t[1:] >= t[:-1]).all())
    torch.manual_seed(11)
    draws = sde.sample_debiasing_t([20000, 1]).double().view(-1).numpy()
    assert draws.min() > 0 and draws.max() <= T
    ks = stats.kstest(draws, cdf)
    assert ks.pvalue > 1e-3, ks
    # the density really is ∝ beta / var above t_epsilon: histogram ratio test on a coarse grid
    edges = np.array([te, 0.01, 0.05, 0.2, 0.5, 1.0])
    emp = np.histogram(draws, bins=edges)[0] / len(draws)
    assert np.allclose(emp, np.diff(cdf(edges)), atol=0.012)
