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


def test_bf16x3_split_products_meet_the_surrogate_tolerances_on_the_reference_fixture():
    """The arithmetic of the tensor-core K4 kernel (csrc/dmip_surrogate_tc.cu), emulated on the CPU: the two 256 x 256
    layers and the output layer as bf16x3 split products (x = hi + lo, both bf16; hi hi + hi lo + lo hi), layer 0 and the
    input gradient in fp32 — against the reference's autograd values of scat_energy.npz in the units of the GPU test:
    f 2e-5, E 2e-4 rel (+1e-2), gradient 2e-4 of its scale on the rows away from a ReLU kink.  Plain bf16 operands must
    MISS them (that is why the kernel splits)."""
    import numpy as np
    from util import surrogate_params
    fx = load_golden("scat_energy")
    sp = [(W.double().numpy(), b.double().numpy()) for W, b in surrogate_params()]
    x, y = fx["x"].double().numpy(), fx["y"].double().numpy()

    def bf(v):
        return torch.from_numpy(np.asarray(v, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).double().numpy()

    def mm(A, B, parts):
        if parts == 0:
            return (A.astype(np.float32) @ B.astype(np.float32)).astype(np.float64)
        a_hi, b_hi = bf(A), bf(B)
        if parts == 1:
            return (a_hi @ b_hi).astype(np.float32).astype(np.float64)
        a_lo, b_lo = bf(A.astype(np.float32).astype(np.float64) - a_hi), bf(B.astype(np.float32).astype(np.float64) - b_hi)
        return (a_hi @ b_hi + a_hi @ b_lo + a_lo @ b_hi).astype(np.float32).astype(np.float64)

    def run(parts):
        f32 = lambda v: v.astype(np.float32).astype(np.float64)
        z, h = [], x
        for l, (W, b) in enumerate(sp):
            zz = f32(mm(h, W.T, 0 if l == 0 else parts) + b)
            z.append(zz)
            h = np.maximum(zz, 0) if l < 3 else zz
        f = h
        pre = (0.2 * f) ** 2 + 0.01 ** 2
        E = 0.5 * np.log(pre).sum(1) + 0.5 * ((y - f) ** 2 / pre).sum(1) + 1000.0 * (np.maximum(x - 1, 0) + np.maximum(-1 - x, 0)).sum(1)
        g = f32(0.04 * f / pre - (y - f) / pre - (y - f) ** 2 * 0.04 * f / pre ** 2)
        for l in (3, 2, 1, 0):
            g = f32(mm(g, sp[l][0], 0 if l == 0 else parts))
            if l > 0:
                g = g * (z[l - 1] > 0)
        g = g + 1000.0 * ((x > 1).astype(float) - (x < -1).astype(float))
        return f, E, g

    # rows away from a ReLU kink (tests/gpu_cases.py:_kink_rows)
    h, mn = x, np.full(x.shape[0], np.inf)
    for W, b in sp[:-1]:
        zz = h @ W.T + b
        mn = np.minimum(mn, np.abs(zz).min(1))
        h = np.maximum(zz, 0)
    keep = mn >= 2e-5
    assert keep.mean() >= 0.95
    ref_f, ref_E, ref_g = fx["fx"].double().numpy(), fx["E"].double().numpy(), fx["grad"].double().numpy()

    def errs(parts):
        f, E, g = run(parts)
        return (np.abs(f - ref_f).max() / 2e-5, (np.abs(E - ref_E) / (2e-4 * np.abs(ref_E) + 1e-2)).max(),
                np.abs(g - ref_g)[keep].max() / (2e-4 * np.abs(ref_g).max()))

    e3 = errs(3)
    assert max(e3) < 0.5, e3          # bf16x3: inside every tolerance with a factor 2 to spare
    e1 = errs(1)
    assert min(e1) > 10.0, e1         # plain bf16 operands: outside every tolerance by more than 10x
