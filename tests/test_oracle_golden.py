"""Pin the CPU oracle (oracle/) against outputs of the reference itself.

The fixtures under tests/golden/ were produced by oracle/make_golden.py, which
runs the unmodified reference classes on seeded inputs with injected noise.
fp32 comparisons: the oracle restates the same fp32 torch ops, so samples agree
to ~1e-6; the loss oracle replaces autograd double-backward by explicit jets, so
losses/grads agree to fp32 round-off of a different summation order (1e-4 rel).
"""
import pytest
import torch

import oracle
from oracle import losses as ol, nets as on, sampler as osamp, scatterometry as oscat, vp
from oracle.weights import make_params
from util import check_grads, load_golden, meta_hidden, surrogate_params


@pytest.mark.parametrize("name", ["mlp_cde_linear", "mlp_cdiffe_scat", "mlp_synth", "mlp_small"])
def test_mlp_forward(name):
    fx = load_golden(name)
    seed, xdim, ydim, out_dim = (int(v) for v in fx["meta"][:4])
    params = make_params(seed, xdim + ydim + 1, out_dim, meta_hidden(fx, 4))
    out = on.mlp(params, fx["x"], fx["y"], fx["t"])
    assert torch.allclose(out, fx["out"], rtol=1e-5, atol=2e-6)
    # single-tanh network must NOT match (double-tanh quirk)
    W, b = params[0]
    h = torch.tanh(torch.cat([fx["x"], fx["y"], fx["t"]], 1) @ W.T + b)
    for W, b in params[1:-1]:
        h = torch.tanh(h @ W.T + b)
    single = h @ params[-1][0].T + params[-1][1]
    assert (single - fx["out"]).abs().max() > 1e-3


@pytest.mark.parametrize("name", ["sampler_cde_linear", "sampler_cde_linear_meanstd", "sampler_cde_scat",
                                  "sampler_cde_synth", "sampler_cde_small"])
def test_sampler_cde(name):
    fx = load_golden(name)
    seed, xdim, ydim, N, S = (int(v) for v in fx["meta"][:5])
    params = make_params(seed, xdim + ydim + 1, xdim, meta_hidden(fx, 5))
    mean, std = fx["mean_std"].tolist()
    out = osamp.em_sampler_cde(params, fx["y"], fx["x0"] * std + mean, fx["noise"], S)
    assert out.shape == (N, xdim)
    assert (out - fx["out"]).abs().max() <= 1e-6 * fx["out"].abs().max() + 1e-6


@pytest.mark.parametrize("name", ["sampler_cdiffe_linear", "sampler_cdiffe_scat"])
def test_sampler_cdiffe(name):
    fx = load_golden(name)
    seed, xdim, ydim, N, S = (int(v) for v in fx["meta"][:5])
    params = make_params(seed, xdim + ydim + 1, xdim + ydim, meta_hidden(fx, 5))
    out = osamp.em_sampler_cdiffe(params, fx["y"], fx["x0"], fx["ynoise"], fx["noise"], S)
    assert (out - fx["out"]).abs().max() <= 1e-6 * fx["out"].abs().max() + 1e-6


def test_sampler_dps():
    fx = load_golden("sampler_dps_scat")
    seed, xdim, ydim, N, S = (int(v) for v in fx["meta"][:5])
    hid = meta_hidden(fx, 5)
    pp = make_params(seed, xdim + 1, xdim, hid)
    lp = make_params(seed + 100, xdim + ydim + 1, xdim, hid)
    out = osamp.em_sampler_dps(pp, lp, fx["y"], fx["x0"], fx["noise"], S)
    assert (out - fx["out"]).abs().max() <= 2e-6 * fx["out"].abs().max() + 1e-6


def _leaf(params):
    return [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in params]


LOSS_CASES = {
    # name: (model, kind, kwargs, gain)
    "loss_dsm_cde_linear": ("CDE", "DSM", {}, 1.0),
    "loss_dsm_cdiffe_linear": ("CDiffE", "DSM", {}, 1.0),
    "loss_dsm_cde_scat": ("CDE", "DSM", {}, 1.0),
    "loss_dsm_small": ("CDE", "DSM", {}, 1.0),
    "loss_pinn_cde_linear": ("CDE", "PINN", dict(lam=0.001, lam2=0.1, pde_loss="FPE", ic_metric="L2", pde_metric="L1"), 1.0),
    "loss_pinn_cde_linear_g3": ("CDE", "PINN", dict(lam=0.001, lam2=0.1, pde_loss="FPE", ic_metric="L2", pde_metric="L1"), 3.0),
    "loss_pinn_cde_linear_l2l1": ("CDE", "PINN", dict(lam=0.5, lam2=0.7, pde_loss="FPE", ic_metric="L1", pde_metric="L2"), 1.0),
    "loss_pinn_cde_linear_cfpe": ("CDE", "PINN", dict(lam=0.01, lam2=0.1, pde_loss="cScoreFPE", ic_metric="L2", pde_metric="L2"), 1.0),
    "loss_pinn_cde_scat": ("CDE", "PINN", dict(lam=0.01, lam2=0.001, pde_loss="FPE", ic_metric="L2", pde_metric="L1"), 1.0),
    "loss_pinn_cdiffe_linear": ("CDiffE", "PINN", dict(lam=0.001, lam2=0.1, pde_loss="FPE", ic_metric="L2", pde_metric="L1"), 1.0),
    "loss_pinn_small": ("CDE", "PINN", dict(lam=0.001, lam2=0.1, pde_loss="FPE", ic_metric="L2", pde_metric="L1"), 1.0),
    "loss_dsmpde_cde_linear": ("CDE", "DSM_PDE", dict(lam=0.1, pde_loss="FPE", pde_metric="L1"), 1.0),
    "loss_dsmpde_cde_linear_cfpe": ("CDE", "DSM_PDE", dict(lam=0.1, pde_loss="cScoreFPE", pde_metric="L1"), 1.0),
    # PINNLoss2 (losses.py:245-291) with the attribute its forward reads but never sets supplied (oracle/make_golden.py)
    "loss_pinn2_cde_linear": ("CDE", "PINN2", dict(lam=0.01, lam2=0.1, pde_loss="FPE", ic_metric="L2"), 1.0),
    "loss_pinn2_cde_scat_cfpe": ("CDE", "PINN2", dict(lam=0.02, lam2=0.05, pde_loss="cScoreFPE", ic_metric="L1"), 1.0),
    "loss_pinn_cdiffe_scat": ("CDiffE", "PINN", dict(lam=0.01, lam2=0.001, pde_loss="FPE", ic_metric="L2", pde_metric="L1"), 1.0),
    "loss_pinn_cde_scat_hutch": ("CDE", "PINN", dict(lam=0.01, lam2=0.001, pde_loss="FPE", ic_metric="L2", pde_metric="L1",
                                                     divergence_method="hutchinson"), 1.0),
    "loss_dsmpde_cdiffe_scat_hutch": ("CDiffE", "DSM_PDE", dict(lam=0.05, pde_loss="FPE", pde_metric="L2",
                                                                divergence_method="approx"), 1.0),
}


@pytest.mark.parametrize("name", sorted(LOSS_CASES))
def test_losses(name):
    model, kind, kw, gain = LOSS_CASES[name]
    fx = load_golden(name, torch.float64)
    seed, xdim, ydim, B = (int(v) for v in fx["meta"][:4])
    out_dim = xdim if model == "CDE" else xdim + ydim
    # fp64 oracle vs fp32 reference: removes the oracle's own round-off from the comparison
    params = _leaf(make_params(seed, xdim + ydim + 1, out_dim, meta_hidden(fx, 4), gain=gain, dtype=torch.float64))
    x, y, t, eps = fx["x"], fx["y"], fx["t"], fx["eps"]
    kw = dict(kw)
    if kw.pop("divergence_method", "exact") != "exact":
        kw["probe"] = fx["probe"]                      # the v of div_estimator (losses.py:28-40) stored with the fixture
    if kind == "DSM":
        loss, info = ol.dsm_loss(params, model, x, y, t, eps), {}
    elif kind == "PINN":
        loss, info = ol.pinn_loss(params, model, x, y, t, eps, fx["ic_target"], **kw)
    elif kind == "PINN2":
        loss, info = ol.pinn2_loss(params, model, x, y, t, eps, fx["ic_target"], **kw)
    else:
        loss, info = ol.dsm_pde_loss(params, model, x, y, t, eps, **kw)
    assert abs(loss.item() - fx["loss"].item()) <= 2e-4 * abs(fx["loss"].item()) + 1e-6
    for k, v in info.items():
        ref = fx["info_" + k.replace(" ", "_").replace("-", "_")].item()
        assert abs(v.item() - ref) <= 3e-4 * abs(ref) + 1e-6, (k, v.item(), ref)
    loss.backward()
    check_grads(fx, [(W.grad, b.grad) for W, b in params], rtol=2e-3, atol_frac=2e-3)


def test_pinn_scat_ic_target_matches_closed_form():
    fx = load_golden("loss_pinn_cde_scat", torch.float64)
    sp = surrogate_params(torch.float64)
    ic = oscat.score_posterior(sp, fx["x"], fx["y"])
    assert torch.allclose(ic, fx["ic_target"], rtol=2e-4, atol=2e-3)


@pytest.mark.parametrize("name", ["loss_posterior_scat", "loss_posterior_small"])
def test_posterior_loss(name):
    fx = load_golden(name, torch.float64)
    seed, xdim, ydim, B = (int(v) for v in fx["meta"][:4])
    hid = meta_hidden(fx, 4)
    pp = _leaf(make_params(seed, xdim + 1, xdim, hid, dtype=torch.float64))
    lp = _leaf(make_params(seed + 100, xdim + ydim + 1, xdim, hid, dtype=torch.float64))
    sp = surrogate_params(torch.float64)
    loss, info = ol.posterior_loss(pp, lp, sp, fx["x"], fx["y"], fx["t"], fx["eps"], float(fx["lam"]))
    assert abs(loss.item() - fx["loss"].item()) <= 5e-4 * abs(fx["loss"].item())
    assert abs(info["PriorLoss"].item() - fx["info_PriorLoss"].item()) <= 5e-4 * abs(fx["info_PriorLoss"].item())
    assert abs(info["LikelihoodLoss"].item() - fx["info_LikelihoodLoss"].item()) <= 5e-4 * abs(fx["info_LikelihoodLoss"].item())
    loss.backward()
    check_grads(fx, [(W.grad, b.grad) for W, b in pp], prefix="prior_", rtol=3e-3, atol_frac=3e-3)
    check_grads(fx, [(W.grad, b.grad) for W, b in lp], prefix="lik_", rtol=3e-3, atol_frac=3e-3)


def test_scat_energy_and_grad():
    fx = load_golden("scat_energy")
    sp = surrogate_params()
    assert torch.allclose(oscat.surrogate(sp, fx["x"]), fx["fx"], rtol=1e-5, atol=1e-5)
    E, g = oscat.energy_and_grad(sp, fx["x"], fx["y"])
    assert torch.allclose(E, fx["E"], rtol=2e-4, atol=1e-2)
    assert torch.allclose(oscat.energy(sp, fx["x"], fx["y"]), fx["E"], rtol=1e-5, atol=1e-3)
    scale = fx["grad"].abs().max()
    assert (g - fx["grad"]).abs().max() <= 2e-4 * scale


def test_vp_closed_forms():
    fx = load_golden("vp_closed_forms")
    t = fx["t"]
    assert torch.allclose(vp.beta(t), fx["beta"])
    assert torch.allclose(vp.mean_weight(t), fx["alpha"])
    assert torch.allclose(vp.var(t), fx["var"])
    yt, std, g = vp.perturb(t, fx["y0"], fx["eps"])
    assert torch.allclose(yt, fx["yt"]) and torch.allclose(std, fx["std"]) and torch.allclose(g, fx["g"])
    assert torch.allclose(vp.f(t, fx["y0"]), fx["f"])
    assert torch.allclose(vp.mean_weight(t) ** 2 + vp.var(t), torch.ones_like(t), atol=1e-6)


def test_philox_known_answers():
    import numpy as np
    r = oracle.philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]      # Random123 KAT
    r = oracle.philox.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(v) for v in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    z = oracle.philox.normals(np.arange(200000), 7, 0, 6, 99)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3


def trained_cde_linear_params(dtype=torch.float32):
    fx = load_golden("trained_cde_linear", dtype)
    return [(fx[f"{k}_weight"], fx[f"{k}_bias"]) for k in (0, 3, 5, 7)]


def test_sampler_trained_cde_linear_philox_noise():
    """Reference default S=200 on a trained net, noise from the keyed Philox stream."""
    import numpy as np
    fx = load_golden("sampler_trained_cde_linear")
    N, S, seed = (int(v) for v in fx["philox"])
    gidx = np.arange(N)
    x0 = torch.from_numpy(oracle.philox.normals(gidx, oracle.philox.STEP_INIT, 0, 2, seed))
    noise = torch.from_numpy(np.stack([oracle.philox.normals(gidx, i, 0, 2, seed) for i in range(S)]))
    out = osamp.em_sampler_cde(trained_cde_linear_params(), fx["y"], x0, noise, S)
    assert (out - fx["out"]).abs().max() < 2e-5
    # the trained model reproduces the analytic posterior (linear_problem.py:41-46)
    assert (out.mean(0) - fx["post_mean"]).abs().max() < 0.08
    assert (torch.cov(out.T) - fx["post_cov"]).abs().max() < 0.05
    # score-MSE metric of main_diffusion_linear.py:78-83 on the stored probe points
    s0 = on.mlp(trained_cde_linear_params(), fx["score_x"], fx["score_y"], torch.zeros(256, 1)) / 0.1 ** 0.5
    assert torch.allclose(s0, fx["score_net"], rtol=1e-4, atol=1e-4)


def test_metrics_oracle_kl_properties():
    """oracle.metrics restates main_diffusion_linear.py:84-117: identical sample sets give KL 0; a shifted set a positive
    KL; the sum over repeats equals the histogram of the concatenated repeats."""
    import numpy as np
    from oracle import metrics as omet
    rng = np.random.default_rng(1)
    a = [rng.standard_normal((5000, 2)).astype(np.float32) for _ in range(3)]
    b = [x + 0.5 for x in a]
    bins, ranges = (75, 75), ((-3.5, 3.5), (-3.5, 3.5))
    ha, hb = omet.hist_sum(a, bins, ranges), omet.hist_sum(b, bins, ranges)
    assert omet.kl2(ha, ha) == 0.0
    assert omet.kl2(ha, hb) > 0.05
    assert np.array_equal(ha, np.histogramdd(np.concatenate(a), bins=bins, range=ranges)[0])


def test_metropolis_oracle_matches_reference_chains():
    """oracle.mcmc.anneal_to_energy vs the reference's anneal_to_energy run on the stored normals / uniforms
    (fixture mcmc_scat): identical accept decisions, final points and energy differences."""
    from oracle import mcmc as omc
    fx = load_golden("mcmc_scat")
    n_obs, n_per, S = (int(v) for v in fx["meta"])
    sp = surrogate_params()
    y = fx["y"].repeat_interleave(n_per, 0)
    x, de = omc.anneal_to_energy(sp, fx["x0"], y, fx["noise"], fx["unif"], float(fx["noise_std"]))
    moved = ((fx["out"] - fx["x0"]).abs().max(1).values > 0).float().mean().item()
    assert moved > 0.5, moved                                      # the chains did move
    same = ((x - fx["out"]).abs().max(1).values <= 1e-6)
    assert same.float().mean().item() >= 0.99                      # a near-tie may flip an accept in fp32
    assert torch.allclose(de[same], fx["de"][same], rtol=1e-4, atol=2e-2)


def test_predictor_corrector_oracle_reduces_to_the_pinned_reference_sampler():
    """oracle/ve.py generalises the sampler to the VE-SDE and a Langevin corrector (no upstream counterpart).  With the
    VP-SDE and no corrector it must be bit-identical to the restatement pinned against the reference fixtures."""
    import torch
    from oracle import sampler as osamp, ve
    from oracle.weights import make_params
    torch.manual_seed(0)
    p, p2, pc = make_params(1, 27, 3, (64, 64)), make_params(2, 4, 3, (64, 64)), make_params(3, 27, 26, (64, 64))
    y, x0, S = torch.randn(23), torch.randn(50, 3), 7
    noise, yn = torch.randn(S, 50, 3), torch.randn(S, 50, 23)
    a = osamp.em_sampler_cde(p, y, x0, noise, S)
    assert torch.equal(a, ve.pc_sampler(ve.net_fn("CDE", p), "CDE", y, x0, noise, S, ve.VP()))
    a = osamp.em_sampler_dps(p2, p, y, x0, noise, S)
    assert torch.equal(a, ve.pc_sampler(ve.net_fn("DPS", p, p2), "DPS", y, x0, noise, S, ve.VP()))
    a = osamp.em_sampler_cdiffe(pc, y, x0, yn, noise, S)
    assert torch.equal(a, ve.pc_sampler(ve.net_fn("CDiffE", pc), "CDiffE", y, x0, noise, S, ve.VP(), ynoise=yn))
    # the corrector changes the result and keeps it finite; VE closed forms
    b = ve.pc_sampler(ve.net_fn("CDE", p), "CDE", y, x0, torch.randn(2 * S, 50, 3), S, ve.VP(), n_corr=1)
    assert torch.isfinite(b).all() and not torch.allclose(a, b)
    sde = ve.VE(0.01, 50.0)
    t = torch.tensor([[0.0], [0.5], [1.0]])
    assert torch.allclose(sde.var(t) ** 0.5, torch.tensor([[0.01], [0.01 * 5000 ** 0.5], [50.0]]), rtol=1e-5)
    # g^2 = d sigma^2 / dt
    h = 1e-3
    num = (sde.var(t + h) - sde.var(t - h)) / (2 * h)
    assert torch.allclose(sde.g(t, t) ** 2, num, rtol=2e-4)


def test_ve_host_class_matches_the_oracle_closed_forms():
    import torch
    from dmip import sdes
    from oracle import ve
    a, b = sdes.VarianceExplodingSDE(0.02, 7.0), ve.VE(0.02, 7.0)
    t = torch.rand(16, 1)
    y = torch.randn(16, 3)
    assert torch.allclose(a.var(t), b.var(t)) and torch.allclose(a.g(t, y), b.g(t, y)) and torch.equal(a.f(t, y), b.f(t, y))
    assert torch.equal(a.mean_weight(t), torch.ones_like(t))
    yt, eps, std, g = a.sample(t, y, return_noise=True)
    assert torch.allclose(yt, y + std * eps)
