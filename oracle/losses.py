"""Score-training losses — oracle restatement of losses.py:14-52,77-98,116-124,
143-164,214-242,349-386 and the loss dispatch of models/diffusion.py:74-105,
123-156,204-229.  TEST INFRASTRUCTURE ONLY.

The Score-FPE terms are restated with explicit forward-mode jets
 instead of the reference's 2d+1 double-backward
autograd calls; `tests/test_oracle_golden.py` pins this restatement against
the reference's own autograd results.  Facts that decide parity:
  Q8  ds_dt is the TOTAL derivative along x_t(t) at fixed eps (t is a leaf, x_t depends on it);
  Q9  grad_x is detached (no create_graph at losses.py:89-90);
  Q10 (B,)+(B,1) broadcasting makes mean(loss) == mean(dsm)+mean(ic)+mean(pde).
"""
import torch

from . import vp
from . import nets as onets
from . import scatterometry as scat


def dsm(s, std, eps):
    """DSMLoss.forward (losses.py:49-52) -> (B,)"""
    return ((s * std + eps) ** 2).reshape(s.shape[0], -1).sum(1) / 2


def _xdot(t, x0, eps):
    """d/dt [eps*std(t) + alpha(t)*x0] at fixed eps."""
    b = vp.beta(t)
    v = vp.var(t)
    std = v ** 0.5
    alpha = vp.mean_weight(t)
    return eps * b * (1 - v) / (2 * std) - 0.5 * b * alpha * x0


def score_and_fpe_terms(params, z_t, cond, t, z0, eps, need_space=True, probe=None):
    """Common jets for one batch.

    z_t  (B,d): diffused state fed to the net (x_t for CDE, [x_t,y_t] for CDiffE)
    cond (B,c) or None: undiffused conditioning columns (y for CDE, none for CDiffE)
    z0   (B,d): clean state the diffusion started from;  eps (B,d) its noise.
    probe (B,d) or None: the vector v of div_estimator (losses.py:28-40); given, the divergence is the one-sample
          Hutchinson estimate v.(J_s^T v) instead of tr J_s, as `divergence_method='hutchinson'` evaluates it.
    Returns dict with a, s, ds_dt (total), and if need_space: div_s, grad_x (detached).
    """
    B, d = z_t.shape
    c = 0 if cond is None else cond.shape[1]
    inp = torch.cat([z_t] + ([cond] if c else []) + [t.reshape(B, 1)], dim=1)
    nin = inp.shape[1]
    zeros = torch.zeros(B, nin, dtype=inp.dtype)
    # direction 0: time  (xdot, 0_cond, 1)
    dirs = [torch.cat([_xdot(t, z0, eps), torch.zeros(B, c, dtype=inp.dtype), torch.ones(B, 1, dtype=inp.dtype)], 1)]
    pairs = []
    if need_space:
        for k in range(d):
            e = zeros.clone()
            e[:, k] = 1.0
            dirs.append(e)
        if probe is None:
            pairs = [(1 + i, 1 + k) for i in range(d) for k in range(i, d)]
        else:
            dirs.append(torch.cat([probe, torch.zeros(B, nin - d, dtype=inp.dtype)], 1))   # index d + 1
            pairs = [(d + 1, 1 + k) for k in range(d)]
    a, ad, add = onets.mlp_jets(params, inp, dirs, pairs)
    b = vp.beta(t)                       # (B,1)
    sb = b ** 0.5
    s = a / sb
    dbeta = vp.BETA_MAX - vp.BETA_MIN
    out = {"a": a, "s": s, "beta": b,
           "ds_dt": ad[0] / sb - a * dbeta / (2 * b ** 1.5)}
    if need_space:
        J = torch.stack(ad[1:1 + d], dim=2)                      # J[b,i,k] = d a_i / d x_k
        gtr = torch.zeros(B, d, dtype=inp.dtype)                 # grad_x tr J   |   grad_x v.(J v)
        if probe is None:
            out["div_s"] = torch.diagonal(J, dim1=1, dim2=2).sum(1, keepdim=True) / sb
            for idx, (i, k) in enumerate(pairs):
                i -= 1
                k -= 1
                gtr[:, k] += add[idx][:, i]
                if i != k:
                    gtr[:, i] += add[idx][:, k]
        else:
            out["div_s"] = (probe * ad[1 + d]).sum(1, keepdim=True) / sb
            for k in range(d):
                gtr[:, k] = (probe * add[k]).sum(1)              # sum_j v_j d^2 a_j [v, e_k]
        JTa = torch.einsum("bik,bi->bk", J, a)
        JTx = torch.einsum("bik,bi->bk", J, z_t)
        out["grad_x"] = (gtr / sb + 2 * JTa / b + (a + JTx) / sb).detach()
    return out


def score_fpe(terms, metric="L1"):
    """ScoreFPELoss.forward, exact divergence (losses.py:77-98) -> (B,1)"""
    R = terms["ds_dt"] - 0.5 * terms["beta"] * terms["grad_x"]
    if metric == "L1":
        return R.abs().mean(1, keepdim=True)
    if metric == "L2":
        return (R ** 2).mean(1, keepdim=True)
    raise ValueError(metric)


def cscore_fpe(terms, t, eps, std, metric="L2"):
    """ConditionalScoreFPELoss.forward (losses.py:116-124) -> (B,)"""
    alpha = vp.mean_weight(t)
    u = 0.5 * eps * terms["beta"] * alpha ** 2
    r = std ** 3 * terms["ds_dt"] - u
    return (r ** 2).sum(1) if metric == "L2" else r.abs().sum(1)


def _pde(terms, t, eps, std, pde_loss, pde_metric):
    if pde_loss == "cScoreFPE":
        return cscore_fpe(terms, t, eps, std, pde_metric)
    return score_fpe(terms, pde_metric)


def _split(model_kind, x, y, t, eps):
    """Diffuse and split as CDE.train_epoch (:80-81) / CDiffE.train_epoch (:129-133) do."""
    if model_kind == "CDE":
        z0, cond = x, y
    elif model_kind == "CDiffE":
        z0, cond = torch.cat([x, y], 1), None
    else:
        raise ValueError(model_kind)
    z_t, std, _ = vp.perturb(t, z0, eps)
    return z0, cond, z_t, std


def dsm_loss(params, model_kind, x, y, t, eps):
    """The `loss_fn.name == 'DSMLoss'` branch of CDE/CDiffE.train_epoch
    (models/diffusion.py:83-85, :134-137): mean_B DSM(a/g, std, eps)."""
    z0, cond, z_t, std = _split(model_kind, x, y, t, eps)
    a = onets.mlp(params, z_t, cond, t)
    s = a / vp.beta(t) ** 0.5
    return dsm(s, std, eps).mean()


def dsm_pde_loss(params, model_kind, x, y, t, eps, lam=1.0, pde_loss="FPE", pde_metric="L1", probe=None):
    """DSM_PDELoss.forward (losses.py:143-164).  Returns (loss, info)."""
    z0, cond, z_t, std = _split(model_kind, x, y, t, eps)
    terms = score_and_fpe_terms(params, z_t, cond, t, z0, eps, need_space=(pde_loss != "cScoreFPE"), probe=probe)
    l_dsm = dsm(terms["s"], std, eps)
    l_pde = lam * _pde(terms, t, eps, std, pde_loss, pde_metric)
    loss = l_dsm.mean() + l_pde.mean()                                  # Q10
    return loss, {"PDE-Loss": l_pde.mean(), "DSM-Loss": l_dsm.mean()}


def pinn_loss(params, model_kind, x, y, t, eps, ic_target, lam=1.0, lam2=1.0,
              pde_loss="FPE", ic_metric="L1", pde_metric="L1", probe=None):
    """PINNLoss.forward (losses.py:214-242).  `ic_target` = initial_condition(x, y),
    (B,xdim) (analytic linear score, linear_problem.py:61-65, or the scatterometry
    score_posterior).  Returns (loss, info)."""
    B, xdim = x.shape
    z0, cond, z_t, std = _split(model_kind, x, y, t, eps)
    t0 = torch.zeros_like(t)
    s0 = onets.mlp(params, x, y, t0) / vp.beta(t0) ** 0.5               # :221-223
    terms = score_and_fpe_terms(params, z_t, cond, t, z0, eps, need_space=(pde_loss != "cScoreFPE"), probe=probe)
    diff = s0[:, :xdim] - ic_target
    l_ic = lam2 * ((diff ** 2).mean(1, keepdim=True) if ic_metric == "L2" else diff.abs().mean(1, keepdim=True))
    l_dsm = dsm(terms["s"], std, eps)
    l_pde = lam * _pde(terms, t, eps, std, pde_loss, pde_metric)
    loss = l_dsm.mean() + l_ic.mean() + l_pde.mean()                    # Q10
    return loss, {"PDE-Loss": l_pde.mean(), "Initial Condition": l_ic.mean(), "DSM-Loss": l_dsm.mean()}


def pinn2_loss(params, model_kind, x, y, t, eps, ic_target, lam=1.0, lam2=1.0,
               pde_loss="FPE", ic_metric="L1", probe=None):
    """PINNLoss2.forward (losses.py:263-291): PINNLoss without the data-driven DSM term —
    mean(lam2 ic + lam pde); the DSM loss is only reported ('DSM_eval').  Upstream the class
    reads an attribute it never sets (self.ic_metric, :276) — `ic_metric` here is that
    attribute; its pde losses are built with their default metrics (:256-259: ScoreFPELoss()
    'L1', ConditionalScoreFPELoss() 'L2')."""
    B, xdim = x.shape
    z0, cond, z_t, std = _split(model_kind, x, y, t, eps)
    t0 = torch.zeros_like(t)
    s0 = onets.mlp(params, x, y, t0) / vp.beta(t0) ** 0.5               # :268-270
    terms = score_and_fpe_terms(params, z_t, cond, t, z0, eps, need_space=(pde_loss != "cScoreFPE"), probe=probe)
    diff = s0[:, :xdim] - ic_target
    l_ic = lam2 * ((diff ** 2).mean(1, keepdim=True) if ic_metric == "L2" else diff.abs().mean(1, keepdim=True))
    l_pde = lam * _pde(terms, t, eps, std, pde_loss, "L2" if pde_loss == "cScoreFPE" else "L1")
    loss = l_ic.mean() + l_pde.mean()                                   # :289 (Q10 for the (B,)+(B,1) case)
    return loss, {"PDE-Loss": l_pde.mean(), "Initial Condition": l_ic.mean(),
                  "DSM_eval": dsm(terms["s"], std, eps).mean()}


def posterior_loss(prior_params, lik_params, surr_params, x, y, t, eps, lam,
                   a=scat.A_NOISE, b=scat.B_NOISE):
    """PosteriorLoss.forward (losses.py:372-386) with likelihood_target (:349-371)
    merged into one VJP through the surrogate and one through the prior net
    (SURVEY.md a15): target = (std^2 J_s^T + I) J_f^T (-a^2 v1 + v2 + a^2 v3), detached."""
    x_t, std, _ = vp.perturb(t, x, eps)
    x_t = x_t.detach().requires_grad_(True)
    s_prior = onets.mlp2(prior_params, x_t, t)
    s_lik = onets.mlp(lik_params, x_t, y, t)
    alpha = vp.mean_weight(t)
    l_prior = dsm(s_prior, std, eps)
    x0_hat = ((x_t + std ** 2 * s_prior) / alpha).detach()
    fx = scat.surrogate(surr_params, x0_hat)
    pre = (a * fx) ** 2 + b ** 2
    w = -a * a * fx / pre + (y - fx) / pre + a * a * (y - fx) ** 2 * fx / pre
    u = scat.surrogate_vjp(surr_params, x0_hat, w)
    vhp = torch.autograd.grad(s_prior, x_t, u, retain_graph=True)[0]
    target = (std ** 2 * vhp + u).detach()
    l_lik = ((alpha * s_lik - target) ** 2).sum(1)
    loss = (l_prior + lam * l_lik).mean()
    return loss, {"PriorLoss": l_prior.mean(), "LikelihoodLoss": lam * l_lik.mean()}
