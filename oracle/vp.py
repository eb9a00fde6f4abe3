"""VP-SDE closed forms — oracle restatement of sdes.py:9-57 (TEST INFRASTRUCTURE ONLY).

beta_min=0.1, beta_max=20, T=1, t_epsilon=1e-3 are the reference defaults
(sdes.py:14-19); no caller ever overrides them.
"""
import torch

BETA_MIN = 0.1
BETA_MAX = 20.0
T = 1.0
T_EPS = 1e-3


def beta(t, bmin=BETA_MIN, bmax=BETA_MAX):
    """sdes.py:21-22"""
    return bmin + (bmax - bmin) * t


def mean_weight(t, bmin=BETA_MIN, bmax=BETA_MAX):
    """alpha(t), sdes.py:24-25"""
    return torch.exp(-0.25 * t ** 2 * (bmax - bmin) - 0.5 * t * bmin)


def var(t, bmin=BETA_MIN, bmax=BETA_MAX):
    """sdes.py:27-28"""
    return 1.0 - torch.exp(-0.5 * t ** 2 * (bmax - bmin) - t * bmin)


def f(t, y):
    """forward drift, sdes.py:30-31"""
    return -0.5 * beta(t) * y


def g(t, y):
    """forward diffusion broadcast to y.shape, sdes.py:33-35"""
    return torch.ones_like(y) * beta(t) ** 0.5


def perturb(t, y0, eps):
    """`VariancePreservingSDE.sample(t, y0, return_noise=True)` with the Gaussian
    draw injected (sdes.py:37-49).  Returns (y_t, std, g)."""
    mu = mean_weight(t) * y0
    std = var(t) ** 0.5
    yt = eps * std + mu
    return yt, std, g(t, yt)


def sample_t_from_uniform(u, eps=1e-4, T=T):
    """`BaseClassDiffusionModel.sample_t` debias branch (models/diffusion.py:48-53)
    with the uniform draw `u` injected; inverse CDF per SURVEY.md App. A.2
    (sdeflow-light, parity unpinned — see oracle/shims)."""
    from .shims.include.sdeflow_light.lib.utils import vp_truncated_inv_cdf
    t = vp_truncated_inv_cdf(u.reshape(-1), BETA_MIN, BETA_MAX, T_EPS, T).reshape(u.shape) + eps
    t = torch.where(t > T, t - eps, t)
    return t
