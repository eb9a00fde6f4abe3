"""Shim for the un-vendored `include.sdeflow_light.lib.utils` (sdes.py:6).

The reference imports three helpers from CW-Huang/sdeflow-light, which its
author vendored under a git-ignored `include/` directory (absent from
/root/reference).  Only `sample_vp_truncated_q` is on a live path
(sdes.py:57 <- models/diffusion.py:51): it draws the training time `t`,
which is an *input* of every loss kernel, so parity tests inject `t` and this
shim never decides a parity result.  PARITY UNPINNED at this boundary: the
formula below restates the published sdeflow-light inverse-CDF sampler
(SURVEY.md App. A.2), not the author's vendored copy.

Test infrastructure only.
"""
import math

import torch


def _B(t, bmin, bmax):
    return 0.5 * t * t * (bmax - bmin) + t * bmin


def _A(t, bmin, bmax):
    # antiderivative of beta/var:  log(1 - exp(-B)) + B
    b = _B(t, bmin, bmax)
    return math.log(1.0 - math.exp(-b)) + b


def vp_truncated_inv_cdf(u, beta_min, beta_max, t_epsilon, T):
    """t = Phi^{-1}(u) for the density  q(t) ∝ beta(t)/var(t)  truncated below t_epsilon."""
    db = beta_max - beta_min
    r_eps = (beta_min + db * t_epsilon) / (1.0 - math.exp(-_B(t_epsilon, beta_min, beta_max)))
    a_eps = _A(t_epsilon, beta_min, beta_max)
    Z = r_eps * t_epsilon + _A(float(T), beta_min, beta_max) - a_eps
    low = (u <= t_epsilon * r_eps / Z).to(u.dtype)
    lin = Z / r_eps * u
    arg = torch.log(1.0 + torch.exp(Z * u + a_eps - r_eps * t_epsilon))
    nl = (-beta_min + (beta_min ** 2 + 2.0 * db * arg) ** 0.5) / db
    return low * lin + (1.0 - low) * nl


def sample_vp_truncated_q(shape, beta_min, beta_max, t_epsilon, T):
    u = torch.rand(*shape)
    return vp_truncated_inv_cdf(u.view(-1), beta_min, beta_max, t_epsilon, T).view(*shape)


def sample_v(shape, vtype="rademacher"):
    if vtype == "rademacher":
        return torch.randint(0, 2, shape).float() * 2.0 - 1.0
    return torch.randn(shape)


def log_normal(x, mean, log_var, eps=1e-5):
    return -0.5 * (x - mean) ** 2 / (log_var.exp() + eps) - 0.5 * log_var - 0.5 * math.log(2 * math.pi)
