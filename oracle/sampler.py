"""Reverse-SDE Euler–Maruyama samplers — oracle restatement of
models/diffusion.py:27-46 (CDE / DPS), :158-180 (CDiffE, with the one-argument
fix of SURVEY.md Q7) and sdes.py:77-87.  TEST INFRASTRUCTURE ONLY.

All Gaussian draws are injected so the CUDA path can be fed identical noise.
RNG order of the reference: x0 = randn(N,xdim); then per step one
randn_like(x_t).  CDiffE per step: randn_like(z_0) (for y_t) then randn_like(z_t).
"""
import torch

from . import vp
from . import nets as onets


def _time_grid(num_steps, T, dtype):
    # models/diffusion.py:34 — linspace is fp32 on CPU in the reference
    return (torch.linspace(0, 1, num_steps + 1) * T).to(dtype)


def em_sampler(drift_a, y, x0, noise, num_steps, T=vp.T):
    """BaseClassDiffusionModel.forward (models/diffusion.py:27-46).

    drift_a(x, ycond, t) -> (N, xdim)   [the net output a = g * score, SURVEY.md Q3]
    y (ydim,), x0 (N, xdim) already scaled by std/mean, noise (S, N, xdim).
    """
    N = x0.shape[0]
    ys = torch.zeros(N, y.numel(), dtype=x0.dtype) + y
    delta = T / num_steps
    ts = _time_grid(num_steps, T, x0.dtype)
    ones = torch.ones(N, 1, dtype=x0.dtype)
    x = x0
    for i in range(num_steps):
        t = ones * ts[i]
        tau = T - t
        gg = vp.g(tau, x)
        mu = gg * drift_a(x, ys, tau) - vp.f(tau, x)       # sdes.py:77-79, lmbd = 0
        sigma = vp.g(tau, x)                                # sdes.py:86-87
        x = x + delta * mu + delta ** 0.5 * sigma * noise[i]
    return x


def em_sampler_cde(params, y, x0, noise, num_steps, T=vp.T):
    return em_sampler(lambda x, c, t: onets.mlp(params, x, c, t), y, x0, noise, num_steps, T)


def em_sampler_dps(prior_params, lik_params, y, x0, noise, num_steps, T=vp.T):
    def a(x, c, t):
        return onets.posterior_score(prior_params, lik_params, x, c, t, vp.g(t, x))
    return em_sampler(a, y, x0, noise, num_steps, T)


def em_sampler_cdiffe(params, y, x0, ynoise, noise, num_steps, T=vp.T):
    """CDiffE.forward (models/diffusion.py:158-180) with `cond = torch.Tensor([])`
    passed to `mu` (the upstream call omits it and raises TypeError).

    ynoise (S, N, ydim): the y-columns of randn_like(z_0) used to re-diffuse y.
    noise  (S, N, xdim): the x-columns of randn_like(z_t); the y-columns of the
                         update are discarded every step (:177).
    """
    N, xdim = x0.shape
    ydim = y.numel()
    ys = torch.zeros(N, ydim, dtype=x0.dtype) + y
    delta = T / num_steps
    ts = _time_grid(num_steps, T, x0.dtype)
    ones = torch.ones(N, 1, dtype=x0.dtype)
    empty = torch.zeros(0, dtype=x0.dtype)
    x = x0
    for i in range(num_steps):
        tau0 = T - ts[i]                                            # 0-dim, :172
        y_t = ynoise[i] * vp.var(tau0) ** 0.5 + vp.mean_weight(tau0) * ys
        z = torch.cat([x, y_t], dim=1)
        tau = T - ones * ts[i]
        a = onets.mlp(params, z, empty, tau)                        # out dim xdim+ydim
        mu = vp.g(tau, z) * a - vp.f(tau, z)
        x = (z + delta * mu)[:, :xdim] + delta ** 0.5 * vp.g(tau, x) * noise[i]
    return x
