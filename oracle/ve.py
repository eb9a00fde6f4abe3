"""VE-SDE closed forms and the predictor–corrector sampler — oracle restatement of what BASELINE.json calls "VP/VE SDE
drift/diffusion" and "predictor sampler" (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the reference ships only the VP-SDE (sdes.py:9-57) and only the Euler–Maruyama predictor
(models/diffusion.py:27-46); there is no upstream VE class, no corrector and therefore no golden vector.  This module
restates Song et al. 2021 ("Score-based generative modeling through SDEs", eq. 30-31, Alg. 4/5) behind the interface of
the reference's `VariancePreservingSDE`, and is what the CUDA samplers are checked against for these modes.

    sigma(t) = smin (smax / smin)^t,  f = 0,  g(t) = sigma(t) sqrt(2 ln(smax / smin)),  x_t | x_0 ~ N(x_0, sigma(t)^2)
Corrector (n_corr Langevin sub-steps after every predictor step, at the time level the predictor reached):
    x <- x + e s + sqrt(2 e) z,   s = a / g,   e = 2 snr^2 std(t)^2
— the closed form of Song's step-size rule e = 2 (snr |z| / |s|)^2 for a calibrated score (|z| ~ sqrt(d), |s| ~
sqrt(d) / std); no norms over coordinates or batch are taken, so particles stay independent (csrc/dmip_sde.cuh).
"""
import math

import torch

from . import vp
from . import nets as onets


class VE:
    kind = "VE"

    def __init__(self, sigma_min=0.01, sigma_max=50.0, T=1.0):
        self.smin, self.smax, self.T = sigma_min, sigma_max, T

    def sigma(self, t):
        return self.smin * (self.smax / self.smin) ** t

    def mean_weight(self, t):
        return torch.ones_like(t)

    def var(self, t):
        return self.sigma(t) ** 2

    def g(self, t, y):
        return torch.ones_like(y) * self.sigma(t) * math.sqrt(2.0 * math.log(self.smax / self.smin))

    def f(self, t, y):
        return torch.zeros_like(y)


class VP:
    """the reference's SDE (oracle/vp.py) behind the same small interface"""
    kind = "VP"

    def __init__(self, T=vp.T):
        self.T = T

    mean_weight = staticmethod(vp.mean_weight)
    var = staticmethod(vp.var)
    g = staticmethod(vp.g)
    f = staticmethod(vp.f)


def _time_grid(num_steps, T, dtype):
    return (torch.linspace(0, 1, num_steps + 1) * T).to(dtype)


def pc_sampler(net, variant, y, x0, noise, num_steps, sde, n_corr=0, snr=0.16, ynoise=None):
    """Predictor–corrector reverse-SDE sampler.

    net(x, ycond, tau) -> net output: the drift a = g * score for variant 'CDE' / 'CDiffE' (models/diffusion.py:27-46,
    :158-180), the score sum prior + likelihood for 'DPS' (nets.py:155-157: a = g * sum).
    noise (S * (1 + n_corr), N, xdim): sub-step u = step * (1 + n_corr) + c uses noise[u]; c = 0 is the predictor.
    CDiffE: ynoise (S * (1 + n_corr), N, ydim) re-diffuses y for every net evaluation.
    """
    N, xdim = x0.shape
    T = sde.T
    ys = torch.zeros(N, y.numel(), dtype=x0.dtype) + y
    delta = T / num_steps
    ts = _time_grid(num_steps, T, x0.dtype)
    ones = torch.ones(N, 1, dtype=x0.dtype)
    x = x0
    per = 1 + n_corr
    for u in range(num_steps * per):
        i, c = divmod(u, per)
        tau = T - ones * ts[i if c == 0 else i + 1]
        if variant == "CDiffE":
            y_t = ynoise[u] * sde.var(tau) ** 0.5 + sde.mean_weight(tau) * ys
            out = net(torch.cat([x, y_t], 1), None, tau)[:, :xdim]
        else:
            out = net(x, ys, tau)
        g = sde.g(tau, x)
        a = g * out if variant == "DPS" else out
        if c == 0:                                   # Euler–Maruyama predictor (sdes.py:77-87)
            x = x + delta * (g * a - sde.f(tau, x)) + delta ** 0.5 * g * noise[u]
        else:                                        # Langevin corrector
            e = 2.0 * snr * snr * sde.var(tau)
            x = x + e * (a / g) + (2.0 * e) ** 0.5 * noise[u]
    return x


def net_fn(variant, params, params2=None):
    """closures over the oracle nets: CDE -> mlp(x, y, t); CDiffE -> mlp([x, y_t], -, t); DPS -> prior(x, t) + lik(x, y, t)"""
    empty = torch.zeros(0)
    if variant == "CDE":
        return lambda x, c, t: onets.mlp(params, x, c, t)
    if variant == "CDiffE":
        return lambda z, c, t: onets.mlp(params, z, empty.to(z.dtype), t)
    return lambda x, c, t: onets.mlp2(params2, x, t) + onets.mlp(params, x, c, t)
