"""Scatterometry surrogate, energy and posterior score — oracle restatement of
utils_scatterometry.py:8-38 and models/SNF.py:234-237 (TEST INFRASTRUCTURE ONLY).

Closed forms: SURVEY.md App. A.6.
"""
import torch

A_NOISE = 0.2       # utils_scatterometry.py:19
B_NOISE = 0.01      # :20
LAMBD_BD = 1000.0   # :21


def surrogate_params_from_state_dict(sd):
    """nn.Sequential(Linear(3,256),ReLU,Linear(256,256),ReLU,Linear(256,256),ReLU,Linear(256,23))
    — keys 0,2,4,6 (utils_scatterometry.py:9-12)."""
    return [(sd[f"{i}.weight"], sd[f"{i}.bias"]) for i in (0, 2, 4, 6)]


def surrogate(params, x):
    h = x
    for W, b in params[:-1]:
        h = torch.relu(h @ W.T + b)
    W, b = params[-1]
    return h @ W.T + b


def energy(params, x, y, a=A_NOISE, b=B_NOISE, lambd_bd=LAMBD_BD):
    """get_log_posterior (utils_scatterometry.py:30-38): the *negative* log posterior E(x)."""
    fx = surrogate(params, x)
    pre = (a * fx) ** 2 + b ** 2
    p = 0.5 * torch.sum(torch.log(pre), dim=1)
    p2 = 0.5 * torch.sum((y - fx) ** 2 / pre, dim=1)
    p3 = lambd_bd * torch.sum(torch.relu(x - 1) + torch.relu(-1 - x), dim=1)
    return p + p2 + p3


def surrogate_vjp(params, x, v):
    """J_f(x)^T v by an explicit reverse sweep through the ReLU masks."""
    hs = [x]
    for W, b in params[:-1]:
        hs.append(torch.relu(hs[-1] @ W.T + b))
    gbar = v @ params[-1][0]
    for (W, _), h in zip(reversed(params[:-1]), reversed(hs[1:])):
        gbar = (gbar * (h > 0).to(gbar.dtype)) @ W
    return gbar


def energy_and_grad(params, x, y, a=A_NOISE, b=B_NOISE, lambd_bd=LAMBD_BD):
    """(E, grad_x E) in closed form; equals `energy_grad`
    (models/SNF.py:234-237) applied to get_log_posterior."""
    fx = surrogate(params, x)
    pre = (a * fx) ** 2 + b ** 2
    r = y - fx
    E = 0.5 * torch.sum(torch.log(pre), 1) + 0.5 * torch.sum(r * r / pre, 1) \
        + lambd_bd * torch.sum(torch.relu(x - 1) + torch.relu(-1 - x), 1)
    dE_df = a * a * fx / pre - r / pre - a * a * fx * r * r / (pre * pre)
    gx = surrogate_vjp(params, x, dE_df) + lambd_bd * ((x > 1).to(x.dtype) - (x < -1).to(x.dtype))
    return E, gx


def score_posterior(params, x, y):
    """Initial condition of the scatterometry PINNLoss
    (main_diffusion_scatterometry.py:142-145): -grad_x E."""
    return -energy_and_grad(params, x, y)[1]
