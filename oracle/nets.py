"""Score networks — oracle restatement of nets.py:17-57,143-157 (TEST INFRASTRUCTURE ONLY).

Parameters are a list [(W0,b0),(W1,b1),...] in layer order; in a reference
state_dict they sit under keys 0,3,5,7 because the activation is
registered a second time under the key 'act' (nets.py:25-30).  That same quirk
makes `Sequential.forward` apply tanh TWICE after layer 0.
"""
import torch


def params_from_state_dict(sd, prefix=""):
    """Collect (weight, bias) pairs in numeric key order (0,3,5,7 for 3 hidden layers)."""
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in sd
                  if k.startswith(prefix) and k[len(prefix):].split(".")[0].isdigit()})
    return [(sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"]) for i in idx]


def mlp_apply(params, inp):
    """nn.Sequential over (L0, tanh, tanh[='act'], L3, tanh, L5, tanh, L7) — nets.py:25-30."""
    W, b = params[0]
    h = torch.tanh(torch.tanh(inp @ W.T + b))
    for W, b in params[1:-1]:
        h = torch.tanh(h @ W.T + b)
    W, b = params[-1]
    return h @ W.T + b


def mlp(params, x, cond, t):
    """MLP.forward: cat[x, cond, t] (nets.py:32-35).  `cond` may be the empty 1-D
    tensor, which torch.cat skips (losses.py:149,219)."""
    parts = [x] + ([cond] if cond is not None and cond.numel() > 0 else []) + [t.reshape(len(x), 1)]
    return mlp_apply(params, torch.cat(parts, dim=1))


def mlp2(params, x, t):
    """MLP2.forward: cat[x, t] (nets.py:52-57)."""
    return mlp_apply(params, torch.cat([x, t.reshape(len(x), 1)], dim=1))


def posterior_score(prior_params, lik_params, x, y, t, g):
    """PosteriorScore.forward (nets.py:155-157): g(t,x) * (prior(x,t) + likelihood(x,y,t))."""
    return g * (mlp2(prior_params, x, t) + mlp(lik_params, x, y, t))


# ---------------------------------------------------------------------------
# Forward-mode jets through the MLP.  Used by the loss
# oracle; written with torch ops so autograd yields parameter gradients.
# ---------------------------------------------------------------------------

def mlp_jets(params, inp, d1, pairs=()):
    """Propagate value, first-order tangents and second-order tangents.

    inp   : (B, in)
    d1    : list of (B, in) input directions u_k
    pairs : list of (i, j) index pairs into d1 — second-order directional
            derivative d^2/ds dr f(inp + s u_i + r u_j) (input second tangent is 0)
    returns (out, [out_dot_k], [out_ddot_(i,j)])
    """
    n_layers = len(params)
    h = inp
    hd = list(d1)
    hdd = [torch.zeros_like(inp) for _ in pairs]
    for li, (W, b) in enumerate(params):
        z = h @ W.T + b
        zd = [v @ W.T for v in hd]
        zdd = [v @ W.T for v in hdd]
        if li == n_layers - 1:
            return z, zd, zdd
        if li == 0:
            u = torch.tanh(z)
            h = torch.tanh(u)
            p1 = (1 - h * h) * (1 - u * u)                      # phi'
            p2 = p1 * (-2 * h * (1 - u * u) - 2 * u)            # phi''
        else:
            h = torch.tanh(z)
            p1 = 1 - h * h
            p2 = -2 * h * p1
        hdd = [p1 * zdd[k] + p2 * zd[i] * zd[j] for k, (i, j) in enumerate(pairs)]
        hd = [p1 * v for v in zd]
    raise AssertionError
