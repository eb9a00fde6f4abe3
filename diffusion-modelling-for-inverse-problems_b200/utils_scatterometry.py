"""Drop-in for the hot functions of the reference's `utils_scatterometry.py` (load_forward_model :8-27,
get_log_posterior :30-38) together with `energy_grad` of models/SNF.py:234-237 as they are composed into
`score_posterior` by main_diffusion_scatterometry.py:142-145.

The surrogate is frozen, so energy, its gradient and the posterior score are ONE fused call into libdmip_sm100.so
(`dmip_surrogate_score`, include/dmip.h): forward through the ReLU MLP, closed-form d energy / d f,
reverse sweep through the stored ReLU masks — instead of the reference's autograd graph.
"""
import ctypes as C
import os

import torch
from torch import nn

from . import _lib

device = 'cuda' if torch.cuda.is_available() else 'cpu'

SURR_ENERGY, SURR_LIK_VJP = 0, 1


class DmipSurrogate(C.Structure):
    _fields_ = [("mode", C.c_int32), ("net", _lib.DmipMlp), ("a", C.c_float), ("b", C.c_float), ("lambd_bd", C.c_float),
                ("n", C.c_int64), ("x", C.c_void_p), ("y", C.c_void_p), ("energy", C.c_void_p), ("grad", C.c_void_p),
                ("fx", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("rows_per_obs", C.c_int64)]


def _bind():
    L = _lib.require_gpu()
    if not getattr(L, "_surr_bound", False):
        L.dmip_surrogate_workspace_bytes.restype = C.c_size_t
        L.dmip_surrogate_workspace_bytes.argtypes = [C.POINTER(DmipSurrogate)]
        L.dmip_surrogate_score.restype = C.c_int
        L.dmip_surrogate_score.argtypes = [C.POINTER(DmipSurrogate), C.c_void_p]
        L._surr_bound = True
    return L


def load_forward_model(src_dir):
    """Same module, checkpoint name and parameter dict as the reference (utils_scatterometry.py:8-27)."""
    forward_model = nn.Sequential(nn.Linear(3, 256), nn.ReLU(),
                                  nn.Linear(256, 256), nn.ReLU(),
                                  nn.Linear(256, 256), nn.ReLU(),
                                  nn.Linear(256, 23)).to(device)
    forward_model.load_state_dict(torch.load(os.path.join(src_dir, 'surrogate.pt'), map_location=torch.device(device)))
    for param in forward_model.parameters():
        param.requires_grad = False
    params = {'a': 0.2, 'b': 0.01, 'lambd_bd': 1000, 'xdim': 3, 'ydim': 23}
    return forward_model, params


def surrogate_call(forward_model, x, y, a, b, lambd_bd=0.0, mode=SURR_ENERGY, want_fx=False):
    """One fused forward + reverse sweep.  Returns (energy | None, grad, fx | None)."""
    L = _bind()
    if not x.is_cuda:
        raise RuntimeError("dmip surrogate kernels run on CUDA (sm_100a) only: there is no CPU fallback")
    x = x.detach().float().contiguous()
    n = x.shape[0]
    y = torch.as_tensor(y, dtype=torch.float32, device=x.device)
    y = y.reshape(-1, y.shape[-1]).contiguous()
    # one observation per row, or k observations for k equal blocks of rows (k = 1: the reference's broadcast of one
    # observation over all samples) — the kernel indexes y by row / rows_per_obs, nothing is expanded
    k = y.shape[0]
    if k == n:
        rows_per_obs = 0
    elif n % k == 0:
        rows_per_obs = n // k
    else:
        raise ValueError(f"y has {k} rows for {n} samples: need one row per sample or a divisor of the sample count")
    keep = [x, y]
    d = DmipSurrogate()
    d.mode = mode
    d.net = _lib.mlp_desc(forward_model, keep)
    d.a, d.b, d.lambd_bd = float(a), float(b), float(lambd_bd)
    d.n = n
    d.rows_per_obs = rows_per_obs
    d.x, d.y = x.data_ptr(), y.data_ptr()
    guard = _lib.Guarded()                 # DMIP_GUARD=1: canaries around every buffer the kernels write (debugging)
    energy = guard.empty(n, torch.float32, x.device) if mode == SURR_ENERGY else None
    grad = guard.empty(x.numel(), torch.float32, x.device).view_as(x)
    fx = guard.empty(n * d.net.out_dim, torch.float32, x.device).view(n, d.net.out_dim) if want_fx else None
    d.energy = energy.data_ptr() if energy is not None else None
    d.grad = grad.data_ptr()
    d.fx = fx.data_ptr() if fx is not None else None
    ws = guard.empty(max(L.dmip_surrogate_workspace_bytes(C.byref(d)), 16), torch.uint8, x.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(x.device):
        _lib.check(L.dmip_surrogate_score(C.byref(d), _lib.stream_ptr()))
    guard.check("dmip_surrogate_score")
    surrogate_call.last_launch_count = L.dmip_last_launch_count()
    return energy, grad, fx


def get_log_posterior(samples, forward_model, a, b, ys, lambd_bd):
    """Negative log posterior E(x) (utils_scatterometry.py:30-38) -> (n,)."""
    return surrogate_call(forward_model, samples, ys, a, b, lambd_bd)[0]


def energy_and_grad(samples, forward_model, a, b, ys, lambd_bd):
    """(grad_x E, E) — what `energy_grad(x, lambda x: get_log_posterior(...))` returns (models/SNF.py:234-237)."""
    e, g, _ = surrogate_call(forward_model, samples, ys, a, b, lambd_bd)
    return g, e


def make_score_posterior(forward_model, forward_model_params):
    """score_posterior(x, y) = -grad_x E, the PINNLoss initial condition (main_diffusion_scatterometry.py:142-145)."""
    p = forward_model_params

    def score_posterior(x, y):
        return -surrogate_call(forward_model, x, y, p['a'], p['b'], p['lambd_bd'])[1]

    return score_posterior
