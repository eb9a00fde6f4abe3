"""Ground-truth posterior samples for scatterometry: the random-walk Metropolis chains of the reference
(`anneal_to_energy`, models/SNF.py:250-275, driven by generate_scatterometry_ground_truth.py:26-63) as ONE kernel
launch for all chains, all observations and all steps (`dmip_metropolis`, include/dmip.h).

Upstream the target energy is a Python closure over `get_log_posterior`; the fused path needs the pieces of that
closure instead: the surrogate, (a, b, lambd_bd) and the observation(s).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib


class DmipMetropolis(C.Structure):
    _fields_ = [("net", _lib.DmipMlp), ("a", C.c_float), ("b", C.c_float), ("lambd_bd", C.c_float),
                ("noise_std", C.c_float), ("n_obs", C.c_int32), ("n_per_obs", C.c_int64), ("steps", C.c_int32),
                ("y", C.c_void_p), ("x", C.c_void_p), ("de", C.c_void_p), ("rng_mode", C.c_int32),
                ("seed", C.c_uint64), ("gidx_base", C.c_uint64), ("noise", C.c_void_p), ("unif", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


def _bind():
    L = _lib.require_gpu()
    if not getattr(L, "_mcmc_bound", False):
        L.dmip_metropolis_workspace_bytes.restype = C.c_size_t
        L.dmip_metropolis_workspace_bytes.argtypes = [C.POINTER(DmipMetropolis)]
        L.dmip_metropolis.restype = C.c_int
        L.dmip_metropolis.argtypes = [C.POINTER(DmipMetropolis), C.c_void_p]
        L._mcmc_bound = True
    return L


_seed_counter = [0]


def anneal_to_energy(x_curr, forward_model, a, b, ys, lambd_bd, metr_steps_per_block, noise_std=0.1, *, seed=None,
                     injected=None, gidx_base=0):
    """`anneal_to_energy(x_curr, energy, metr_steps_per_block, noise_std)` with
    energy = get_log_posterior(., forward_model, a, b, ys, lambd_bd).  x_curr (n, xdim) CUDA tensor; ys (ydim,) one
    observation for all chains, or (n_obs, ydim) with n = n_obs * n_per_obs chains grouped by observation.
    Returns (x_final, E(x_final) - E(x_start)) like upstream.  injected={'noise': (S, n, xdim), 'unif': (S, n)}
    replays a given random stream (parity tests); otherwise Philox keyed by (seed, gidx_base + chain, step)."""
    L = _bind()
    if not x_curr.is_cuda:
        raise RuntimeError("dmip Metropolis chains run on CUDA (sm_100a) only: there is no CPU fallback")
    dev = x_curr.device
    x = x_curr.detach().to(torch.float32).contiguous().clone()
    n = x.shape[0]
    ys = torch.as_tensor(ys, dtype=torch.float32, device=dev).reshape(-1, ys.shape[-1]).contiguous()
    n_obs = ys.shape[0]
    if n % n_obs:
        raise ValueError("the number of chains must be a multiple of the number of observations")
    keep = [x, ys]
    d = DmipMetropolis()
    d.net = _lib.mlp_desc(forward_model, keep)
    d.a, d.b, d.lambd_bd, d.noise_std = float(a), float(b), float(lambd_bd), float(noise_std)
    d.n_obs, d.n_per_obs, d.steps = n_obs, n // n_obs, int(metr_steps_per_block)
    de = torch.empty(n, device=dev, dtype=torch.float32)
    d.y, d.x, d.de = ys.data_ptr(), x.data_ptr(), de.data_ptr()
    if injected is not None:
        d.rng_mode = _lib.RNG_INJECTED
        for name in ('noise', 'unif'):
            tns = injected[name].to(dev, torch.float32).contiguous()
            keep.append(tns)
            setattr(d, name, tns.data_ptr())
        assert injected['noise'].numel() == d.steps * n * x.shape[1] and injected['unif'].numel() == d.steps * n
    else:
        d.rng_mode = _lib.RNG_PHILOX
        if seed is None:
            _seed_counter[0] += 1
            seed = (torch.initial_seed() + 0xD1B54A32D192ED03 * _seed_counter[0]) & (2 ** 64 - 1)
        d.seed = int(seed) & (2 ** 64 - 1)
    d.gidx_base = int(gidx_base)
    ws = torch.empty(max(L.dmip_metropolis_workspace_bytes(C.byref(d)), 16), dtype=torch.uint8, device=dev)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(dev):
        _lib.check(L.dmip_metropolis(C.byref(d), _lib.stream_ptr()))
    anneal_to_energy.last_launch_count = L.dmip_last_launch_count()
    return x, de


def generate_gt_samples(forward_model, forward_model_params, ys, n_samples_x, metr_steps, noise_std, n_repeats=10,
                        gt_dir=None, seed=None):
    """generate_scatterometry_ground_truth.py:26-63 for all test observations at once: per repeat j one launch of
    n_obs * n_samples_x chains started at U(-1,1)^xdim.  Returns a CUDA tensor (n_obs, n_repeats, n_samples_x, xdim);
    with gt_dir also writes <gt_dir>/<i>/<j>.npy, the layout `evaluate` reads upstream."""
    p = forward_model_params
    dev = next(forward_model.parameters()).device
    ys = torch.as_tensor(ys, dtype=torch.float32, device=dev).reshape(-1, p['ydim'])
    n_obs = ys.shape[0]
    out = torch.empty(n_obs, n_repeats, n_samples_x, p['xdim'], device=dev)
    for j in range(n_repeats):
        x0 = torch.rand(n_obs * n_samples_x, p['xdim'], device=dev) * 2 - 1
        x, _ = anneal_to_energy(x0, forward_model, p['a'], p['b'], ys, p['lambd_bd'], metr_steps, noise_std,
                                seed=None if seed is None else seed + j)
        out[:, j] = x.view(n_obs, n_samples_x, -1)
    if gt_dir is not None:
        host = out.cpu().numpy()
        for i in range(n_obs):
            os.makedirs(os.path.join(gt_dir, str(i)), exist_ok=True)
            for j in range(n_repeats):
                np.save(os.path.join(gt_dir, str(i), '%d.npy' % j), host[i, j])
    return out
