"""Deterministic synthetic parameters for fixtures (TEST INFRASTRUCTURE ONLY).

numpy PCG64 is bit-stable across platforms, so fixtures need not store the
2 MB weight tensors: both `make_golden.py` (reference side) and the tests
(oracle / CUDA side) regenerate them from (seed, dims).  Scale follows
nn.Linear's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)) times `gain`.
"""
import numpy as np
import torch


def make_params(seed, in_dim, out_dim, hidden=(512, 512, 512), gain=1.0, dtype=torch.float32):
    rng = np.random.Generator(np.random.PCG64(seed))
    dims = [in_dim] + list(hidden) + [out_dim]
    params = []
    for fi, fo in zip(dims[:-1], dims[1:]):
        bound = gain / np.sqrt(fi)
        W = rng.uniform(-bound, bound, size=(fo, fi)).astype(np.float32)
        b = rng.uniform(-bound, bound, size=(fo,)).astype(np.float32)
        params.append((torch.from_numpy(W).to(dtype), torch.from_numpy(b).to(dtype)))
    return params


def state_dict_from_params(params):
    """Reference key naming 0,3,5,7,...."""
    keys = [0] + [3 + 2 * i for i in range(len(params) - 1)]
    sd = {}
    for k, (W, b) in zip(keys, params):
        sd[f"{k}.weight"] = W.clone()
        sd[f"{k}.bias"] = b.clone()
    return sd
