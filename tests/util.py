"""Shared helpers for the test-suite (fixtures, parameter regeneration)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name, dtype=torch.float32):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        v = z[k]
        if k.startswith("gidx") or "_gidx_" in k or k == "meta":
            out[k] = torch.from_numpy(v.astype(np.int64))
        else:
            npd = np.float64 if dtype == torch.float64 else np.float32      # float64 tables keep their precision
            out[k] = torch.from_numpy(np.asarray(v, dtype=npd)).to(dtype)
    return out


def meta_hidden(fx, n_head):
    return tuple(int(v) for v in fx["meta"][n_head:])


def surrogate_params(dtype=torch.float32):
    fx = load_golden("surrogate", dtype)
    return [(fx[f"{i}_weight"], fx[f"{i}_bias"]) for i in (0, 2, 4, 6)]


KEYS = {4: (0, 3, 5, 7), 3: (0, 3, 5)}


def check_grads(fx, grads, prefix="", rtol=1e-3, atol_frac=1e-3):
    """Compare a list of (W,b) gradients with the fixture summary written by make_golden.py."""
    keys = KEYS[len(grads)]
    for k, (gW, gb) in zip(keys, grads):
        for nm, g in ((f"{k}_weight", gW), (f"{k}_bias", gb)):
            ref_norm = float(fx[f"{prefix}gnorm_{nm}"])
            scale = ref_norm / max(g.numel() ** 0.5, 1.0)
            if f"{prefix}g_{nm}" in fx:
                ref = fx[f"{prefix}g_{nm}"]
                got = g
            else:
                idx = fx[f"{prefix}gidx_{nm}"]
                ref = fx[f"{prefix}gval_{nm}"]
                got = g.reshape(-1)[idx]
            err = (got.double() - ref.double()).abs().max().item()
            assert err <= atol_frac * scale + rtol * ref.abs().max().item(), (nm, err, scale)
            assert abs(float(g.norm()) - ref_norm) <= rtol * ref_norm + 1e-12, (nm, float(g.norm()), ref_norm)
