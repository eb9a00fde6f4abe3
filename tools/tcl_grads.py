"""Per-parameter comparison of the tcgen05 loss path against the FFMA path and the fixture (debug aid).
Usage: python tools/tcl_grads.py fixture [fixture ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import gpu_cases as gc
from dmip import losses as dl
from oracle.weights import make_params, state_dict_from_params
from util import load_golden, meta_hidden


def run(name, path):
    os.environ["DMIP_LOSS_PATH"] = path
    model_kind, kind, kw, gain = gc.LOSS_CASES[name]
    fx = load_golden(name)
    seed, xdim, ydim, B = (int(v) for v in fx["meta"][:4])
    hidden = meta_hidden(fx, 4)
    from dmip.models.diffusion import CDE, CDiffE
    m = (CDE if model_kind == "CDE" else CDiffE)(xdim, ydim, list(hidden))
    out_dim = xdim if model_kind == "CDE" else xdim + ydim
    m.sde.a.load_state_dict(state_dict_from_params(make_params(seed, xdim + ydim + 1, out_dim, hidden, gain=gain)))
    m.sde.to("cuda")
    x, y, t, eps = (fx[k].to("cuda") for k in ("x", "y", "t", "eps"))
    if kind == "DSM":
        loss, _ = dl.dsm_fused(m, x, y, t, eps)
        info = {}
    else:
        if kind == "PINN":
            ic = fx["ic_target"].to("cuda")
            loss_fn = dl.PINNLoss(lambda xx, yy: ic, **kw)
        else:
            loss_fn = dl.DSM_PDELoss(**kw)
        if "probe" in fx:
            loss_fn.probe = fx["probe"].to("cuda")
        z = x if model_kind == "CDE" else torch.cat([x, y], 1)
        loss, info = loss_fn(m.sde, x, y, z, t, eps, None, None)
    m.sde.a.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), {k: v.item() for k, v in info.items()}, {n: p.grad.detach().cpu().clone() for n, p in m.sde.a.named_parameters()}


for name in sys.argv[1:]:
    la, ia, ga = run(name, "ffma")
    lb, ib, gb = run(name, "tc")
    print(f"== {name}: loss ffma {la:.7g} tc {lb:.7g}  info ffma {ia} tc {ib}")
    for k in ga:
        a, b = ga[k], gb[k]
        d = (a - b).abs()
        print(f"   {k:10s} shape {tuple(a.shape)}  |ffma| {a.norm():.4e} |tc| {b.norm():.4e}  max|diff| {d.max():.3e}  rel {d.max() / a.abs().max():.3e}"
              f"  cos {torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item():.6f}")
        if a.ndim == 2 and d.max() / a.abs().max() > 1e-2:
            rows = (d.max(1).values / a.abs().max() > 1e-2).nonzero().flatten()
            cols = (d.max(0).values / a.abs().max() > 1e-2).nonzero().flatten()
            print(f"      bad rows {len(rows)} {rows[:12].tolist()}  bad cols {len(cols)} {cols[:12].tolist()}")
            r0 = int(rows[0])
            print("      ffma row", r0, a[r0, :6].tolist())
            print("      tc   row", r0, b[r0, :6].tolist())
