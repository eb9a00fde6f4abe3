"""Training-step time at the reference's own batch sizes (config_linear.yml: batch 1000): at these sizes a step is a
sequence of short launches, so the host-side path (descriptor building, torch glue kernels) is what is measured.
    python tools/small_batch_steps.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmip import distributed as dd, losses as dl
from dmip.linear_problem import LinearForwardProblem
from dmip.models.diffusion import CDE


def run(kind, B, path, steps=100):
    torch.manual_seed(0)
    m = CDE(2, 2, [512, 512, 512])
    opt = torch.optim.Adam(m.sde.a.parameters(), lr=1e-4, fused=True)
    lin = LinearForwardProblem()
    x = torch.randn(B, 2)
    y = lin(x) + 0.3 * torch.randn(B, 2)
    xd, yd = x.cuda(), y.cuda()
    loss_fn = dl.DSMLoss() if kind == "DSM" else dl.PINNLoss(lambda xx, yy: -xx, lam=0.001, lam2=0.1, pde_loss="FPE",
                                                              ic_metric="L2", pde_metric="L1")

    def step_bucket():
        t = m.sample_t(xd)
        return dd.train_step_data_parallel(m, opt, loss_fn, xd, yd, t)[0]

    def step_reference_loop():          # the body of the reference's train_epoch: zero_grad / backward / step
        t = m.sample_t(xd)
        loss, _ = dl.fused_train_step(m, loss_fn, xd, yd, t)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    if path in ("train_epoch", "graph"):   # the drop-in method itself, 50 batches per call; "graph": graph=True
        kw = dict(graph=True) if path == "graph" else {}

        def loader():
            for _ in range(50):
                yield xd, yd
        m.train_epoch(opt, loss_fn, loader, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.train_epoch(opt, loss_fn, loader, **kw)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / 50 * 1e3
    fn = step_bucket if path == "bucket" else step_reference_loop
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


if __name__ == "__main__":
    for kind in ("DSM", "PINN"):
        for B in (1000, 4096, 16384):
            print(f"{kind:4s} batch {B:6d}: zero_grad / backward / step loop {run(kind, B, 'loop'):7.3f} ms/step   "
                  f"fused optimizer step {run(kind, B, 'bucket'):7.3f} ms/step   model.train_epoch "
                  f"{run(kind, B, 'train_epoch'):7.3f} ms/step   train_epoch(graph=True) {run(kind, B, 'graph'):7.3f} ms/step",
                  flush=True)
