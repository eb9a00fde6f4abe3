"""Multi-GPU plumbing for the hot path: one process per GPU, `torch.distributed` for the rendezvous.

Sampling shards with NO collective on the hot loop: particles of one observation are i.i.d. given y and observations
are independent (the reference loops over them serially, main_diffusion_linear.py:65-74), so every rank integrates a
contiguous range and the Philox counters are keyed by the GLOBAL particle index — the union of the shards is
bit-identical to a single-GPU run of the same seed.  Training is data parallel: every rank evaluates the fused loss on
its slice of the batch with means taken over the GLOBAL batch (`batch_global`), so one SUM all-reduce of the flat
gradient (2-4 MB, NCCL over NVLink) plus the 4 loss scalars reproduces the single-process step.
"""
import gc

import torch
import torch.distributed as dist


def world():
    """(rank, world_size) — (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous, balanced partition of range(n): returns (start, count); counts differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(int(n), world_size)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def sample_sharded(model, y, num_samples=2000, num_steps=200, mean=0, std=1, *, seed, shard='particles',
                   gather=False, group=None, **kw):
    """Posterior sampling split across the ranks of `group`.

    shard='particles': every rank integrates its slice of the `num_samples` particles of each observation
    (BASELINE configs 3 and 5); shard='observations': y is (n_obs, ydim) and every rank takes a slice of the
    observations (config 4: 256 observations x 64K particles over 8 GPUs).  Returns this rank's samples as a CUDA
    tensor, or with gather=True the full result on every rank (one all_gather at the end — never inside the SDE loop).
    """
    rank, ws = world()
    y = torch.as_tensor(y, dtype=torch.float32)
    if shard == 'particles':
        start, count = shard_range(num_samples, rank, ws)
        if y.ndim == 2 and y.shape[0] > 1:
            raise ValueError("shard='particles' takes one observation; use shard='observations' for a batch")
        out = model(y.reshape(-1), num_samples=count, num_steps=num_steps, mean=mean, std=std, seed=seed,
                    gidx_base=start, return_tensor=True, **kw)
        cat_dim = 0
    elif shard == 'observations':
        ys = y.reshape(-1, model.ydim)
        start, count = shard_range(ys.shape[0], rank, ws)
        out = model(ys[start:start + count], num_samples=num_samples, num_steps=num_steps, mean=mean, std=std,
                    seed=seed, gidx_base=start * num_samples, return_tensor=True, **kw)
        cat_dim = 0
    else:
        raise ValueError("shard has to be one of 'particles' or 'observations'")
    if not gather or ws == 1:
        return out
    sizes = [shard_range(num_samples if shard == 'particles' else ys.shape[0], r, ws)[1] for r in range(ws)]
    parts = [torch.empty((s,) + tuple(out.shape[1:]), device=out.device, dtype=out.dtype) for s in sizes]
    dist.all_gather(parts, out.contiguous(), group=group)
    return torch.cat(parts, dim=cat_dim)


def allreduce_gradients(params, extra=None, group=None):
    """SUM all-reduce of the gradients of `params` (and optionally a small tensor `extra`, e.g. the loss scalars) as ONE
    flat bucket: the nets are 2-4 MB, so a single latency-bound NCCL call is the right size on NVSwitch."""
    rank, ws = world()
    params = [p for p in params if p.grad is not None]
    if ws == 1 or not params:
        return extra
    flat = [p.grad.reshape(-1) for p in params]
    if extra is not None:
        flat.append(extra.reshape(-1).to(flat[0].dtype))
    bucket = torch.cat(flat)
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(bucket[off:off + n].view_as(p.grad))
        off += n
    if extra is not None:
        return bucket[off:].view_as(extra).to(extra.dtype)
    return None


class GradBucket:
    """One persistent flat fp32 buffer  [gradients of every parameter, in the kernels' flat order | 8 scalar slots]  whose
    slices ARE the parameters' `.grad`: the fused loss kernels write the gradient straight into it (`grad_out`), one
    `all_reduce` sums gradients and loss scalars of all ranks in place, and `optimizer.step()` reads the same memory —
    the 2-4 MB are touched once per step, with no concatenation, no copy back and no host synchronisation."""

    N_SCALARS = 8

    def __init__(self, nets):
        from ._lib import linear_layers
        self.params = [p for net in nets for lin in linear_layers(net) for p in (lin.weight, lin.bias)]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n + self.N_SCALARS, device=dev, dtype=torch.float32)
        self.n_grad = n
        self.bind()

    def bind(self):
        """(Re-)attach the parameters' .grad to the bucket (e.g. after `optimizer.zero_grad(set_to_none=True)`)."""
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def grads(self):
        return self.flat[:self.n_grad]

    @property
    def scalars(self):
        return self.flat[self.n_grad:]

    def all_reduce(self, group=None):
        rank, ws = world()
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


def _bucket_nets(model):
    a = model.sde.a
    if hasattr(a, 'prior_net'):                       # PosteriorScore: flat order [prior | likelihood] (posterior.py)
        return [a.prior_net, a.likelihood_net]
    return [a]


def fused_optimizer_step(model, optimizer, loss_fn, x, y, t=None, *, data_parallel=False, group=None, batch_global=None,
                         step=True):
    """`optimizer.zero_grad(); loss.backward(); optimizer.step()` of the reference's train loops (models/diffusion.py:
    100-102, :150-152, :223-225) without autograd in between: the fused loss kernels write the gradient into the model's
    GradBucket, whose slices ARE the parameters' `.grad`, and the optimizer steps on that memory.  The gradient is the
    one `loss.backward()` after `zero_grad()` leaves behind; what is saved is host time (at the reference's batch of
    1000 the autograd round trip is a third of the step).  Returns (loss, info) as device tensors.

    data_parallel=True: x, y are THIS rank's rows, the loss means run over the global batch (`batch_global` rows over all
    ranks; default: this rank's rows x world size — what `shard_range` gives when the batch divides evenly; a host
    integer, so no collective and no device->host read is spent on it), and ONE all-reduce sums gradients and loss
    scalars of all ranks in the bucket before the optimizer steps."""
    from .losses import fused_train_step
    ws = world()[1] if data_parallel else 1
    bucket = getattr(model, '_grad_bucket', None)
    if bucket is None:
        bucket = model._grad_bucket = GradBucket(_bucket_nets(model))
    bucket.bind()                                     # cheap: re-points .grad at the bucket if something dropped it
    if t is None:
        t = model.sample_t(x)
    loss_fn.batch_global = int(batch_global) if batch_global else x.shape[0] * ws
    loss_fn.grad_out = bucket.grads
    try:
        with torch.no_grad():
            loss, info = fused_train_step(model, loss_fn, x, y, t)
    finally:                                          # module state must not leak into a later call
        loss_fn.batch_global = 0
        loss_fn.grad_out = None
    if ws == 1:
        if step:                                      # step=False: the caller steps (GraphedTrainStep, eager optimizers)
            optimizer.step()
        return loss, info
    keys = sorted(info)
    sc = bucket.scalars
    sc.zero_()
    sc[0] = loss
    for i, k in enumerate(keys):
        sc[1 + i] = info[k]
    bucket.all_reduce(group)
    optimizer.step()
    out = bucket.scalars.clone()
    return out[0], {k: out[1 + i] for i, k in enumerate(keys)}


def train_step_data_parallel(model, optimizer, loss_fn, x, y, t=None, group=None, batch_global=None):
    """One optimisation step of CDE/CDiffE/PosteriorDiffusionEstimator.train_epoch (models/diffusion.py:80-102) with the
    batch split across ranks (see fused_optimizer_step)."""
    return fused_optimizer_step(model, optimizer, loss_fn, x, y, t, data_parallel=True, group=group,
                                batch_global=batch_global)


class GraphedTrainStep:
    """One training step of `train_epoch` at a fixed batch shape as a CUDA graph: the t draw, the noise draw, the fused loss
    kernels writing into the gradient bucket and — if the optimizer is capturable (`torch.optim.Adam(..., capturable=True)`)
    — the optimizer step are captured once and replayed per batch; an eager optimizer steps after the replay.  At the
    reference's batch of 1000 a step is ~0.15 ms of kernels behind ~0.3 ms of host work (descriptors, ~20 launches): the
    replay is one launch.  The captured region must not synchronise: `loss_fn.initial_condition` has to be plain device
    code, and DMIP_GUARD must be off."""

    def __init__(self, model, optimizer, loss_fn, x, y):
        self.key = self.key_of(optimizer, loss_fn, x, y)
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.xs, self.ys = x.detach().clone(), y.detach().clone()
        self.step_in_graph = all(g.get('capturable', False) for g in optimizer.param_groups)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):                 # lazy initialisations happen outside the capture; no parameter moves
            for _ in range(2):
                self._body(step=False)
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        # No garbage collection inside the capture: collecting an older model's captured step there destroys ITS graph
        # while this stream is capturing, which invalidates the capture (seen with a second model in one process).
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(self.graph):
                self.loss, self.info = self._body(step=self.step_in_graph)
        finally:
            if gc_was_on:
                gc.enable()

    @staticmethod
    def key_of(optimizer, loss_fn, x, y):
        return (tuple(x.shape), tuple(y.shape), x.dtype, x.device, id(optimizer), id(loss_fn))

    def _body(self, step):
        t = self.model.sample_t(self.xs)
        return fused_optimizer_step(self.model, self.optimizer, self.loss_fn, self.xs, self.ys, t, step=step)

    def __call__(self, x, y):
        self.xs.copy_(x)
        self.ys.copy_(y)
        self.graph.replay()
        if not self.step_in_graph:
            self.optimizer.step()
        return self.loss, self.info
