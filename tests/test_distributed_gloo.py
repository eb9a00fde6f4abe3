"""Host-side multi-GPU logic on CPU: world_size-2 `gloo` processes.  The kernels themselves need a GPU;
what is covered here is the sharding arithmetic, the Philox global-index contract that makes shards independent of the
GPU count, and the single-bucket gradient all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import philox


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port_holder[0], fn, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        return dict(ret)


port_holder = [0]


@pytest.fixture(autouse=True)
def _port():
    port_holder[0] = _free_port()


def test_shard_range_is_a_balanced_partition():
    from dmip.distributed import shard_range
    for n in (0, 1, 7, 8, 1000, 1 << 20, 256):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_philox_streams_do_not_depend_on_the_gpu_count():
    """The sampler keys its noise by (global particle index, step): the union of two ranks' draws with
    gidx_base = shard start equals the single-rank draw."""
    from dmip.distributed import shard_range
    N, seed = 1001, 77
    full = philox.normals(np.arange(N), 5, 0, 3, seed)
    parts = []
    for r in range(2):
        s, c = shard_range(N, r, 2)
        parts.append(philox.normals(s + np.arange(c), 5, 0, 3, seed))
    assert np.array_equal(np.concatenate(parts), full)


def _allreduce_case(rank, world):
    from dmip.distributed import allreduce_gradients, world as w
    assert w() == (rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.Tanh(), torch.nn.Linear(16, 2))
    g = torch.Generator().manual_seed(100 + rank)
    for p in net.parameters():
        p.grad = torch.randn(p.shape, generator=g)
    extra = torch.tensor([1.0 + rank, 10.0 * (rank + 1)])
    out = allreduce_gradients(net.parameters(), extra)
    return [p.grad.clone() for p in net.parameters()], out


def test_single_bucket_gradient_allreduce_gloo():
    res = _run(_allreduce_case)
    want = None
    for r in range(2):
        g = torch.Generator().manual_seed(100 + r)
        gs = [torch.randn(s, generator=g) for s in ((16, 5), (16,), (2, 16), (2,))]
        want = gs if want is None else [a + b for a, b in zip(want, gs)]
    for r in range(2):
        grads, extra = res[r]
        for a, b in zip(grads, want):
            assert torch.allclose(a, b)
        assert torch.allclose(extra, torch.tensor([3.0, 30.0]))


def _dp_loss_case(rank, world):
    """Data-parallel mean over the GLOBAL batch: per-rank partial sums / B_global, summed by all-reduce, equal the
    single-process mean — checked with the CPU oracle's DSM loss standing in for the fused kernel."""
    from dmip.distributed import shard_range
    from oracle import losses as ol
    from oracle.weights import make_params
    B = 37
    g = torch.Generator().manual_seed(3)
    x, y = torch.randn(B, 2, generator=g), torch.randn(B, 2, generator=g)
    t = torch.rand(B, 1, generator=g) * 0.9 + 0.05
    eps = torch.randn(B, 2, generator=g)
    params = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in make_params(1, 5, 2, (32, 32))]
    s, c = shard_range(B, rank, world)
    sl = slice(s, s + c)
    part = ol.dsm_loss(params, "CDE", x[sl], y[sl], t[sl], eps[sl]) * c / B      # local mean re-weighted to the global batch
    part.backward()
    flat = torch.cat([p.grad.reshape(-1) for W, b in params for p in (W, b)] + [part.detach().reshape(1)])
    dist.all_reduce(flat)
    full_params = [(W.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)) for W, b in params]
    full = ol.dsm_loss(full_params, "CDE", x, y, t, eps)
    full.backward()
    ref = torch.cat([p.grad.reshape(-1) for W, b in full_params for p in (W, b)] + [full.detach().reshape(1)])
    return float((flat - ref).abs().max() / ref.abs().max())


def test_data_parallel_loss_equals_single_process():
    res = _run(_dp_loss_case)
    assert max(res.values()) < 1e-5


def _bucket_case(rank, world):
    """GradBucket: the parameters' .grad alias one flat buffer; one all_reduce sums gradients and scalars in place."""
    from dmip.distributed import GradBucket
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.Tanh(), torch.nn.Linear(16, 2))
    b = GradBucket([net])
    assert b.flat.numel() == 5 * 16 + 16 + 16 * 2 + 2 + GradBucket.N_SCALARS
    g = torch.Generator().manual_seed(100 + rank)
    b.grads.copy_(torch.randn(b.n_grad, generator=g))          # what the fused kernel does through grad_out
    b.scalars[0] = 1.0 + rank
    b.all_reduce()
    lin0 = net[0]
    aliased = lin0.weight.grad.data_ptr() == b.flat.data_ptr()
    for p in net.parameters():
        p.grad = None                                           # optimizer.zero_grad(set_to_none=True)
    b.bind()
    aliased = aliased and lin0.weight.grad.data_ptr() == b.flat.data_ptr()
    return b.flat.clone(), aliased


def test_grad_bucket_allreduce_gloo():
    res = _run(_bucket_case)
    n = 5 * 16 + 16 + 16 * 2 + 2
    want = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(2))
    for r in range(2):
        flat, aliased = res[r]
        assert aliased
        assert torch.allclose(flat[:n], want)
        assert float(flat[n]) == 3.0
