"""CPU, world_size 2 over gloo: the ray-sharding + single flat gradient all-reduce logic of
lush_nerf_b200.train (the N>1 path of bench.py).  The sharded step must reproduce the single-process
gradients of the global batch (mean of equal-size shard means == global mean), parameters without a
gradient (NeRF_Noise.alpha_linear in the reference) must stay untouched."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lush_nerf_b200.train import FlatGradAllReduce, lush_loss, shard_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_net():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3), torch.nn.Sigmoid())
    unused = torch.nn.Linear(4, 1)   # never part of the loss: grad stays None
    return net, unused


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net, unused = _make_net()
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(32, 6, generator=g), torch.rand(32, 3, generator=g)
        lo, hi = shard_rows(32, rank, world)
        out = net(x[lo:hi])
        lush_loss(out, out, y[lo:hi]).backward()
        params = list(net.parameters()) + list(unused.parameters())
        FlatGradAllReduce(params)()
        # numpy copies travel by value: a torch tensor in a Queue is a shared-memory handle that dies with this process
        q.put((rank, [p.grad.numpy().copy() if p.grad is not None else None for p in params]))
    finally:
        dist.destroy_process_group()


def test_shard_rows():
    assert [shard_rows(1024, r, 4) for r in range(4)] == [(0, 256), (256, 512), (512, 768), (768, 1024)]
    with pytest.raises(ValueError):
        shard_rows(10, 0, 4)


def test_sharded_step_matches_global_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {r: [None if a is None else torch.from_numpy(a) for a in arrs]
               for r, arrs in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    net, unused = _make_net()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(32, 6, generator=g), torch.rand(32, 3, generator=g)
    out = net(x)
    lush_loss(out, out, y).backward()
    ref = [p.grad for p in net.parameters()] + [None, None]
    for r in range(world):
        for got, want in zip(results[r], ref):
            if want is None:
                assert got is None
            else:
                assert torch.allclose(got, want, atol=1e-6), (got - want).abs().max()
    for a, b in zip(results[0], results[1]):
        assert (a is None and b is None) or torch.equal(a, b)   # replicas stay bit-identical
