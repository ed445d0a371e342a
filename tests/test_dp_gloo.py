"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding, unique-id exchange plumbing and the
numerical identity the library's all-reduce design relies on (sum of shard gradients with the criterion
divided by the GLOBAL element count == big-batch gradient)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from dcgan_super_resolution_b200 import parallel
    from oracle import ops
    from util import oracle_net, rng, t64
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert parallel.env_rank() == (rank, rank, world)

    # 1. the rendezvous helper broadcasts rank 0's 128-byte id to everyone
    class FakeCtx:
        def comm_unique_id(self):
            return bytes(range(128))

        def comm_init(self, uid):
            self.uid = uid
    fc = FakeCtx()
    uid = parallel.exchange_unique_id(fc, dist)
    assert uid == bytes(range(128)) and fc.uid == uid

    # 2. shard -> local fwd/bwd (global-count criterion) -> all-reduce(sum) == single-rank big batch
    specs = [dict(kind="conv", cin=1, cout=4, k=4, s=2, p=1), dict(kind="lrelu", negval=0.2),
             dict(kind="conv", cin=4, cout=1, k=4, s=1, p=0), dict(kind="sigmoid"), dict(kind="view")]
    D = oracle_net(specs, 3)
    x = t64(rng(6).uniform(0, 1, (8, 1, 8, 8)))
    lo, hi = parallel.shard_bounds(8, world, rank)
    xs = t64(parallel.shard_batch(x.numpy(), world, rank))
    assert xs.shape[0] == hi - lo == 4
    o = D.forward(xs)
    g = -(torch.ones_like(o) - o) / ((1 - o + ops.BCE_EPS) * (o + ops.BCE_EPS)) / 8
    D.zero_grad_parameters()
    D.backward(xs, g)
    grads = D.get_flat_grads().clone()
    dist.all_reduce(grads)
    loss_local = torch.tensor([-(torch.log(o + ops.BCE_EPS)).sum() / 8])
    dist.all_reduce(loss_local)
    out = D.forward(x)
    D.zero_grad_parameters()
    D.backward(x, ops.bce_bwd(out, torch.ones_like(out)))
    ok = torch.allclose(grads, D.get_flat_grads(), atol=1e-12) and abs(float(loss_local) - ops.bce_fwd(out, torch.ones_like(out))) < 1e-12
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_shard_bounds_errors():
    sys.path.insert(0, ROOT)
    from dcgan_super_resolution_b200 import parallel
    import pytest
    assert parallel.shard_bounds(512, 8, 3) == (192, 256)
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 4, 0)
    a = np.arange(24).reshape(6, 4)
    assert np.array_equal(parallel.shard_batch(a, 3, 1), a[2:4])
