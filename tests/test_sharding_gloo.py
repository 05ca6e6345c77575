"""N>1 host logic on CPU: world-size-2 gloo processes shard a batch, transform their slices
(with the oracle standing in for the device kernel -- the sharding code is backend agnostic),
and the assembled result equals the unsharded one; timings reduce to the slowest rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle.scattering1d_oracle import ScatteringOracle
    from tebscat.sharding import max_over_ranks, shard_range, sharded_apply
    orc = ScatteringOracle(4, 256, 2, 16, 2)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(7, 256, generator=g)                        # odd batch: ragged shards
    fn = lambda xs: torch.from_numpy(orc(xs.numpy())).float()
    lo, hi, full = sharded_apply(fn, x, gather=True)
    ref = fn(x)
    slowest = max_over_ranks(10.0 + rank)
    dist.barrier()
    ret[rank] = (lo, hi, bool(torch.equal(full, ref)), slowest, shard_range(7, rank, world))
    dist.destroy_process_group()


def test_shard_range_properties():
    from tebscat.sharding import shard_range
    for n in (0, 1, 7, 16384, 16385):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_two_rank_gloo_sharding():
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0][:2] == (0, 4) and ret[1][:2] == (4, 7)
    assert ret[0][2] and ret[1][2]                               # assembled result == unsharded result
    assert ret[0][3] == ret[1][3] == 11.0                        # slowest rank wins
