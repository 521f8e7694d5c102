"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: batch partitioning and the reference-sharded
all-gather.  The kernels themselves are single-GPU and covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mrefsr_b200 import parallel as P


def test_shard_range_partitions():
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            spans = [P.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == P.shard_counts(n, world)


def test_shard_batch_structures():
    t = torch.arange(10).view(5, 2)
    assert torch.equal(P.shard_batch(t, 1, 2), t[3:5])
    d = P.shard_batch({'a': t, 'b': t * 2}, 0, 2)
    assert torch.equal(d['a'], t[:3]) and torch.equal(d['b'], t[:3] * 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_refs, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n, c, h, w = 2, 3, 4, 5
        full = torch.arange(n * n_refs * c * h * w, dtype=torch.float32).view(n, n_refs, c, h, w)
        lo, hi = P.shard_range(n_refs, rank, world)
        got = P.all_gather_refs(full[:, lo:hi].contiguous(), n_refs)
        ok1 = torch.equal(got, full)
        got2 = P.align_reference_sharded(lambda r: full[:, r] * 2, list(range(lo, hi)), n_refs)
        ok2 = torch.equal(got2, full * 2)
        # batch sharding needs no collective: the union of the shards is the batch
        b = torch.arange(7)
        mine = P.shard_batch(b, rank, world)
        parts = [None] * world
        dist.all_gather_object(parts, mine.tolist())
        ok3 = sum(parts, []) == b.tolist()
        out[rank] = bool(ok1 and ok2 and ok3)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_refs', [4, 5])
def test_reference_sharded_all_gather_gloo(n_refs):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_refs, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_all_gather_refs_single_process_passthrough():
    x = torch.randn(2, 3, 4, 5, 6)
    assert P.all_gather_refs(x, 3) is x
    with pytest.raises(RuntimeError):
        P.all_gather_refs(x, 4)
