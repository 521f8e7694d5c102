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
        # pixel-slab exchange: each rank fuses its rows of every reference; all_gather_slabs reassembles the rows
        fused_full = full.sum(1)                                   # stand-in for the fusion over references
        mine_rows = P.slab_of(full, rank, world, dim=3).sum(1)     # what this rank computes from its slab
        ok4 = torch.equal(P.all_gather_slabs(mine_rows), fused_full)
        out[rank] = bool(ok1 and ok2 and ok3 and ok4)
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


# ---------------------------------------------------------------------------------------------------------
# ragged reference groups (BASELINE config 3: LMR-shaped groups, 2-6 references per image)
def _fake_forward(calls):
    def forward(lq, up, refs):
        assert lq.dim() == 4 and up.dim() == 4 and refs.dim() == 5 and refs.shape[0] == lq.shape[0] == up.shape[0]
        calls.append((int(refs.shape[1]), int(lq.shape[0])))
        return up + refs.mean(dim=1) + lq.mean(dim=(1, 2, 3)).view(-1, 1, 1, 1)
    return forward


def _ragged_samples(counts, seed=0, sizes=None):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i, r in enumerate(counts):
        hw = 8 if sizes is None else sizes[i]
        out.append((torch.rand(3, hw // 4, hw // 4, generator=g), torch.rand(3, hw, hw, generator=g),
                    torch.rand(r, 3, hw, hw, generator=g)))
    return out


def test_shard_ragged_partitions_and_balances():
    counts = [2, 6, 3, 3, 5, 4, 2, 6, 6, 2, 4, 5, 3, 2, 6, 4, 5]
    for world in (1, 2, 3, 8):
        owned = P.shard_ragged(counts, world)
        assert sorted(sum(owned, [])) == list(range(len(counts)))           # every image exactly once
        assert all(o == sorted(o) for o in owned)
        assert [P.shard_ragged(counts, world, r) for r in range(world)] == owned
        loads = [sum(P.image_cost(counts[i]) for i in o) for o in owned]
        assert max(loads) - min(loads) <= max(P.image_cost(c) for c in counts) + 1e-9      # LPT bound
    # size-aware: a 300^2 image costs more than a 160^2 one with the same reference count
    owned = P.shard_ragged([5, 5, 5, 5], 2, pixels=[300 * 300, 160 * 160, 160 * 160, 160 * 160])
    assert owned[0] == [0] and owned[1] == [1, 2, 3]
    with pytest.raises(ValueError):
        P.shard_ragged([1, 2], 2, pixels=[1.0])
    assert P.shard_ragged([], 3) == [[], [], []]


def test_batches_by_shape_preserves_order_and_limits_size():
    keys = ['a', 'b', 'a', 'a', 'b', 'a', 'c', 'a']
    got = P.batches_by_shape(keys, 2)
    assert got == [('a', [0, 2]), ('b', [1, 4]), ('a', [3, 5]), ('c', [6]), ('a', [7])]
    assert P.batches_by_shape([], 4) == []
    with pytest.raises(ValueError):
        P.batches_by_shape(keys, 0)


def test_run_ragged_equals_one_by_one():
    counts = [2, 6, 3, 2, 6, 2, 4, 2, 2]
    samples = _ragged_samples(counts, sizes=[8, 8, 8, 8, 12, 8, 8, 12, 8])
    calls = []
    fwd = _fake_forward(calls)
    got = P.run_ragged(fwd, samples, max_batch=3)
    assert len(got) == len(samples)
    for (lq, up, refs), sr in zip(samples, got):
        want = fwd(lq[None], up[None], refs[None])[0]
        assert torch.equal(sr, want)
    batched = calls[:len(calls) - len(samples)]
    # R=2 at size 8: images 0, 3, 5 (one call of 3) then 8; R=2 at size 12: image 7 alone; R=6: two sizes -> two calls
    assert sorted(batched) == sorted([(2, 3), (2, 1), (2, 1), (6, 1), (6, 1), (3, 1), (4, 1)])
    with pytest.raises(ValueError):
        P.run_ragged(fwd, [(samples[0][0], samples[0][1], samples[4][2])])       # refs of another size


def _ragged_worker(rank, world, port, counts, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        samples = _ragged_samples(counts, seed=3)
        mine = P.shard_ragged(counts, world, rank)
        got = P.run_ragged(_fake_forward([]), [samples[i] for i in mine], max_batch=4)
        parts = [None] * world
        dist.all_gather_object(parts, (mine, [float(t.sum()) for t in got]))        # bookkeeping only, no data path
        merged = {}
        for idxs, sums in parts:
            merged.update(dict(zip(idxs, sums)))
        want = P.run_ragged(_fake_forward([]), samples, max_batch=4)
        out[rank] = sorted(merged) == list(range(len(counts))) and all(
            abs(merged[i] - float(want[i].sum())) < 1e-4 for i in range(len(counts)))
    finally:
        dist.destroy_process_group()


def test_ragged_groups_sharded_gloo():
    """World size 2: each rank derives its share of a ragged batch locally (no collective on the data path), runs it
    bucketed by reference count, and the union of the ranks' results is the whole batch."""
    world, counts = 2, [2, 6, 3, 3, 5, 4, 2, 6, 6, 2, 4]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ragged_worker, args=(world, _free_port(), counts, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
