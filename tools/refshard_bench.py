"""Reference-sharded alignment (BASELINE config 4 shape): H=W=512 HR, R=8 references split across the ranks,
matcher + fused DynAgg/DCN per local reference, one NCCL all-gather of the aligned features per scale, fusion on
every rank.  Launch: python -m torch.distributed.run --nproc-per-node N tools/refshard_bench.py [--check]
Prints one JSON line on rank 0 (time = max over ranks, CUDA events)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrefsr_b200 as M  # noqa: E402
from mrefsr_b200 import parallel as P  # noqa: E402
from mrefsr_b200.dcn import dynagg_dcn_forward, dynagg_dcn_forward_into  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--hw', type=int, default=512)
ap.add_argument('--refs', type=int, default=8)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--fused-gather', action='store_true', help='DCN epilogue stores into every GPU\'s gathered buffer over NVLink (no NCCL all-gather)')
ap.add_argument('--check', action='store_true', help='rank 0 recomputes everything alone and compares')
args = ap.parse_args()

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)

R, n = args.refs, 1
h = args.hw // 4
SCALES = ((256, 1), (128, 2), (64, 4))
g = torch.Generator().manual_seed(7)       # same data on every rank (each uses only its references)
feat_in = torch.randn(n, 256, h, h, generator=g).to(dev)
feat_ref = torch.randn(R, n, 256, h, h, generator=g)
x = {c: torch.randn(R, n, c, h * s, h * s, generator=g) for c, s in SCALES}
conv = {c: torch.randn(R, n, 216, h * s, h * s, generator=g) * 0.5 for c, s in SCALES}
wgt = {c: (torch.randn(c, c, 3, 3, generator=g) * (c * 9) ** -0.5).to(dev) for c, s in SCALES}
bias = {c: torch.zeros(c).to(dev) for c, s in SCALES}
emb_t = {c: (torch.randn(n, c, h * s, h * s, generator=g) * 0.2).to(dev) for c, s in SCALES}


_dev_cache = {}


def on_dev(name, t, r):
    """device copy of reference r's slice of `t`, made once (the timed region holds no host->device traffic)"""
    key = (name, r)
    if key not in _dev_cache:
        _dev_cache[key] = t[r].to(dev)
    return _dev_cache[key]


def run(ref_ids, gather):
    fr = torch.stack([on_dev('fr', feat_ref, r) for r in ref_ids], 0).flatten(0, 1)    # [r_local*n, ...]
    idx, _ = M.feature_match_index_batched(feat_in, fr, is_norm=True, norm_input=True, normalize_pixels=True,
                                           in_div=1)
    outs = []
    for c, s in SCALES:
        xs = torch.stack([on_dev(('x', c), x[c], r) for r in ref_ids], 0).flatten(0, 1)
        cs = torch.stack([on_dev(('conv', c), conv[c], r) for r in ref_ids], 0).flatten(0, 1)
        if gather and args.fused_gather:
            pg = peer_bufs[c]
            pg.begin()
            dynagg_dcn_forward_into(xs, cs, idx, s, wgt[c], bias[c], 8, pg.ptrs, len(ref_ids), R, ref_ids[0])
            full = pg.finish()                                                        # [n, R, C, H, W]
        else:
            y = dynagg_dcn_forward(xs, cs, idx, s, wgt[c], bias[c], 8)                # [r_local*n, C, H, W]
            y = y.view(len(ref_ids), n, c, h * s, h * s).transpose(0, 1).contiguous() # [n, r_local, C, H, W]
            full = P.all_gather_refs(y, R) if gather else y                           # [n, R, C, H, W]
        ass = full.flatten(0, 1).repeat(1, 2, 1, 1)
        emb = full.flatten(0, 1)
        outs.append(M.mrapa_attention(emb_t[c], emb, ass, R))
    return outs


peer_bufs = {}
if world > 1 and args.fused_gather:
    assert n == 1, 'the bench stacks references first; with n == 1 that is the [n, R] slot order'
    peer_bufs = {c: P.PeerGatherBuffer((n, R, c, h * s, h * s), dev) for c, s in SCALES}
lo, hi = P.shard_range(R, rank, world)
mine = list(range(lo, hi))
for _ in range(2):
    outs = run(mine, world > 1)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    outs = run(mine, world > 1)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
err = None
if args.check and rank == 0:
    _dev_cache.clear()
    torch.cuda.empty_cache()
    ref = run(list(range(R)), False)
    err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs, ref))
if rank == 0:
    print(json.dumps({'test': 'reference_sharded', 'hw': args.hw, 'refs': R, 'n_gpus': world, 'fused_gather': bool(args.fused_gather and world > 1), 'ms_per_image': float(ms),
                      'images_per_s': 1e3 / float(ms), 'max_rel_diff_vs_single_gpu': err,
                      'all_gather_bytes_per_image': sum(4 * c * (h * s) ** 2 * R for c, s in SCALES)}), flush=True)
if world > 1:
    dist.destroy_process_group()
