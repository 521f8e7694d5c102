"""Extra measurement legs of bench.py (the driver-visible record of BASELINE configs 3, 4, 5 and of a sustained run).

Each leg is a function bench.py calls inside its own process group (rank / world / dist are bench.py's) and returns a
JSON-able dict for an extra key of the bench line; none of them touches the contract keys.  All timing is on the
device (CUDA events on the launching stream), max over ranks.

    ragged_leg      config 3: LMR-shaped groups (300x300 HR, 2..6 references per image), batch-sharded by cost
                    (parallel.shard_ragged), whole network through MRefSRPipeline.forward_ragged
    refshard_leg    config 4: 512x512 HR, 8 references split across the ranks, per-scale exchange of the aligned
                    features (NCCL all-gather, and the DCN epilogue's NVLink peer stores), fusion on every rank
    train_step_leg  config 5: stage-3 restoration training step (L1 loss, Adam, DDP), 5 references @160^2
    sustained_leg   the hot path of the headline metric repeated for >= 3 s with the clocks sampled
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _max_over_ranks(x, dev, dist):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(dist):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


# ----------------------------------------------------------------------------------------------------------
def ragged_leg(dev, rank, world, dist, images_per_gpu=6, hr=300, steps=2):
    """BASELINE config 3 (options/train/stage3_5ref_restoration_mse_lp.yml:17,23: gt_size 300): a ragged global batch of
    images_per_gpu * world images with 2..6 references each, assigned to ranks by parallel.shard_ragged (identical on
    every rank, no communication), each rank bucketing its images by reference count into batched forwards."""
    from mrefsr_b200 import parallel as P
    from mrefsr_b200.models import MRefSRPipeline
    torch.manual_seed(10)
    net = MRefSRPipeline().eval().to(dev).channels_last_()
    n = images_per_gpu * world
    g = torch.Generator().manual_seed(77)
    counts = [int(torch.randint(2, 7, (1,), generator=g)) for _ in range(n)]
    mine = P.shard_ragged(counts, world, rank)
    gl = torch.Generator().manual_seed(1000 + rank)
    samples = []
    for i in mine:
        lq = torch.rand(3, hr // 4, hr // 4, generator=gl)
        up = torch.nn.functional.interpolate(lq[None], scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)[0]
        refs = torch.rand(counts[i], 3, hr, hr, generator=gl)
        samples.append((lq.to(dev), up.to(dev), refs.to(dev)))
    res = {}
    with torch.no_grad():
        for mode in ('eager', 'graphs'):
            kw = {'graphs': mode == 'graphs'}
            out = net.forward_ragged(samples, **kw)          # warm-up (cuDNN autotune, workspaces, graph capture)
            _barrier(dist)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = net.forward_ragged(samples, **kw)
            e1.record()
            _barrier(dist)
            res[mode] = (_max_over_ranks(e0.elapsed_time(e1) / steps, dev, dist), [o.clone() for o in out])
    ms = res['graphs'][0]
    out = res['graphs'][1]
    same = max(float((a - b).abs().max()) for a, b in zip(res['eager'][1], res['graphs'][1])) if out else 0.0
    ok = all(o.shape == (3, hr, hr) and bool(torch.isfinite(o).all()) for o in out)
    eager_ms = res['eager'][0]
    net.clear_graphs()
    del net, samples, out, res
    torch.cuda.empty_cache()
    return {'config': 'BASELINE config 3: LMR-shaped groups, %dx%d HR, 2-6 references per image, %d images per step '
                      'batch-sharded over %d GPU(s) by reference-count cost, no collective' % (hr, hr, n, world),
            'images_per_s': n / (ms / 1e3), 'ms_per_step': ms, 'images_per_step': n, 'ref_counts': counts,
            'images_on_rank0': len(mine), 'outputs_finite': ok,
            'eager_ms_per_step': eager_ms, 'eager_images_per_s': n / (eager_ms / 1e3),
            'max_abs_diff_graphs_vs_eager': same,
            'note': 'whole network (cuDNN convolutions + this library), forward_ragged(graphs=True): one CUDA graph per '
                    'shape class (batch, R, H, W), captured before the timed region; eager_* = the same call launched '
                    'eagerly (host-bound); inputs resident'}


# ----------------------------------------------------------------------------------------------------------
def refshard_leg(dev, rank, world, dist, hw=512, n_refs=8, steps=2):
    """BASELINE config 4: the alignment path on one 512x512 image with 8 references split across the ranks.  Two
    exchange mechanisms are timed back to back on the same inputs: NCCL all-gather of the aligned features per scale
    (parallel.all_gather_refs) and the DCN epilogue storing its tiles straight into every GPU's gathered tensor over
    NVLink (mrefsr_dynagg_dcn_forward_multi + parallel.PeerGatherBuffer), and the pixel-slab exchange (each tile stored
    only to the GPU that owns its rows, fusion of 1 / world of the pixels per GPU, all-gather of the fused result:
    mrefsr_dynagg_dcn_forward_slabs + parallel.PeerSlabBuffer); outputs are compared bit for bit."""
    import mrefsr_b200 as M
    from mrefsr_b200 import parallel as P
    from mrefsr_b200.dcn import dynagg_dcn_forward, dynagg_dcn_forward_into
    R, n, h = n_refs, 1, hw // 4
    scales = ((256, 1), (128, 2), (64, 4))
    g = torch.Generator().manual_seed(7)           # same data on every rank (each uses only its references)
    feat_in = torch.randn(n, 256, h, h, generator=g).to(dev)
    feat_ref = torch.randn(R, n, 256, h, h, generator=g)
    x = {c: torch.randn(R, n, c, h * s, h * s, generator=g) for c, s in scales}
    conv = {c: torch.randn(R, n, 216, h * s, h * s, generator=g) * 0.5 for c, s in scales}
    wgt = {c: (torch.randn(c, c, 3, 3, generator=g) * (c * 9) ** -0.5).to(dev) for c, s in scales}
    bias = {c: torch.zeros(c).to(dev) for c, s in scales}
    emb_t = {c: (torch.randn(n, c, h * s, h * s, generator=g) * 0.2).to(dev) for c, s in scales}
    lo, hi = P.shard_range(R, rank, world)
    mine = list(range(lo, hi))
    fr = torch.stack([feat_ref[r] for r in mine], 0).flatten(0, 1).to(dev)
    xs = {c: torch.stack([x[c][r] for r in mine], 0).flatten(0, 1).to(dev) for c, s in scales}
    cs = {c: torch.stack([conv[c][r] for r in mine], 0).flatten(0, 1).to(dev) for c, s in scales}
    del feat_ref, x, conv
    peer = {c: P.PeerGatherBuffer((n, R, c, h * s, h * s), dev) for c, s in scales} if world > 1 else {}
    slabs = {c: P.PeerSlabBuffer(n, R, c, h * s, h * s, dev) for c, s in scales} if world > 1 else {}
    emb_t_slab = {c: P.slab_of(emb_t[c], rank, world) for c, s in scales} if world > 1 else {}

    def run_slabs():
        """pixel-slab exchange: every rank receives all references for ITS rows (from the DCN epilogues), fuses 1/world
        of the pixels and all-gathers the fused result"""
        idx, _ = M.feature_match_index_batched(feat_in, fr, is_norm=True, norm_input=True, normalize_pixels=True, in_div=1)
        outs = []
        for c, s in scales:
            pg = slabs[c]
            pg.begin()
            dynagg_dcn_forward_into(xs[c], cs[c], idx, s, wgt[c], bias[c], 8, pg.ptrs_by_rank, len(mine), R, mine[0],
                                    slab_rows=pg.slab_rows)
            emb = pg.finish().flatten(0, 1)                                            # [n*R, C, H/world, W]
            outs.append(P.all_gather_slabs(M.mrapa_attention(emb_t_slab[c], emb, emb.repeat(1, 2, 1, 1), R)))
        return outs

    def run(fused_gather):
        idx, _ = M.feature_match_index_batched(feat_in, fr, is_norm=True, norm_input=True, normalize_pixels=True, in_div=1)
        outs = []
        for c, s in scales:
            if world > 1 and fused_gather:
                pg = peer[c]
                pg.begin()
                dynagg_dcn_forward_into(xs[c], cs[c], idx, s, wgt[c], bias[c], 8, pg.ptrs, len(mine), R, mine[0])
                full = pg.finish()                                                     # [n, R, C, H, W]
            else:
                y = dynagg_dcn_forward(xs[c], cs[c], idx, s, wgt[c], bias[c], 8)
                y = y.view(len(mine), n, c, h * s, h * s).transpose(0, 1).contiguous()
                full = P.all_gather_refs(y, R) if world > 1 else y
            emb = full.flatten(0, 1)
            outs.append(M.mrapa_attention(emb_t[c], emb, emb.repeat(1, 2, 1, 1), R))
        return outs

    res = {}
    outs = {}
    for name, fg in (('nccl_all_gather', False), ('peer_store_epilogue', True), ('peer_store_pixel_slabs', None)):
        if fg is not False and world == 1:
            continue
        fn = run_slabs if fg is None else (lambda fg=fg: run(fg))
        for _ in range(2):
            outs[name] = fn()
        _barrier(dist)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            outs[name] = fn()
        e1.record()
        _barrier(dist)
        res[name] = _max_over_ranks(e0.elapsed_time(e1) / steps, dev, dist)
    same = None
    if len(outs) > 1:
        same = all(torch.equal(a, b) for name in outs if name != 'nccl_all_gather'
                   for a, b in zip(outs['nccl_all_gather'], outs[name]))
    checksum = float(sum(o.double().sum() for o in outs['nccl_all_gather']))
    ex_bytes = sum(4 * c * (h * s) ** 2 * n * (R - len(mine)) for c, s in scales)       # received per GPU per image
    # pixel slabs: (R - r_local) references x 1/world of the rows in, then (world - 1)/world of the fused 2C planes
    ex_slabs = sum(4 * c * (h * s) ** 2 * n * ((R - len(mine)) + 2 * (world - 1)) // world for c, s in scales)
    best = min(res.values())
    out = {'config': 'BASELINE config 4: %dx%d HR, %d references sharded over %d GPU(s), aligned features exchanged per '
                     'scale, fusion on every rank; alignment path only, inputs resident' % (hw, hw, R, world),
           'ms_per_image': res, 'images_per_s': 1e3 / best,
           'exchange_bytes_received_per_gpu': {'all_gather_of_aligned_features': ex_bytes, 'pixel_slabs': ex_slabs},
           'mechanisms_bit_identical': same, 'checksum': checksum}
    del peer, slabs, outs
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------
def train_step_leg(dev, rank, world, dist, batch=12, refs=5, hr=160, steps=3, bf16=True):
    """BASELINE config 5: stage-3 restoration training step (options/train/stage3_5ref_restoration_mse.yml:23,77: batch
    12 per GPU, L1 pixel loss), Adam on net_g, DDP gradient all-reduce when world > 1, bf16 autocast for the plain
    convolutions; DynAgg / DCNv2 / MRAPAFusion forward and backward through this library's autograd Functions."""
    from mrefsr_b200.models import MRefSRPipeline
    torch.manual_seed(10)
    pipe = MRefSRPipeline().to(dev)
    pipe.net_extractor.eval()
    pipe.net_map.eval()
    for p in list(pipe.net_extractor.parameters()) + list(pipe.net_map.parameters()):
        p.requires_grad_(False)
    pipe.channels_last_()                    # cuDNN's native layout for the plain convolutions, frozen nets included
    net_g = pipe.net_g.train()
    for name in ('small', 'medium', 'large'):      # zero-init in the reference: give every backward path a signal
        getattr(net_g.dyn_agg_restore, f'{name}_dyn_agg').conv_offset_mask.weight.data.normal_(0, 1e-3)
    model = torch.nn.parallel.DistributedDataParallel(net_g, device_ids=[dev.index]) if world > 1 else net_g
    opt = torch.optim.Adam(net_g.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1234 + rank)
    gt = torch.rand(batch, 3, hr, hr, generator=g).to(dev)
    lq = torch.nn.functional.interpolate(gt, scale_factor=0.25, mode='bicubic', align_corners=False).clamp(0, 1)
    up = torch.nn.functional.interpolate(lq, scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)
    ref_list = [torch.rand(batch, 3, hr, hr, generator=g).to(dev) for _ in range(refs)]

    refs_stacked = torch.stack(ref_list, 1)                      # [B, R, 3, H, W]

    def step():
        # frozen half (extractor, matcher, VGG19 of the references) batched over the references, then net_g on the stacked
        # tensors: the per-reference loops of the reference's training step, same arithmetic per (image, reference)
        pres, rfs, n_refs = pipe.correspondences(up, refs_stacked)
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bf16):
            out = model(lq.contiguous(memory_format=torch.channels_last), pres, rfs, n_refs)
        loss = torch.nn.functional.l1_loss(out.float(), gt)
        loss.backward()
        opt.step()
        return loss

    losses = [float(step()) for _ in range(2)]
    _barrier(dist)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses.append(float(step()))
    e1.record()
    _barrier(dist)
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, dev, dist)
    finite = all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in net_g.parameters())
    del pipe, model, opt
    torch.cuda.empty_cache()
    return {'config': 'BASELINE config 5: stage-3 restoration training step, batch %d per GPU, %d refs @%d^2, L1 loss, '
                      'Adam, %s, channels_last net_g, %d GPU(s)%s' % (batch, refs, hr, 'bf16 autocast convolutions' if bf16 else 'fp32',
                                                 world, ' (DDP)' if world > 1 else ''),
            'ms_per_step': ms, 'images_per_s': batch * world / (ms / 1e3), 'losses': [round(v, 5) for v in losses],
            'all_grads_finite': finite,
            'note': 'whole step inside the timed region: frozen extractor / matcher / VGG19 of the references (evaluated once '
                    'for all references: MRefSRPipeline.correspondences), net_g forward + backward (cuDNN convolutions; '
                    'DynAgg / DCNv2 / fusion / bias-activation epilogues = this library), Adam'}


# ----------------------------------------------------------------------------------------------------------
def sustained_leg(step_fn, images_per_step, sampler_cls, gpu_index, dev, dist, seconds=3.0):
    """The headline hot path repeated back to back for >= `seconds` (the timed region of the headline number is only a
    fraction of a second): images/s and the clocks nvidia-smi reports during the run."""
    _barrier(dist)
    sampler = sampler_cls(gpu_index) if sampler_cls is not None else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    n = 0
    while True:
        for _ in range(20):
            step_fn()
        n += 20
        torch.cuda.synchronize()
        if time.perf_counter() - t0 >= seconds:
            break
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None
    ms = e0.elapsed_time(e1)
    ms_max = _max_over_ranks(ms, dev, dist)
    return {'seconds': ms_max / 1e3, 'steps': n, 'ms_per_step': ms_max / n, 'images_per_s_per_gpu': images_per_step * n / (ms_max / 1e3),
            'clocks': clocks}
