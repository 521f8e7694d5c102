"""GPU-side comparison set of SURVEY.md section 8(d): the reference's own GPU route for the three pieces of the
alignment path, timed on the same B200 and the same BASELINE config-2 inputs as this library's kernels.

    python tests/perf_reference_gpu.py [--batch 16] [--refs 5]        (run on a GPU box; not collected by pytest)

"Reference route" = what a user of wdmwhh/MRefSR runs on a GPU today, restated with stock PyTorch / library calls:
  matcher : per (image, reference) pair, patches unfolded and correlated with F.conv2d, max over the reference axis
            (basicsr/archs/ref_map_util.py:26-86, called in a Python loop by corres_generation_arch.py:53-68)
  DynAgg  : chunk / cat / repeat / strided adds / sigmoid / mean-abs host sync in torch
            (basicsr/archs/ref_mrapa_restoration_arch.py:55-73), then the DCN itself through
            (a) the reference's own deform_conv_ext compiled unmodified for sm_100a (oracle/_ref, checker build) and
            (b) torchvision.ops.deform_conv2d (what the mmcv stand-in resolves to)
  fusion  : permute / contiguous / batched matmul / softmax / matmul (ref_mrapa_restoration_arch.py:321-335)
This file lives under tests/ because it loads oracle/_ref (test infrastructure); nothing here is a product path.
One JSON line per measurement; the last line is the summary.
"""
import argparse
import importlib.util
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mrefsr_b200 as M  # noqa: E402
from mrefsr_b200.dcn import dcn_forward_raw, dynagg_dcn_forward  # noqa: E402
from mrefsr_b200.dynagg import DynAggOffsetsFunction  # noqa: E402

DEV = 'cuda:0'
REF_SO = os.path.join(ROOT, 'oracle', '_ref', 'deform_conv_ext_ref.so')


def timeit(fn, iters=3, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(**kw):
    print(json.dumps(kw), flush=True)


# ---------------------------------------------------------------- reference route, stock PyTorch on the GPU
def ref_match_pair(f_in, f_ref):
    """One (image, reference) pair the reference's way: both feature maps per-pixel unit-normalised, reference
    patches L2-normalised and used as conv filters, arg-max over them, input-patch norm applied afterwards."""
    c, h, w = f_in.shape
    a = F.normalize(f_in.reshape(c, -1), dim=0).view(1, c, h, w)
    r = F.normalize(f_ref.reshape(c, -1), dim=0).view(1, c, h, w)
    filt = F.unfold(r, 3)[0].t().reshape(-1, c, 3, 3)                     # [N_ref, C, 3, 3]
    filt = filt / (filt.flatten(1).norm(dim=1).view(-1, 1, 1, 1) + 1e-5)
    sim = F.conv2d(a, filt)[0]                                            # [N_ref, h-2, w-2]
    val, idx = sim.max(0)
    pin = F.unfold(a, 3)[0].norm(dim=0).view(h - 2, w - 2) + 1e-5
    return idx, val / pin


def ref_dynagg_glue(conv_out, pre, dg):
    o1, o2, mask = torch.chunk(conv_out, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1)
    pre = pre.repeat([1, dg, 1, 1, 1])
    reorder = torch.zeros_like(offset)
    reorder[:, 0::2] = pre[..., 1]
    reorder[:, 1::2] = pre[..., 0]
    offset = offset + reorder
    mask = torch.sigmoid(mask)
    if float(torch.mean(torch.abs(offset - reorder))) > 100:   # the reference's host sync
        pass
    return offset, mask


def ref_fusion(emb_t, emb, ass, t):
    """emb_t is conv_emb1's output already multiplied by C^-0.5 (:321), as for the operator under test."""
    n, c, h, w = emb_t.shape
    q = emb_t.permute(0, 2, 3, 1).unsqueeze(3).contiguous().flatten(0, 2)       # (n*h*w, 1, c)
    k = emb.unflatten(0, (n, t)).permute(0, 3, 4, 2, 1).contiguous().flatten(0, 2)            # (n*h*w, c, t)
    v = ass.unflatten(0, (n, t)).permute(0, 3, 4, 1, 2).contiguous().flatten(0, 2)            # (n*h*w, t, 2c)
    p = F.softmax(torch.matmul(q, k), dim=2)
    out = torch.matmul(p, v).squeeze(1).unflatten(0, (n, h, w))
    return out.permute(0, 3, 1, 2).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--refs', type=int, default=5)
    args = ap.parse_args()
    b, r = args.batch, args.refs
    torch.backends.cudnn.allow_tf32 = True          # torch defaults: what a reference user gets
    d = bench.make_inputs(b, r, 1234, DEV)
    emit(what='env', gpu=torch.cuda.get_device_name(0), batch=b, refs=r, torch=torch.__version__,
         cudnn_tf32=torch.backends.cudnn.allow_tf32, matmul_tf32=torch.backends.cuda.matmul.allow_tf32)
    res = {}

    # ---- matcher
    def ref_matcher():
        out = []
        for p in range(b * r):
            out.append(ref_match_pair(d['feat_in'][p // r], d['feat_ref'][p]))
        return out

    def our_matcher():
        return M.feature_match_index_batched(d['feat_in'], d['feat_ref'], is_norm=True, norm_input=True,
                                             normalize_pixels=True, in_div=r)

    res['match_ref_ms'] = timeit(ref_matcher)
    res['match_ours_ms'] = timeit(our_matcher, iters=10, warmup=2)
    idx, val = our_matcher()
    ref_out = ref_matcher()
    agree = sum(int((ref_out[p][0] == idx[p]).sum()) for p in range(b * r)) / float(idx.numel())
    emit(what='matcher', ref_ms=res['match_ref_ms'], ours_ms=res['match_ours_ms'], argmax_agreement_with_tf32_cudnn=agree)

    # ---- DynAgg glue + DCN, three scales
    pre = M.pre_offsets(idx)
    ext = None
    if os.path.exists(REF_SO):
        spec = importlib.util.spec_from_file_location('deform_conv_ext_ref', REF_SO)
        ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ext)
    import torchvision
    tot = dict(glue_ref=0.0, dcn_ext=0.0, dcn_tv=0.0, ours_ops=0.0, ours_fused=0.0)
    for k, (c, hw) in enumerate(bench.SCALES):
        x, conv_out, wgt, bias = d[f'x{c}'], d[f'conv_out{c}'], d[f'w{c}'], d[f'b{c}']
        off, mask = ref_dynagg_glue(conv_out, pre[k], bench.DG)
        t_glue = timeit(lambda: ref_dynagg_glue(conv_out, pre[k], bench.DG))
        t_tv = timeit(lambda: torchvision.ops.deform_conv2d(x, off, wgt, bias, padding=1, mask=mask))
        t_ext = None
        if ext is not None:
            out_ref = x.new_empty(x.shape[0], c, hw, hw)
            empty = x.new_empty(0)
            t_ext = timeit(lambda: ext.modulated_deform_conv_forward(x, wgt, bias, empty, off, mask, out_ref, empty, 3, 3,
                                                                     1, 1, 1, 1, 1, 1, 1, bench.DG, True))

        def ours_ops():
            o, m = DynAggOffsetsFunction.apply(conv_out, pre[k], bench.DG, None)
            return dcn_forward_raw(x, o, m, wgt, bias, (1, 1), (1, 1), (1, 1), 1, bench.DG)

        t_ops = timeit(ours_ops, iters=10, warmup=2)
        t_fused = timeit(lambda: dynagg_dcn_forward(x, conv_out, idx, hw // 40, wgt, bias, bench.DG), iters=10, warmup=2)
        y_tv = torchvision.ops.deform_conv2d(x, off, wgt, bias, padding=1, mask=mask)
        y = dynagg_dcn_forward(x, conv_out, idx, hw // 40, wgt, bias, bench.DG)
        rel = float((y - y_tv).abs().max() / y_tv.abs().max())
        emit(what='dcn', C=c, hw=hw, samples=x.shape[0], glue_ref_ms=t_glue, dcn_ref_ext_ms=t_ext, dcn_torchvision_ms=t_tv,
             ours_operator_boundaries_ms=t_ops, ours_fused_ms=t_fused, rel_diff_vs_torchvision=rel)
        tot['glue_ref'] += t_glue
        tot['dcn_tv'] += t_tv
        tot['dcn_ext'] += t_ext if t_ext is not None else float('nan')
        tot['ours_ops'] += t_ops
        tot['ours_fused'] += t_fused
    res.update({k + '_ms': v for k, v in tot.items()})

    # ---- fusion, three scales
    f_ref = f_ours = 0.0
    for c, hw in bench.SCALES:
        et, em, av = d[f'emb_t{c}'], d[f'emb{c}'], d[f'ass{c}']
        t_ref = timeit(lambda: ref_fusion(et, em, av, r))
        t_ours = timeit(lambda: M.mrapa_attention(et, em, av, r), iters=10, warmup=2)
        rel = float((M.mrapa_attention(et, em, av, r) - ref_fusion(et, em, av, r)).abs().max())
        emit(what='fusion', C=c, hw=hw, ref_ms=t_ref, ours_ms=t_ours, max_abs_diff=rel)
        f_ref += t_ref
        f_ours += t_ours
    res['fusion_ref_ms'], res['fusion_ours_ms'] = f_ref, f_ours

    ref_total = res['match_ref_ms'] + res['glue_ref_ms'] + min(res['dcn_ext_ms'], res['dcn_tv_ms']) \
        if res['dcn_ext_ms'] == res['dcn_ext_ms'] else res['match_ref_ms'] + res['glue_ref_ms'] + res['dcn_tv_ms']
    ref_total += res['fusion_ref_ms']
    ours_total = res['match_ours_ms'] + res['ours_fused_ms'] + res['fusion_ours_ms']
    emit(what='summary', images_per_step=b, **{k: round(v, 4) for k, v in res.items()},
         reference_route_ms=round(ref_total, 3), ours_ms=round(ours_total, 3),
         reference_route_images_per_s=round(b / ref_total * 1e3, 1), ours_images_per_s=round(b / ours_total * 1e3, 1),
         speedup=round(ref_total / ours_total, 2))


if __name__ == '__main__':
    main()
