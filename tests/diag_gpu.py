"""First-contact diagnostics on the GPU box: matcher descriptor variants, parity stats, rough kernel timings.
Usage: python tests/diag_gpu.py  (lives under tests/: it checks against the oracle)  (prints JSON lines; also written to gpurun_out/diag.jsonl)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrefsr_b200 as M  # noqa: E402
from mrefsr_b200 import _lib, matcher as MM, dcn as D  # noqa: E402
from tests.util import match_parity, unit_features, rel_err  # noqa: E402
import oracle  # noqa: E402

DEV = 'cuda:0'
out_path = os.path.join(ROOT, 'gpurun_out', 'diag.jsonl')
os.makedirs(os.path.dirname(out_path), exist_ok=True)
fout = open(out_path, 'a')


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    fout.write(line + '\n')
    fout.flush()


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def matcher_variants():
    kw = dict(is_norm=True, norm_input=True)
    variants = {
        'fp32': MM.MATCH_FP32,
        'x3_nostrip': MM.MATCH_TC_BF16X3 | MM.FLAG_NO_STRIP,
        'x3_strip': MM.MATCH_TC_BF16X3,
        'x3_strip_baseoff': MM.MATCH_TC_BF16X3 | MM.FLAG_BASE_OFFSET,
        'bf16_nostrip': MM.MATCH_TC_BF16 | MM.FLAG_NO_STRIP,
        'bf16_strip': MM.MATCH_TC_BF16,
        'bf16_strip_baseoff': MM.MATCH_TC_BF16 | MM.FLAG_BASE_OFFSET,
    }
    for (c, h, w) in [(64, 16, 18), (256, 40, 40)]:
        fi, fr = unit_features(1, c, h, w, 1)[0], unit_features(1, c, h, w, 2)[0]
        for name, mode in variants.items():
            try:
                idx, val = M.feature_match_index(fi.to(DEV), fr.to(DEV), mode=mode, **kw)
                torch.cuda.synchronize()
                r = match_parity(idx, val, fi, fr, kw)
                emit(test='matcher_variant', shape=[c, h, w], variant=name, **r)
            except Exception as e:  # noqa: BLE001
                emit(test='matcher_variant', shape=[c, h, w], variant=name, error=repr(e)[:300])
                torch.cuda.synchronize()


def matcher_timing():
    b, r, c, h, w = 16, 5, 256, 40, 40
    fin = unit_features(b, c, h, w, 3).to(DEV)
    fref = unit_features(b * r, c, h, w, 4).to(DEV)
    flops = 2.0 * (38 * 38) ** 2 * 2304 * b * r
    for name, mode in [('x3_nostrip', MM.MATCH_TC_BF16X3 | MM.FLAG_NO_STRIP), ('x3_strip', MM.MATCH_TC_BF16X3),
                       ('x3_strip_baseoff', MM.MATCH_TC_BF16X3 | MM.FLAG_BASE_OFFSET),
                       ('bf16_nostrip', MM.MATCH_TC_BF16 | MM.FLAG_NO_STRIP), ('bf16_strip', MM.MATCH_TC_BF16),
                       ('fp32', MM.MATCH_FP32)]:
        try:
            _lib.timing_enable(True)
            _lib.timing_read()
            ms = timeit(lambda: M.feature_match_index_batched(fin, fref, is_norm=True, norm_input=True, mode=mode),
                        iters=3 if name == 'fp32' else 10)
            t = _lib.timing_read()
            _lib.timing_enable(False)
            main = t['match_main'][0] / max(1, t['match_main'][1])
            emit(test='matcher_timing', variant=name, ms_total=ms, ms_main=main, ms_prep=t['match_prep'][0] / max(1, t['match_prep'][1]),
                 algo_tflops_main=flops / main / 1e9)
        except Exception as e:  # noqa: BLE001
            emit(test='matcher_timing', variant=name, error=repr(e)[:300])


def dcn_timing():
    for (c, hw) in [(256, 40), (128, 80), (64, 160)]:
        n = 80
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n, c, hw, hw, generator=g).to(DEV)
        off = (torch.randn(n, 144, hw, hw, generator=g) * 3).to(DEV)
        mask = torch.rand(n, 72, hw, hw, generator=g).to(DEV)
        wgt = (torch.randn(c, c, 3, 3, generator=g) * 0.02).to(DEV)
        bias = torch.zeros(c).to(DEV)
        for mode in ('tf32',):
            try:
                ms = timeit(lambda: D.dcn_forward_raw(x, off, mask, wgt, bias, (1, 1), (1, 1), (1, 1), 1, 8, mode=mode), iters=3)
                flops = 2.0 * c * c * 9 * hw * hw * n
                byt = 4.0 * hw * hw * (c + 216 + c) * n
                emit(test='dcn_timing', C=c, hw=hw, mode=mode, ms=ms, tflops=flops / ms / 1e9, gbs=byt / ms / 1e6)
            except Exception as e:  # noqa: BLE001
                emit(test='dcn_timing', C=c, hw=hw, mode=mode, error=repr(e)[:300])
        try:
            _lib.timing_enable(True); _lib.timing_read()
            conv_out = (torch.randn(n, 216, hw, hw, generator=g) * 0.5).to(DEV)
            mi = torch.randint(0, 38 * 38, (n, 38, 38), generator=g).to(DEV)
            ms = timeit(lambda: D.dynagg_dcn_forward(x, conv_out, mi, hw // 40, wgt, bias, 8), iters=5)
            t = _lib.timing_read(); _lib.timing_enable(False)
            emit(test='dcn_fused_timing', C=c, hw=hw, ms=ms, ms_main=t['dcn_fwd'][0] / max(1, t['dcn_fwd'][1]),
                 ms_aux=t['dcn_aux'][0] / max(1, t['dcn_aux'][1]), gbs=4.0 * hw * hw * (c + 216 + c) * n / ms / 1e6)
        except Exception as e:  # noqa: BLE001
            emit(test='dcn_fused_timing', C=c, hw=hw, error=repr(e)[:300])
        # spot parity on sample 0 against the C oracle
        try:
            out = D.dcn_forward_raw(x[:1], off[:1], mask[:1], wgt, bias, (1, 1), (1, 1), (1, 1), 1, 8, mode='auto')
            from oracle.dcn import modulated_deform_conv_c
            ref = modulated_deform_conv_c(x[:1].cpu(), off[:1].cpu(), mask[:1].cpu(), wgt.cpu(), bias.cpu(), 1, 1, 1, 1, 8)
            emit(test='dcn_parity', C=c, hw=hw, rel_err=rel_err(out, ref))
        except Exception as e:  # noqa: BLE001
            emit(test='dcn_parity', C=c, hw=hw, error=repr(e)[:300])


def fusion_timing():
    for (c, hw) in [(256, 40), (128, 80), (64, 160)]:
        n, t = 16, 5
        q = torch.randn(n, c, hw, hw, device=DEV)
        k = torch.randn(n * t, c, hw, hw, device=DEV)
        v = torch.randn(n * t, 2 * c, hw, hw, device=DEV)
        ms = timeit(lambda: M.mrapa_attention(q, k, v, t), iters=5)
        byt = 4.0 * hw * hw * c * (3 + 3 * t) * n
        emit(test='fusion_timing', C=c, hw=hw, ms=ms, gbs=byt / ms / 1e6)


if __name__ == '__main__':
    emit(test='env', gpu=torch.cuda.get_device_name(0), sms=_lib.lib().mrefsr_sm_count(), torch=torch.__version__)
    which = sys.argv[1:] or ['variants', 'mtime', 'dcn', 'fusion']
    if 'variants' in which:
        matcher_variants()
    if 'mtime' in which:
        matcher_timing()
    if 'dcn' in which:
        dcn_timing()
    if 'fusion' in which:
        fusion_timing()
