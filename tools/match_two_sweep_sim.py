"""CPU simulation behind DESIGN.md section 6, item 3 (a two-sweep exact arg-max for the matcher).

Question: if the correlation is computed in ONE bf16 pass (a third of today's tensor work), how wrong are the
similarities, and how many reference positions per LR position lie so close to the row maximum that they must be
re-evaluated exactly to keep the arg-max exact?  Emulation: operands rounded to bf16, products accumulated in fp32
(what tcgen05 kind::f16 does), against an fp64 evaluation, on bench.py's synthetic config-2 features (translated,
independent and zero-padded references).

    python tools/match_two_sweep_sim.py [--images 2]

Prints per reference kind: max |error| of the one-pass similarity, and for several safety margins the mean / 99.9th
percentile / max number of candidates per row and whether the exact arg-max is always among them; then the accuracy
of the split-operand schemes (how many tensor passes, bf16 or fp16 halves) that could replace today's three bf16
passes.  Findings of round 1 (2 images x 5 references): one bf16 pass is off by up to 1e-3, so a rigorous margin is
2e-3 and leaves 1.0-1.1 candidates per row (max 4) -- but every row's winner would still need an exact re-evaluation
for max_val (2 GB of patch reads), which eats the gain; no two-pass scheme gets below 8e-5; three passes with FP16
halves are 3.6x more accurate than with bf16 halves (1.7e-6 vs 6.1e-6 max error) at the same cost.
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def patches(f):                       # [C,h,w] -> [N, 9C] rows = 3x3 patches, per-pixel normalised features
    c, h, w = f.shape
    f = F.normalize(f.reshape(c, -1), dim=0).view(1, c, h, w)
    return F.unfold(f, 3)[0].t().contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=2)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    r = 5
    d = bench.make_inputs(args.images, r, 1234, 'cpu')
    kinds = {0: 'translated', 1: 'translated', 2: 'independent', 3: 'independent', 4: 'zero-padded translation'}
    margins = (2e-4, 5e-4, 1e-3, 2e-3)
    stats = {}
    for p in range(args.images * r):
        a = patches(d['feat_in'][p // r])
        b = patches(d['feat_ref'][p])
        b = b / (b.norm(dim=1, keepdim=True) + 1e-5)
        exact = a.double() @ b.double().t()
        one = (a.bfloat16().float() @ b.bfloat16().float().t())           # bf16 operands, fp32 accumulation
        err = float((one.double() - exact).abs().max())
        top2 = exact.topk(2, dim=1).values
        gap = top2[:, 0] - top2[:, 1]
        arg = exact.argmax(dim=1)
        k = kinds[p % r]
        s = stats.setdefault(k, {'err': 0.0, 'rows': 0, 'cand': {m: [] for m in margins}, 'miss': {m: 0 for m in margins},
                                 'flip': 0})
        s['err'] = max(s['err'], err)
        s['rows'] += a.shape[0]
        s['flip'] += int(((one.argmax(dim=1) != arg) & (gap >= 1e-5)).sum())
        mx = one.max(dim=1, keepdim=True).values
        for m in margins:
            cand = one >= mx - m
            s['cand'][m].append(cand.sum(dim=1))
            s['miss'][m] += int((~cand.gather(1, arg[:, None])[:, 0] & (gap >= 1e-5)).sum())
    for k, s in stats.items():
        print('%-24s rows %6d  max |one-pass - exact| = %.2e   one-pass arg-max wrong on %d rows with gap >= 1e-5'
              % (k, s['rows'], s['err'], s['flip']))
        for m in margins:
            c = torch.cat(s['cand'][m]).double()
            print('   margin %.0e: candidates per row mean %.2f  p99.9 %.0f  max %.0f   exact arg-max missed on %d rows'
                  % (m, c.mean(), c.quantile(0.999), c.max(), s['miss'][m]))


def pass_schemes(images=2):
    r = 5
    d = bench.make_inputs(images, r, 1234, 'cpu')
    res = {}
    rows = 0
    for p in range(images * r):
        a = patches(d['feat_in'][p // r])
        b = patches(d['feat_ref'][p])
        b = b / (b.norm(dim=1, keepdim=True) + 1e-5)
        exact = a.double() @ b.double().t()
        top2 = exact.topk(2, dim=1).values
        gap, arg = top2[:, 0] - top2[:, 1], exact.argmax(dim=1)
        rows += a.shape[0]

        def split(x, dt):
            hi = x.to(dt).float()
            return hi, (x - hi).to(dt).float()
        v = {}
        for name, dt in (('bf16', torch.bfloat16), ('fp16', torch.float16)):
            ah, al = split(a, dt)
            bh, bl = split(b, dt)
            v[name + ' x1 (hi.hi)'] = ah @ bh.t()
            v[name + ' x2 (hi.hi + lo.hi)'] = ah @ bh.t() + al @ bh.t()
            v[name + ' x3 (hi.hi + hi.lo + lo.hi)'] = ah @ bh.t() + ah @ bl.t() + al @ bh.t()
        for k, m in v.items():
            e = float((m.double() - exact).abs().max())
            f = int(((m.argmax(dim=1) != arg) & (gap >= 1e-5)).sum())
            x = res.setdefault(k, [0.0, 0])
            x[0], x[1] = max(x[0], e), x[1] + f
    print('split-operand schemes (fp32 accumulation), %d rows:' % rows)
    for k, (e, f) in res.items():
        print('   %-34s max |error| %.2e   arg-max wrong on %d rows with gap >= 1e-5' % (k, e, f))


if __name__ == '__main__':
    main()
    pass_schemes()
