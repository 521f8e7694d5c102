"""Parity of the CUDA matcher (through the C ABI) against the oracle and the reference-generated golden vectors."""
import pytest
import torch

import mrefsr_b200 as M
from mrefsr_b200 import matcher as MM
from tests.util import match_parity, unit_features

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
CASES = ['rand', 'shift', 'zeropad', 'raw', 'nonorm', 'strided', 'c256']
TC_MODES = [MM.MATCH_TC_BF16X3, MM.MATCH_TC_BF16X3 | MM.FLAG_NO_BSTRIP, MM.MATCH_TC_BF16X3 | MM.FLAG_NO_DIAG, MM.MATCH_TC_BF16X3 | MM.FLAG_NO_STRIP]


def _run(fi, fr, kw, mode):
    idx, val = M.feature_match_index(fi.to(DEV), fr.to(DEV), mode=mode, **kw)
    torch.cuda.synchronize()
    return idx, val


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('mode', ['auto', 'fp32'])
def test_golden(golden, case, mode):
    g = golden('matcher')
    kw = eval(str(g(f'{case}.kw')))
    fi, fr = g(f'{case}.fi'), g(f'{case}.fr')
    idx, val = _run(fi, fr, kw, mode)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == tuple(g(f'{case}.idx').shape)
    r = match_parity(idx, val, fi, fr, kw)
    assert r['n_bad'] == 0 and r['val_err'] <= 1e-3, r
    # against the reference's own output: exact wherever the gap is resolvable
    ref_idx = g(f'{case}.idx')
    if case != 'zeropad':
        assert (idx.cpu() != ref_idx).sum().item() <= r['n_lowgap']
    else:
        assert torch.equal(idx.cpu()[12:], ref_idx[12:])   # exact-tie plateau: lowest index wins


@pytest.mark.parametrize('case', ['rand', 'shift', 'zeropad', 'c256'])
@pytest.mark.parametrize('mode', TC_MODES)
def test_tensor_core_modes(golden, case, mode):
    g = golden('matcher')
    kw = eval(str(g(f'{case}.kw')))
    fi, fr = g(f'{case}.fi'), g(f'{case}.fr')
    idx, val = _run(fi, fr, kw, mode)
    r = match_parity(idx, val, fi, fr, kw)
    assert r['n_bad'] == 0 and r['val_err'] <= 1e-3, r


def test_bf16_fast_mode_tolerance():
    """Single-pass bf16: stated tolerance 2e-3 absolute on the similarity; indices may differ where the gap is
    below that tolerance."""
    fi, fr = unit_features(1, 256, 40, 40, 1)[0], unit_features(1, 256, 40, 40, 2)[0]
    kw = dict(is_norm=True, norm_input=True)
    idx, val = _run(fi, fr, kw, 'bf16')
    r = match_parity(idx, val, fi, fr, kw, gap_tol=2e-3)
    assert r['n_bad'] == 0, r
    r2 = match_parity(idx, val, fi, fr, kw)
    assert r2['val_err'] < 0.05, r2


@pytest.mark.parametrize('shape', [(256, 40, 40), (256, 75, 75), (64, 3, 3), (64, 3, 40), (128, 17, 23), (64, 5, 130)])
def test_shapes_vs_oracle(shape):
    c, h, w = shape
    fi, fr = unit_features(1, c, h, w, 11)[0], unit_features(1, c, h, w, 12)[0]
    kw = dict(is_norm=True, norm_input=True)
    idx, val = _run(fi, fr, kw, 'auto')
    r = match_parity(idx, val, fi, fr, kw, dtype=torch.float32 if h * w > 3000 else torch.float64)
    assert r['n_bad'] == 0 and r['val_err'] <= 1e-3, r


@pytest.mark.parametrize('w_ref', [3, 8, 33, 40, 75, 96, 97, 104, 130, 259])
def test_diagonal_form_across_reference_widths(w_ref):
    """The diagonal-form kernels take row offsets i * w_ref into their tiles: the strip variant as shared-memory
    descriptor offsets (any w_ref up to 96, odd ones included: swizzle on absolute address bits), the plain variant as
    TMA coordinates (wider grids).  Every variant against the oracle, and the variants against each other, at widths on
    both sides of the strip limit, with an input grid of a different width and tiles that end inside a row."""
    h_ref, h_in, w_in, c = 7, 9, 21, 64
    fi, fr = unit_features(1, c, h_in, w_in, 100 + w_ref)[0], unit_features(1, c, h_ref, w_ref, 200 + w_ref)[0]
    kw = dict(is_norm=True, norm_input=True)
    outs = []
    for mode in TC_MODES:
        idx, val = _run(fi, fr, kw, mode)
        r = match_parity(idx, val, fi, fr, kw)
        assert r['n_bad'] == 0 and r['val_err'] <= 1e-3, (mode, r)
        outs.append((idx.cpu(), val.cpu()))
    for idx, val in outs[1:]:
        same = idx == outs[0][0]
        assert int((~same).sum()) <= r['n_lowgap']
        assert float((val - outs[0][1]).abs().max()) <= 1e-5


def test_different_input_and_ref_sizes():
    fi, fr = unit_features(1, 64, 12, 20, 5)[0], unit_features(1, 64, 30, 17, 6)[0]
    kw = dict(is_norm=True, norm_input=True)
    for mode in ('auto', 'fp32'):
        idx, val = _run(fi, fr, kw, mode)
        r = match_parity(idx, val, fi, fr, kw)
        assert r['n_bad'] == 0 and r['val_err'] <= 1e-3, (mode, r)


@pytest.mark.parametrize('layout', ['BR', 'RB'])
def test_batched_pairs(layout):
    b, r, c, h, w = 2, 3, 256, 20, 24
    fin = torch.randn(b, c, h, w, generator=torch.Generator().manual_seed(3))
    fref = torch.randn(b * r, c, h, w, generator=torch.Generator().manual_seed(4))
    in_div = r if layout == 'BR' else 1
    idx, val = M.feature_match_index_batched(fin.to(DEV), fref.to(DEV), is_norm=True, norm_input=True,
                                             normalize_pixels=True, in_div=in_div)
    import torch.nn.functional as F
    kw = dict(is_norm=True, norm_input=True)
    for p in range(b * r):
        i = (p // in_div) % b
        a = F.normalize(fin[i].reshape(c, -1), dim=0).view(c, h, w)
        rr = F.normalize(fref[p].reshape(c, -1), dim=0).view(c, h, w)
        res = match_parity(idx[p], val[p], a, rr, kw)
        assert res['n_bad'] == 0 and res['val_err'] <= 1e-3, (p, res)


def test_full_size_shift_property():
    """BASELINE config 2 shape (16 images x 5 refs, 256 x 40 x 40): each reference is the input translated by a
    known integer shift, so arg-max and similarity (= 1) are known analytically -- no oracle needed."""
    b, r, c, h, w = 16, 5, 256, 40, 40
    big = unit_features(b, c, h + 8, w + 8, 77)
    shifts = [(0, 0), (1, 3), (4, 2), (6, 6), (8, 1)]
    fin = big[:, :, 4:4 + h, 4:4 + w].contiguous()
    refs = torch.stack([big[:, :, dy:dy + h, dx:dx + w] for dy, dx in shifts], dim=1)   # [B,R,C,h,w]
    idx, val = M.feature_match_index_batched(fin.to(DEV), refs.flatten(0, 1).contiguous().to(DEV), is_norm=True,
                                             norm_input=True)
    idx, val = idx.cpu().view(b, r, h - 2, w - 2), val.cpu().view(b, r, h - 2, w - 2)
    ys, xs = torch.meshgrid(torch.arange(h - 2), torch.arange(w - 2), indexing='ij')
    for k, (dy, dx) in enumerate(shifts):
        # input pixel (y, x) = big (y+4, x+4) = ref_k (y+4-dy, x+4-dx)
        ry, rx = ys + 4 - dy, xs + 4 - dx
        ok = (ry >= 0) & (ry < h - 2) & (rx >= 0) & (rx < w - 2)
        exp = ry * (w - 2) + rx
        assert torch.equal(idx[:, k][:, ok], exp[ok].expand(b, -1)), k
        assert (val[:, k][:, ok] - 1).abs().max() < 1e-4


def test_pre_offsets_golden(golden):
    g = golden('correspondence')
    o1, o2, o4 = M.pre_offsets(g('max_idx').to(DEV))
    assert torch.equal(o1.cpu(), g('relu3_1')) and torch.equal(o2.cpu(), g('relu2_1')) and torch.equal(o4.cpu(), g('relu1_1'))


def _check_flows_against_reference(pre, g):
    """North-star criterion on the module's output (integer flows): a position may differ from the reference's result
    only where the top-2 similarity gap of the fp64 oracle is below 1e-5; every other scale / tap must be the exact
    function of the arg-max map that corres_generation_arch.py:70-105 defines."""
    import torch.nn.functional as F
    import oracle
    f1, f2 = g('f1'), g('f2')
    b, c, h, w = f1.shape
    hp, wp = h - 2, w - 2
    ours = pre['relu3_1'].cpu()[:, 0, :hp, :wp]            # tap (0, 0): the unshifted flow, (x, y) order
    ref = g('relu3_1')[:, 0, :hp, :wp]
    ys, xs = torch.meshgrid(torch.arange(hp), torch.arange(wp), indexing='ij')
    n_diff = 0
    for i in range(b):
        a = F.normalize(f1[i].reshape(c, -1), dim=0).view(c, h, w)
        q = F.normalize(f2[i].reshape(c, -1), dim=0).view(c, h, w)
        _, _, gap = oracle.feature_match_index_oracle(a, q, is_norm=True, norm_input=True, return_gap=True,
                                                      dtype=torch.float64)
        diff = (ours[i] != ref[i]).any(-1)
        assert not (diff & (gap >= 1e-5)).any(), 'arg-max differs from the reference where the gap is resolvable'
        n_diff += int(diff.sum())
        idx = ((ys + ours[i, ..., 1]) * wp + xs + ours[i, ..., 0]).long()
        exp = oracle.pre_offsets_oracle(idx)
        for k in ('relu3_1', 'relu2_1', 'relu1_1'):
            assert torch.equal(pre[k][i].cpu(), exp[k]), k
    return n_diff


def test_correspondence_golden(golden):
    g = golden('correspondence')
    pre = M.correspondence(g('f1').to(DEV), g('f2').to(DEV))
    _check_flows_against_reference(pre, g)


def test_correspondence_arch_module(golden):
    """Module-level drop-in: same forward contract as the reference CorrespondenceGenerationArch."""
    g = golden('correspondence')
    net = M.CorrespondenceGenerationArch(patch_size=3, stride=1, vgg_layer_list=['relu1_1', 'relu2_1', 'relu3_1']).to(DEV)
    f1, f2 = g('f1').to(DEV), g('f2').to(DEV)
    img = torch.rand(f1.shape[0], 3, 4 * f1.shape[2], 4 * f1.shape[3], device=DEV)
    with torch.no_grad():
        pre, feats = net({'dense_features1': f1, 'dense_features2': f2}, img)
    for k in ('relu3_1', 'relu2_1', 'relu1_1'):
        assert pre[k].shape == g(k).shape
    _check_flows_against_reference(pre, g)
    assert feats['relu3_1'].shape == (f1.shape[0], 256, f1.shape[2], f1.shape[3])


@pytest.mark.parametrize('hw', [125, 128])
def test_native_validation_grids(hw):
    """The matcher at the feature grids of the real validation / config-4 shapes: 125 x 125 (CUFED5 val, images padded
    to 500^2: basicsr/data/multi_ref_dataset.py:174-179; N = 15129) and 128 x 128 (512^2 references; N = 15876),
    C = 256, against the fp64 oracle evaluated in chunks.  One pair each: a reference that is partly a translation of
    the input (known answers), partly independent, with a zero-padded border (tie plateau) as the padded CUFED5
    images have."""
    g = torch.Generator().manual_seed(hw)
    big = torch.randn(256, hw + 8, hw + 8, generator=g)
    fi = big[:, 4:4 + hw, 4:4 + hw].contiguous()
    fr = big[:, 2:2 + hw, 7:7 + hw].clone()
    fr[:, :, hw // 2:] = torch.randn(256, hw, hw - hw // 2, generator=g)      # right half: independent content
    fr[:, hw - 9:, :] = 0                                                      # zero padding at the bottom
    import torch.nn.functional as F
    fi_n = F.normalize(fi.reshape(256, -1), dim=0).view(256, hw, hw)
    fr_n = F.normalize(fr.reshape(256, -1), dim=0).view(256, hw, hw)
    kw = dict(is_norm=True, norm_input=True)
    idx, val = _run(fi_n, fr_n, kw, 'auto')
    import oracle
    o_idx, o_val, gap = oracle.feature_match_index_oracle(fi_n, fr_n, return_gap=True, dtype=torch.float64, chunk=2048,
                                                          **kw)
    diff = idx.cpu() != o_idx
    assert not (diff & (gap >= 1e-5)).any(), int((diff & (gap >= 1e-5)).sum())
    assert float(((val.cpu().double() - o_val).abs() / (o_val.abs() + 1e-3)).max()) <= 1e-3
    # known answer where the translated half matches: input (y, x) <-> reference (y + 2, x - 3)
    hp = hw - 2
    ys, xs = torch.meshgrid(torch.arange(hp), torch.arange(hp), indexing='ij')
    known = (xs - 3 >= 0) & (xs - 3 + 2 < hw // 2) & (ys + 2 + 2 < hw - 9)
    assert torch.equal(idx.cpu()[known], ((ys + 2) * hp + xs - 3)[known])


def test_cpu_tensor_raises():
    with pytest.raises(NotImplementedError):
        M.feature_match_index(torch.randn(8, 6, 6), torch.randn(8, 6, 6))
