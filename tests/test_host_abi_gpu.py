"""The *_host C entry points (host pointers in, host pointers out: H2D, kernels, D2H, stream sync) against the oracle.
These are what a non-PyTorch caller of the C ABI would use."""
import ctypes

import pytest
import torch

import oracle
from mrefsr_b200 import _lib
from tests.util import match_parity, rel_err, unit_features

pytestmark = pytest.mark.gpu


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def test_feature_match_host():
    lib = _lib.lib()
    torch.cuda.set_device(0)
    fi, fr = unit_features(1, 64, 14, 17, 3), unit_features(2, 64, 14, 17, 4)
    idx = torch.empty(2, 12, 15, dtype=torch.int64)
    val = torch.empty(2, 12, 15, dtype=torch.float32)
    rc = lib.mrefsr_feature_match_batched_host(_p(fi), _p(fr), 1, 2, 2, 64, 14, 17, 14, 17, 3, 1, 1, 1, 1, 0, 0, _p(idx),
                                               _p(val), None)
    _lib.check(rc, 'mrefsr_feature_match_batched_host')
    kw = dict(is_norm=True, norm_input=True)
    for p in range(2):
        r = match_parity(idx[p], val[p], fi[0], fr[p], kw)
        assert r['n_bad'] == 0 and r['val_err'] <= 1e-3, r


def test_dcn_forward_host():
    lib = _lib.lib()
    g = torch.Generator().manual_seed(2)
    b, c, h, w, co, dg = 2, 32, 9, 11, 64, 4
    x = torch.randn(b, c, h, w, generator=g)
    off = torch.randn(b, 2 * dg * 9, h, w, generator=g) * 2
    mask = torch.rand(b, dg * 9, h, w, generator=g)
    wgt = torch.randn(co, c, 3, 3, generator=g) * 0.06
    bias = torch.randn(co, generator=g) * 0.1
    out = torch.empty(b, co, h, w)
    rc = lib.mrefsr_modulated_deform_conv_forward_host(_p(x), _p(wgt), _p(bias), _p(off), _p(mask), _p(out), b, c, h, w, co,
                                                       3, 3, 1, 1, 1, 1, 1, 1, 1, dg, 1, 0, None)
    _lib.check(rc, 'mrefsr_modulated_deform_conv_forward_host')
    ref = oracle.modulated_deform_conv_oracle(x, off, mask, wgt, bias, 1, 1, 1, 1, dg, dtype=torch.float64)
    assert rel_err(out, ref) <= 1e-3


def test_attention_forward_host():
    lib = _lib.lib()
    g = torch.Generator().manual_seed(3)
    n, t, c, cv, h, w = 2, 3, 16, 32, 6, 8
    q = torch.randn(n, c, h, w, generator=g) * 0.3
    k = torch.randn(n * t, c, h, w, generator=g)
    v = torch.randn(n * t, cv, h, w, generator=g)
    out = torch.empty(n, cv, h, w)
    rc = lib.mrefsr_mrapa_attention_forward_host(_p(q), _p(k), _p(v), _p(out), n, t, c, cv, h, w, None)
    _lib.check(rc, 'mrefsr_mrapa_attention_forward_host')
    assert rel_err(out, oracle.mrapa_attention_oracle(q, k, v, t, dtype=torch.float64)) <= 1e-5
    lib.mrefsr_arena_release()
