"""Parity of the CUDA multi-reference attention core (through the C ABI) against the oracle and golden vectors."""
import pytest
import torch

import mrefsr_b200 as M
import oracle
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-3


@pytest.mark.parametrize('case', ['t5', 't1', 'pad'])
def test_core_golden(golden, case):
    g = golden('fusion')
    t = int(g(f'{case}.t'))
    out = M.mrapa_attention(g(f'{case}.emb_t').to(DEV), g(f'{case}.emb').to(DEV), g(f'{case}.ass').to(DEV), t)
    assert rel_err(out, g(f'{case}.core_out')) <= 1e-5


@pytest.mark.parametrize('case', ['t5', 't1', 'pad'])
def test_module_golden(golden, case):
    g = golden('fusion')
    t = int(g(f'{case}.t'))
    target, refs = g(f'{case}.target'), g(f'{case}.refs')
    m = M.MRAPAFusion(nf=target.shape[1], ref_nf=refs.shape[2]).to(DEV).eval()
    sd = {k.split('.sd.', 1)[1]: g(k) for k in g.keys() if k.startswith(f'{case}.sd.')}
    m.load_state_dict(sd)          # identical state-dict keys to the reference module
    with torch.no_grad():
        y = m(target.to(DEV), [refs[i].to(DEV) for i in range(t)])
    assert rel_err(y, g(f'{case}.y')) <= TOL


@pytest.mark.parametrize('shape', [(2, 5, 64, 128, 19, 33), (1, 8, 256, 512, 8, 8), (1, 2, 30, 7, 5, 3),
                                   (1, 12, 16, 32, 6, 6)])
def test_forward_backward_vs_oracle(shape):
    n, t, c, cv, h, w = shape
    g = torch.Generator().manual_seed(7)
    q = torch.randn(n, c, h, w, generator=g) * 0.3
    k = torch.randn(n * t, c, h, w, generator=g)
    v = torch.randn(n * t, cv, h, w, generator=g)
    go = torch.randn(n, cv, h, w, generator=g)
    ts = [x.to(DEV).requires_grad_(True) for x in (q, k, v)]
    out = M.mrapa_attention(ts[0], ts[1], ts[2], t)
    out.backward(go.to(DEV))
    rs = [x.double().requires_grad_(True) for x in (q, k, v)]
    ref = oracle.mrapa_attention_oracle(rs[0], rs[1], rs[2], t)
    ref.backward(go.double())
    assert rel_err(out, ref) <= 1e-5
    for a, b, name in zip(ts, rs, ('emb_t', 'emb', 'ass')):
        assert rel_err(a.grad, b.grad) <= 1e-4, name


def test_full_size_convexity():
    """BASELINE config 2 large scale: the output is a convex combination of the t value maps, so it lies
    between their per-pixel min and max; with identical references it equals them."""
    n, t, c, h, w = 2, 5, 64, 160, 160
    g = torch.Generator().manual_seed(1)
    q = torch.randn(n, c, h, w, generator=g).to(DEV)
    k = torch.randn(n * t, c, h, w, generator=g).to(DEV)
    v = torch.randn(n * t, 2 * c, h, w, generator=g).to(DEV)
    out = M.mrapa_attention(q, k, v, t)
    vv = v.view(n, t, 2 * c, h, w)
    assert (out <= vv.max(1).values + 1e-5).all() and (out >= vv.min(1).values - 1e-5).all()
    same = vv[:, :1].expand(-1, t, -1, -1, -1).reshape(n * t, 2 * c, h, w).contiguous()
    out2 = M.mrapa_attention(q, k, same, t)
    assert rel_err(out2, vv[:, 0]) <= 1e-5


def test_too_many_refs_errors():
    q = torch.randn(1, 8, 4, 4, device=DEV)
    with pytest.raises(RuntimeError):
        M.mrapa_attention(q, torch.randn(17, 8, 4, 4, device=DEV), torch.randn(17, 8, 4, 4, device=DEV), 17)


@pytest.mark.parametrize('c,t', [(64, 5), (128, 3), (256, 8), (64, 1)])
def test_nhwc_variant_with_folded_epilogues(c, t):
    """channels-last kernel (bias / PReLU / scale applied inside) against the NCHW core fed the finished tensors."""
    from mrefsr_b200.fusion import mrapa_attention_nhwc
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(c + t)
    n, h, w = 2, 12, 20
    cl = torch.channels_last
    q_raw = torch.randn(n, c, h, w, generator=g).to(DEV)
    k_raw = torch.randn(n * t, c, h, w, generator=g).to(DEV)
    v_raw = torch.randn(n * t, 2 * c, h, w, generator=g).to(DEV)
    bq, bk = (torch.randn(c, generator=g).to(DEV) for _ in range(2))
    bv = torch.randn(2 * c, generator=g).to(DEV)
    sq, sk = torch.rand(1, generator=g).to(DEV), torch.rand(c, generator=g).to(DEV)       # nn.PReLU(1) and nn.PReLU(C)
    scale = c ** -0.5
    q = F.prelu(q_raw + bq.view(1, -1, 1, 1), sq) * scale
    k = F.prelu(k_raw + bk.view(1, -1, 1, 1), sk)
    v = v_raw + bv.view(1, -1, 1, 1)
    want = M.mrapa_attention(q, k, v, t)
    assert rel_err(want, oracle.mrapa_attention_oracle(q.cpu(), k.cpu(), v.cpu(), t)) <= 1e-5
    got = mrapa_attention_nhwc(q_raw.contiguous(memory_format=cl), k_raw.contiguous(memory_format=cl),
                               v_raw.contiguous(memory_format=cl), t, bq, bk, bv, sq, sk, scale)
    assert got.is_contiguous(memory_format=cl)
    assert rel_err(got, want) <= 1e-5
    # no bias / no activation
    want2 = M.mrapa_attention(q_raw, k_raw, v_raw, t)
    got2 = mrapa_attention_nhwc(q_raw.contiguous(memory_format=cl), k_raw.contiguous(memory_format=cl),
                                v_raw.contiguous(memory_format=cl), t)
    assert rel_err(got2, want2) <= 1e-5


@pytest.mark.parametrize('shape', [(2, 5, 64, 128, 40, 40), (1, 8, 128, 256, 16, 24), (1, 3, 256, 512, 8, 12)])
def test_bf16_io(shape):
    """bf16 tensors in / out (the north star's separately stated bf16 tolerance): against the fp64 oracle evaluated on the
    SAME bf16-rounded inputs the result may differ by one bf16 rounding of the output plus fp32 accumulation error:
    4e-3 of the output scale.  Against the fp32-I/O kernel on those inputs the difference is the output rounding alone."""
    n, t, c, cv, h, w = shape
    g = torch.Generator().manual_seed(11)
    q = (torch.randn(n, c, h, w, generator=g) * 0.3).bfloat16()
    k = torch.randn(n * t, c, h, w, generator=g).bfloat16()
    v = torch.randn(n * t, cv, h, w, generator=g).bfloat16()
    out = M.mrapa_attention(q.to(DEV), k.to(DEV), v.to(DEV), t)
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == (n, cv, h, w)
    ref = oracle.mrapa_attention_oracle(q.float(), k.float(), v.float(), t, dtype=torch.float64)
    assert rel_err(out.float(), ref) <= 4e-3
    out32 = M.mrapa_attention(q.float().to(DEV), k.float().to(DEV), v.float().to(DEV), t)
    assert torch.equal(out, out32.bfloat16())


def test_bf16_channels_last_training_path():
    """The autograd Function on bf16 channels-last tensors (the convolutions' outputs under autocast): values and
    gradients equal the fp32 NCHW run on the same bf16-rounded inputs up to the bf16 rounding of outputs / gradients, and
    gradients come back bf16 channels-last (the layout the producing convolutions' backward wants)."""
    n, t, c, cv, h, w = 2, 5, 64, 128, 12, 20
    g = torch.Generator().manual_seed(9)
    cl = torch.channels_last
    q = (torch.randn(n, c, h, w, generator=g) * 0.3).to(DEV).bfloat16()
    k = torch.randn(n * t, c, h, w, generator=g).to(DEV).bfloat16()
    v = torch.randn(n * t, cv, h, w, generator=g).to(DEV).bfloat16()
    go = torch.randn(n, cv, h, w, generator=g).to(DEV).bfloat16()
    a = [x.contiguous(memory_format=cl).requires_grad_(True) for x in (q, k, v)]
    out = M.mrapa_attention(a[0], a[1], a[2], t)
    assert out.dtype == torch.bfloat16 and out.is_contiguous(memory_format=cl)
    out.backward(go.contiguous(memory_format=cl))
    r = [x.float().requires_grad_(True) for x in (q, k, v)]
    ref = M.mrapa_attention(r[0], r[1], r[2], t)
    ref.backward(go.float())
    assert rel_err(out.float(), ref) <= 8e-3
    for x, y, name in zip(a, r, ('emb_t', 'emb', 'ass')):
        assert x.grad.dtype == torch.bfloat16 and x.grad.is_contiguous(memory_format=cl), name
        assert rel_err(x.grad.float(), y.grad) <= 8e-3, name
