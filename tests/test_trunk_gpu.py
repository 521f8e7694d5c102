"""Trunk glue kernels (csrc/trunk.cu) against the plain torch ops they replace, and the fused
inference path of the trunk modules against their ordinary nn.Module path."""
import pytest
import torch
import torch.nn.functional as F

import mrefsr_b200 as M
from mrefsr_b200 import trunk as T
from mrefsr_b200.models import ResidualBlockNoBN, ContrasExtractorLayer, ContentExtractor
from mrefsr_b200.archs import VGGFeatureExtractor
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _ref(x, bias, act, slope, residual, scale):
    y = x + (bias.view(1, -1, 1, 1) if bias is not None else 0)
    if act == T.ACT_LEAKY:
        s = slope.view(1, -1, 1, 1) if torch.is_tensor(slope) else slope
        y = torch.where(y > 0, y, y * s)
    elif act == T.ACT_SIGMOID:
        y = torch.sigmoid(y)
    y = y * scale
    return y + residual if residual is not None else y


@pytest.mark.parametrize('shape', [(3, 8, 12, 16), (2, 5, 7, 9), (1, 64, 40, 40)])      # vectorised and scalar paths
@pytest.mark.parametrize('act', [T.ACT_NONE, T.ACT_LEAKY, T.ACT_SIGMOID])
@pytest.mark.parametrize('with_bias,with_res', [(True, False), (True, True), (False, True), (False, False)])
def test_bias_act(shape, act, with_bias, with_res):
    g = torch.Generator().manual_seed(sum(shape) + act)
    x = torch.randn(*shape, generator=g).to(DEV)
    bias = torch.randn(shape[1], generator=g).to(DEV) if with_bias else None
    res = torch.randn(*shape, generator=g).to(DEV) if with_res else None
    want = _ref(x, bias, act, 0.1, res, 0.75)
    got = T.bias_act_(x.clone(), bias, act, 0.1, res, None, 0.75)
    assert rel_err(got, want) <= 2e-6


@pytest.mark.parametrize('n_slope', [1, 6])
def test_bias_act_prelu(n_slope):
    g = torch.Generator().manual_seed(n_slope)
    x = torch.randn(2, 6, 8, 8, generator=g).to(DEV)
    bias = torch.randn(6, generator=g).to(DEV)
    w = torch.rand(n_slope, generator=g).to(DEV)
    want = F.prelu(x + bias.view(1, -1, 1, 1), w) * 0.125
    got = T.bias_act_(x.clone(), bias, T.ACT_LEAKY, 0.0, None, w, 0.125)
    assert rel_err(got, want) <= 2e-6


def test_bias_act_errors():
    x = torch.zeros(1, 4, 4, 4, device=DEV)
    with pytest.raises(RuntimeError):
        T.bias_act_(x, None, 7)                                           # unknown activation
    with pytest.raises(RuntimeError):
        T.bias_act_(x, None, T.ACT_LEAKY, 0.0, None, torch.zeros(3, device=DEV))   # slope tensor of the wrong size
    with pytest.raises(ValueError):
        T.bias_act_(x, None, T.ACT_NONE, 0.0, torch.zeros(1, 4, 4, 5, device=DEV))


@pytest.mark.parametrize('shape', [(2, 16, 12, 12), (1, 3, 5, 7)])
def test_attn_modulate(shape):
    g = torch.Generator().manual_seed(shape[1])
    refs, mul, add = (torch.randn(*shape, generator=g).to(DEV) for _ in range(3))
    bm, ba = (torch.randn(shape[1], generator=g).to(DEV) for _ in range(2))
    want = refs * torch.sigmoid(mul + bm.view(1, -1, 1, 1)) * 2 + (add + ba.view(1, -1, 1, 1))
    got = T.attn_modulate_(refs.clone(), mul, add, bm, ba)
    assert rel_err(got, want) <= 2e-6


def _module_vs_fast(m, *xs):
    m = m.to(DEV).eval()
    for p in m.parameters():
        p.data.normal_(0, 0.05)
    with torch.enable_grad():          # the ordinary nn.Module path
        want = m(*xs)
    with torch.no_grad():              # the fused inference path
        got = m(*xs)
    return got, want


def test_resblock_fast_path():
    x = torch.randn(2, 64, 24, 24, device=DEV)
    got, want = _module_vs_fast(ResidualBlockNoBN(64, res_scale=0.5), x)
    assert rel_err(got, want) <= 1e-3      # the two convolutions may pick different cuDNN (TF32) algorithms


def test_content_extractor_fast_path():
    x = torch.rand(2, 3, 20, 20, device=DEV)
    got, want = _module_vs_fast(ContentExtractor(n_blocks=3), x)
    assert rel_err(got, want) <= 1e-3


def test_extractor_and_vgg_fast_path():
    x = torch.rand(2, 3, 32, 32, device=DEV)
    got, want = _module_vs_fast(ContrasExtractorLayer(), x)
    assert rel_err(got, want) <= 1e-3
    vgg = VGGFeatureExtractor(['relu1_1', 'relu2_1', 'relu3_1'])
    got, want = _module_vs_fast(vgg, x)
    assert sorted(got) == sorted(want) == ['relu1_1', 'relu2_1', 'relu3_1']
    for k in want:
        assert rel_err(got[k], want[k]) <= 1e-3, k


def test_mrapa_fusion_fast_path():
    m = M.MRAPAFusion(nf=16, ref_nf=32)
    target = torch.randn(2, 16, 10, 12, device=DEV)          # 10 -> reflect-padded to 12
    refs = [torch.randn(2, 32, 10, 12, device=DEV) for _ in range(3)]
    got, want = _module_vs_fast(m, target, refs)
    assert got.shape == want.shape == (2, 16, 10, 12)
    assert rel_err(got, want) <= 1e-3


@pytest.mark.parametrize('shape', [(2, 8, 6, 10), (2, 6, 5, 7)])           # C % 4 == 0 (float4) and not
def test_glue_kernels_channels_last(shape):
    g = torch.Generator().manual_seed(shape[1])
    cl = torch.channels_last
    x = torch.randn(*shape, generator=g).to(DEV).contiguous(memory_format=cl)
    res = torch.randn(*shape, generator=g).to(DEV)                          # NCHW residual: converted by the wrapper
    bias, w = torch.randn(shape[1], generator=g).to(DEV), torch.rand(shape[1], generator=g).to(DEV)
    want = F.prelu(x + bias.view(1, -1, 1, 1), w) * 0.5 + res
    got = T.bias_act_(x.clone(memory_format=cl), bias, T.ACT_LEAKY, 0.0, res, w, 0.5)
    assert got.is_contiguous(memory_format=cl) and rel_err(got, want) <= 2e-6
    mul, add = (torch.randn(*shape, generator=g).to(DEV).contiguous(memory_format=cl) for _ in range(2))
    want = x * torch.sigmoid(mul + bias.view(1, -1, 1, 1)) * 2 + (add + w.view(1, -1, 1, 1))
    got = T.attn_modulate_(x.clone(memory_format=cl), mul, add, bias, w)
    assert rel_err(got, want) <= 2e-6


@pytest.mark.parametrize('in_cl,out_cl', [(True, True), (True, False), (False, True)])
def test_dynagg_dcn_layout_flags(in_cl, out_cl):
    """NHWC input / NHWC output / folded leaky-ReLU of the fused DynAgg DCN against the plain NCHW call."""
    from mrefsr_b200.dcn import dynagg_dcn_forward
    g = torch.Generator().manual_seed(3)
    b, c, hw, dg, s = 3, 64, 24, 8, 2
    x = torch.randn(b, c, hw, hw, generator=g).to(DEV)
    conv_out = torch.randn(b, 3 * dg * 9, hw, hw, generator=g).to(DEV)
    hp = hw // s - 2
    max_idx = torch.randint(0, hp * hp, (b, hp, hp), generator=g).to(DEV)
    wgt = (torch.randn(c, c, 3, 3, generator=g) * 0.05).to(DEV)
    bias = torch.randn(c, generator=g).to(DEV)
    want = F.leaky_relu(dynagg_dcn_forward(x, conv_out, max_idx, s, wgt, bias, dg), 0.1)
    xin = x.contiguous(memory_format=torch.channels_last) if in_cl else x
    got = dynagg_dcn_forward(xin, conv_out, max_idx, s, wgt, bias, dg, out_slope=0.1, out_channels_last=out_cl)
    assert got.is_contiguous(memory_format=torch.channels_last) == out_cl
    assert rel_err(got, want) <= 1e-6          # same kernel, same arithmetic: only addressing differs


@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('pre', [False, True])
def test_bias_act_shared_residual(cl, pre):
    """residual of sample b taken from sample b // res_div, before or after the activation."""
    g = torch.Generator().manual_seed(11)
    b, r, c, h, w = 2, 3, 8, 6, 10
    x = torch.randn(b * r, c, h, w, generator=g).to(DEV)
    res = torch.randn(b, c, h, w, generator=g).to(DEV)
    bias = torch.randn(c, generator=g).to(DEV)
    rr = res.repeat_interleave(r, dim=0)
    y = x + bias.view(1, -1, 1, 1)
    want = F.leaky_relu(y + rr, 0.1) * 0.5 if pre else F.leaky_relu(y, 0.1) * 0.5 + rr
    xin = x.contiguous(memory_format=torch.channels_last) if cl else x.clone()
    got = T.bias_act_(xin, bias, T.ACT_LEAKY, 0.1, res, None, 0.5, res_div=r, res_pre=pre)
    assert rel_err(got, want) <= 2e-6
    with pytest.raises(ValueError):
        T.bias_act_(x.clone(), bias, T.ACT_LEAKY, 0.1, res, None, 0.5, res_div=2)


@pytest.mark.parametrize('shape', [(3, 216, 24, 24), (2, 64, 7, 9), (1, 4, 33, 31)])
def test_layout_convert(shape):
    x = torch.randn(*shape, device=DEV)
    cl = T.to_nhwc(x)
    assert cl.is_contiguous(memory_format=torch.channels_last) and torch.equal(cl, x)
    back = T.to_nchw(cl)
    assert back.is_contiguous() and torch.equal(back, x)
    assert T.to_nchw(x) is x and T.to_nhwc(cl) is cl
    odd = torch.randn(2, 6, 5, 5, device=DEV)                 # C % 4 != 0: torch's copy
    assert torch.equal(T.to_nchw(T.to_nhwc(odd)), odd)


def test_conv_bias_to_nchw():
    conv = torch.nn.Conv2d(8, 216, 3, 1, 1).to(DEV)
    x = torch.randn(2, 8, 12, 12, device=DEV)
    with torch.no_grad():
        want = conv(x)
        a = T.conv_bias_to_nchw(x, conv)                                             # NCHW convolution
        conv_cl = conv.to(memory_format=torch.channels_last)
        b = T.conv_bias_to_nchw(x.contiguous(memory_format=torch.channels_last), conv_cl)   # bias on the conversion
    assert a.is_contiguous() and b.is_contiguous()
    assert rel_err(a, want) <= 1e-3 and rel_err(b, want) <= 1e-3


def test_maxpool2x2_nhwc():
    x = torch.randn(3, 64, 10, 14, device=DEV)
    want = F.max_pool2d(x, 2, 2)
    got = T.maxpool2x2(x.contiguous(memory_format=torch.channels_last), torch.nn.MaxPool2d(2, 2))
    assert got.is_contiguous(memory_format=torch.channels_last) and torch.equal(got, want)
    odd = torch.randn(1, 6, 9, 9, device=DEV).contiguous(memory_format=torch.channels_last)     # torch path
    assert torch.equal(T.maxpool2x2(odd, torch.nn.MaxPool2d(2, 2)), F.max_pool2d(odd, 2, 2))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('case', ['leaky', 'relu', 'residual', 'residual_scaled', 'bias_only'])
@pytest.mark.parametrize('shape', [(3, 64, 10, 12), (2, 256, 5, 7), (1, 8, 3, 3), (2, 216, 9, 11)])
def test_bias_act_training_function(dtype, case, shape):
    """BiasActFunction (training epilogue: in-place forward, one-pass backward with the bias gradient) against the torch
    expression it replaces, channels-last fp32 and bf16."""
    if not M._lib.lib().mrefsr_bias_act_train_supported(shape[1], 1 if dtype == torch.bfloat16 else 0):
        pytest.skip('channel count not served by the 16-byte channel vectors')
    g = torch.Generator().manual_seed(3)
    cl = torch.channels_last
    x0 = torch.randn(*shape, generator=g).to(DEV, dtype).contiguous(memory_format=cl)
    r0 = torch.randn(*shape, generator=g).to(DEV, dtype).contiguous(memory_format=cl)
    b0 = torch.randn(shape[1], generator=g).to(DEV)
    go = torch.randn(*shape, generator=g).to(DEV, dtype).contiguous(memory_format=cl)
    act, slope, scale, use_res = {'leaky': (T.ACT_LEAKY, 0.1, 1.0, False), 'relu': (T.ACT_LEAKY, 0.0, 1.0, False),
                                  'residual': (T.ACT_NONE, 0.0, 1.0, True), 'residual_scaled': (T.ACT_NONE, 0.0, 0.5, True),
                                  'bias_only': (T.ACT_NONE, 0.0, 1.0, False)}[case]

    def run(fused):
        x = x0.clone().requires_grad_(True)
        r = r0.clone().requires_grad_(True)
        b = b0.clone().requires_grad_(True)
        pre = x * 1.0                      # a non-leaf the Function may overwrite (a convolution's output in the network)
        if fused:
            y = T.BiasActFunction.apply(pre, b, act, slope, r if use_res else None, scale)
        else:
            y = pre.float() + b.view(1, -1, 1, 1)
            if dtype == torch.bfloat16:
                y = y.to(dtype).float()    # torch rounds the bias add to bf16 before the activation
            if act == T.ACT_LEAKY:
                y = F.leaky_relu(y, slope)
            y = y * scale
            if use_res:
                y = y + r.float()
            y = y.to(dtype)
        y.backward(go)
        return y.detach().float(), x.grad.float(), b.grad.float(), (r.grad.float() if use_res else None)

    a, c = run(True), run(False)
    tol = 1e-6 if dtype == torch.float32 else 1.6e-2
    for u, v, name in zip(a, c, ('y', 'gx', 'gb', 'gr')):
        if v is None:
            assert u is None
            continue
        assert rel_err(u, v) <= tol, (name, rel_err(u, v))


def test_resblock_training_path_matches_torch():
    """ResidualBlockNoBN / ContentExtractor under autograd in channels-last: fused epilogues (T.TRAIN_FUSED) against the
    ordinary nn.Module expressions, outputs and every parameter gradient."""
    torch.manual_seed(0)
    net = ContentExtractor(n_blocks=2).to(DEV).to(memory_format=torch.channels_last).train()
    x = torch.randn(2, 3, 20, 24, device=DEV).contiguous(memory_format=torch.channels_last)
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    res = {}
    try:
        for fused in (False, True):
            T.TRAIN_FUSED = fused
            net.zero_grad(set_to_none=True)
            y = net(x)
            y.square().mean().backward()
            res[fused] = (y.detach(), {k: p.grad.clone() for k, p in net.named_parameters()})
    finally:
        T.TRAIN_FUSED = True
        torch.backends.cudnn.allow_tf32 = old_tf32
    assert rel_err(res[True][0], res[False][0]) <= 1e-5
    for k in res[False][1]:
        assert rel_err(res[True][1][k], res[False][1][k]) <= 1e-4, k


@pytest.mark.parametrize('shape', [(2, 64, 9, 13), (1, 216, 16, 16), (3, 8, 5, 5), (2, 260, 4, 7), (1, 192, 6, 10)])
def test_layout_convert_bf16_one_pass(shape):
    """bf16 channels-last <-> fp32 NCHW in one pass (training under autocast) against torch's two-step conversion:
    exact both ways (bf16 -> fp32 is exact; fp32 -> bf16 rounds to nearest even like torch)."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(*shape, generator=g).to(DEV)
    xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    a = T.to_nchw_f32(xb)
    assert a.dtype == torch.float32 and a.is_contiguous() and torch.equal(a, xb.float().contiguous())
    b = T.from_nchw_f32(x, torch.bfloat16, True)
    assert b.dtype == torch.bfloat16 and b.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(b, x.to(torch.bfloat16))
    # other combinations fall back to torch's conversions with the same values
    assert torch.equal(T.to_nchw_f32(x), x)
    assert torch.equal(T.from_nchw_f32(x, torch.float32, True), x)
