"""Host-side logic of the training path (no GPU): the fused epilogues and one-pass conversions apply only to CUDA
channels-last tensors under autograd; everything else keeps the plain torch expressions with the same values."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from mrefsr_b200 import trunk as T


def test_train_ok_is_false_off_the_gpu_and_conv_act_falls_back():
    conv = nn.Conv2d(8, 8, 3, 1, 1)
    x = torch.randn(2, 8, 6, 6).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert not T.train_ok(x)
    y = T.conv_act(x, conv, T.ACT_LEAKY, 0.1)
    ref = F.leaky_relu(conv(x), 0.1)
    assert torch.equal(y, ref)
    assert T.conv_bias_act_train(x, conv, T.ACT_LEAKY, 0.1) is None
    y.sum().backward()
    assert conv.bias.grad is not None and x.grad is not None


def test_train_fused_switch():
    x = torch.randn(1, 8, 4, 4)
    old = T.TRAIN_FUSED
    try:
        T.TRAIN_FUSED = False
        assert not T.train_ok(x)
    finally:
        T.TRAIN_FUSED = old


def test_one_pass_conversions_fall_back_with_equal_values():
    x = torch.randn(2, 8, 5, 7)
    xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    a = T.to_nchw_f32(xb)
    assert a.dtype == torch.float32 and a.is_contiguous() and torch.equal(a, xb.float().contiguous())
    b = T.from_nchw_f32(x, torch.bfloat16, True)
    assert b.dtype == torch.bfloat16 and b.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(b, x.to(torch.bfloat16))
    # with the leaky ReLU folded in, and its backward on the way back (gate = the activation's output)
    c = T.from_nchw_f32(x, torch.bfloat16, True, slope=0.1)
    assert torch.equal(c, F.leaky_relu(x, 0.1).to(torch.bfloat16))
    g = torch.randn(2, 8, 5, 7).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    d = T.to_nchw_f32(g, gate=c, slope=0.1)
    assert torch.equal(d, torch.where(c > 0, g, g * 0.1).float().contiguous())
    assert T.is_channels_last(xb) and not T.is_channels_last(x)


def test_batched_reference_forward_is_selected_only_for_equal_shapes():
    from mrefsr_b200.models import DynamicAggregationRestoration
    m = DynamicAggregationRestoration(ngf=64, n_blocks=1, groups=8)
    called = {}
    m._forward_refs_batched = lambda *a: called.setdefault('batched', True)
    keys = ('relu3_1', 'relu2_1', 'relu1_1')
    feats = [{k: torch.zeros(1, 4, 2, 2) for k in keys} for _ in range(2)]
    pres = [{k: torch.zeros(1, 9, 2, 2, 2) for k in keys} for _ in range(2)]
    assert m.forward(torch.zeros(1, 64, 2, 2), pres, feats) is True and called
    m.batch_refs = False
    called.clear()
    try:
        m.forward(torch.zeros(1, 64, 2, 2), pres, feats)     # the per-reference loop: CUDA-only ops further down
    except Exception:
        pass
    assert not called


def test_dynagg_nodes_refuse_cpu_tensors():
    """No CPU path behind the autograd nodes either: like the reference's extension (deform_conv_ext.cpp:124)."""
    import pytest
    from mrefsr_b200.dynagg import DynAgg
    for fused in (True, False):
        m = DynAgg(8, 8, 3, stride=1, padding=1, dilation=1, deform_groups=2, extra_offset_mask=True)
        m.fused_autograd = fused
        with pytest.raises(NotImplementedError):
            m([torch.randn(1, 8, 5, 5), torch.randn(1, 8, 5, 5)], torch.zeros(1, 9, 5, 5, 2))
    m = DynAgg(8, 8, 3, stride=2, padding=1, dilation=1, deform_groups=2, extra_offset_mask=False)
    with pytest.raises(NotImplementedError):        # the folded activation needs the fused node's configuration
        m(torch.randn(1, 8, 5, 5), torch.zeros(1, 9, 3, 3, 2), out_slope=0.1)
