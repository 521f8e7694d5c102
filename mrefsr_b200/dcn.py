"""Modulated deformable convolution (DCNv2): host-side mirror of basicsr/ops/dcn/deform_conv.py on the sm_100a
kernels of csrc/dcn.cu / csrc/dcn_tc.cu.

Same operator API as the reference (deform_conv.py:121-188, :289-379):
    ModulatedDeformConvFunction.apply(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
                                      groups=1, deformable_groups=1)
    modulated_deform_conv = ModulatedDeformConvFunction.apply
    ModulatedDeformConv, ModulatedDeformConvPack  (same parameters / state-dict keys / init)
stride / padding / dilation may be ints (basicsr) or pairs (mmcv.ops, see mmcv_ops.py).
DCNv1 (DeformConv, deform_conv) is exported by the reference package but is not on the MRefSR path; the
names exist here and raise NotImplementedError.
"""
import math

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair, _single

from . import _lib

DCN_AUTO, DCN_FP32, DCN_TF32 = 0, 1, 2
_MODES = {'auto': DCN_AUTO, 'fp32': DCN_FP32, 'tf32': DCN_TF32}
_default_mode = DCN_AUTO


def set_default_mode(mode):
    """'auto' | 'fp32' (exact, CUDA cores) | 'tf32' (tcgen05)."""
    global _default_mode
    _default_mode = _MODES[mode] if isinstance(mode, str) else int(mode)


def _out_hw(h, w, kh, kw, stride, padding, dilation):
    ho = (h + 2 * padding[0] - (dilation[0] * (kh - 1) + 1)) // stride[0] + 1
    wo = (w + 2 * padding[1] - (dilation[1] * (kw - 1) + 1)) // stride[1] + 1
    return ho, wo


def _check_shapes(input, offset, mask, weight, groups, dg, ho, wo):
    b, c = input.shape[:2]
    co, cg, kh, kw = weight.shape
    if c != cg * groups:
        raise RuntimeError("Input shape and kernel channels won't match: (%d vs %d)." % (c, cg * groups))
    if tuple(offset.shape) != (b, 2 * dg * kh * kw, ho, wo):
        raise RuntimeError('offset shape %s, expected %s' % (tuple(offset.shape), (b, 2 * dg * kh * kw, ho, wo)))
    if tuple(mask.shape) != (b, dg * kh * kw, ho, wo):
        raise RuntimeError('mask shape %s, expected %s' % (tuple(mask.shape), (b, dg * kh * kw, ho, wo)))


def dcn_forward_raw(input, offset, mask, weight, bias, stride, padding, dilation, groups, dg, mode=None):
    """Forward on contiguous fp32 CUDA tensors -> output [B,Co,Ho,Wo] (no autograd)."""
    lib = _lib.lib()
    b, c, h, w = input.shape
    co, _, kh, kw = weight.shape
    ho, wo = _out_hw(h, w, kh, kw, stride, padding, dilation)
    _check_shapes(input, offset, mask, weight, groups, dg, ho, wo)
    m = _default_mode if mode is None else (_MODES[mode] if isinstance(mode, str) else int(mode))
    out = torch.empty(b, co, ho, wo, dtype=torch.float32, device=input.device)
    with torch.cuda.device(input.device):
        nbytes = lib.mrefsr_dcn_workspace_bytes(b, c, h, w, co, kh, kw, stride[0], stride[1], padding[0], padding[1],
                                                dilation[0], dilation[1], groups, dg, m, 0)
        ws, ws_bytes = _lib.workspace(nbytes, input.device)
        rc = lib.mrefsr_modulated_deform_conv_forward(
            _lib.ptr(input), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(offset), _lib.ptr(mask), _lib.ptr(out), b, c, h,
            w, co, kh, kw, stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1], groups, dg,
            int(bias is not None), m, ws, ws_bytes, _lib.stream_ptr(input.device))
    _lib.check(rc, 'mrefsr_modulated_deform_conv_forward')
    return out


def dynagg_dcn_forward(input, conv_out, max_idx, flow_scale, weight, bias, deformable_groups):
    """Fused DynAgg forward for inference (no autograd): DCNv2 (3x3, stride 1, pad 1) whose offsets and masks are
    assembled inside the gather from the raw conv_offset_mask output and the matcher's arg-max map
    (ref_mrapa_restoration_arch.py:45-76 + corres_generation_arch.py:30-47, :70-105 for one scale).
    input [B,C,H,W], conv_out [B,3*dg*9,H,W], max_idx int64 [B,H/s-2,W/s-2] -> [B,Co,H,W]."""
    _lib.require_cuda(input, conv_out, max_idx, weight, bias)
    lib = _lib.lib()
    x, co_, wgt = (t.contiguous().float() for t in (input, conv_out, weight))
    bs = bias.contiguous().float() if bias is not None else None
    mi = max_idx.contiguous()
    b, c, h, w = x.shape
    co = wgt.shape[0]
    dg = deformable_groups
    if tuple(co_.shape) != (b, 3 * dg * 9, h, w):
        raise RuntimeError('conv_out shape %s, expected %s' % (tuple(co_.shape), (b, 3 * dg * 9, h, w)))
    if mi.dtype != torch.int64 or tuple(mi.shape) != (b, h // flow_scale - 2, w // flow_scale - 2):
        raise RuntimeError('max_idx must be int64 [%d,%d,%d]' % (b, h // flow_scale - 2, w // flow_scale - 2))
    out = torch.empty(b, co, h, w, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        nbytes = lib.mrefsr_dcn_workspace_bytes(b, c, h, w, co, 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, DCN_TF32, 0)
        ws, ws_bytes = _lib.workspace(nbytes, x.device)
        rc = lib.mrefsr_dynagg_dcn_forward(_lib.ptr(x), _lib.ptr(wgt), _lib.ptr(bs), _lib.ptr(co_), _lib.ptr(mi),
                                           int(flow_scale), _lib.ptr(out), b, c, h, w, co, dg, int(bs is not None), ws,
                                           ws_bytes, _lib.stream_ptr(x.device))
    _lib.check(rc, 'mrefsr_dynagg_dcn_forward')
    return out


class ModulatedDeformConvFunction(Function):

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1):
        ctx.stride, ctx.padding, ctx.dilation = _pair(stride), _pair(padding), _pair(dilation)
        ctx.groups = groups
        ctx.deformable_groups = deformable_groups
        ctx.with_bias = bias is not None
        if not input.is_cuda:
            raise NotImplementedError  # same as deform_conv.py:143-144
        _lib.require_cuda(offset, mask, weight, bias)
        ctx.in_dtype = input.dtype
        x, off, msk, wgt = (t.contiguous().float() for t in (input, offset, mask, weight))
        bs = bias.contiguous().float() if bias is not None else None
        if weight.requires_grad or mask.requires_grad or offset.requires_grad or input.requires_grad:
            ctx.save_for_backward(x, off, msk, wgt)
        out = dcn_forward_raw(x, off, msk, wgt, bs, ctx.stride, ctx.padding, ctx.dilation, groups, deformable_groups)
        return out.to(input.dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        x, off, msk, wgt = ctx.saved_tensors
        go = grad_output.contiguous().float()
        lib = _lib.lib()
        b, c, h, w = x.shape
        co, _, kh, kw = wgt.shape
        need_gi = ctx.needs_input_grad[0]
        grad_input = torch.empty_like(x) if need_gi else None
        grad_offset = torch.empty_like(off)
        grad_mask = torch.empty_like(msk)
        grad_weight = torch.zeros_like(wgt)
        grad_bias = torch.zeros(co, dtype=torch.float32, device=x.device) if ctx.with_bias else None
        s, p, d = ctx.stride, ctx.padding, ctx.dilation
        with torch.cuda.device(x.device):
            nbytes = lib.mrefsr_dcn_workspace_bytes(b, c, h, w, co, kh, kw, s[0], s[1], p[0], p[1], d[0], d[1],
                                                    ctx.groups, ctx.deformable_groups, _default_mode, 1)
            ws, ws_bytes = _lib.workspace(nbytes, x.device)
            rc = lib.mrefsr_modulated_deform_conv_backward(
                _lib.ptr(x), _lib.ptr(wgt), _lib.ptr(off), _lib.ptr(msk), _lib.ptr(go), _lib.ptr(grad_input),
                _lib.ptr(grad_weight), _lib.ptr(grad_bias), _lib.ptr(grad_offset), _lib.ptr(grad_mask), b, c, h, w, co,
                kh, kw, s[0], s[1], p[0], p[1], d[0], d[1], ctx.groups, ctx.deformable_groups, int(ctx.with_bias),
                _default_mode, ws, ws_bytes, _lib.stream_ptr(x.device))
        _lib.check(rc, 'mrefsr_modulated_deform_conv_backward')
        dt = ctx.in_dtype
        cast = (lambda t: None if t is None else t.to(dt))
        return (cast(grad_input), cast(grad_offset), cast(grad_mask), cast(grad_weight), cast(grad_bias), None, None,
                None, None, None)

    @staticmethod
    def _infer_shape(ctx, input, weight):
        n = input.size(0)
        kh, kw = weight.shape[2:4]
        ho, wo = _out_hw(input.shape[2], input.shape[3], kh, kw, _pair(ctx.stride), _pair(ctx.padding),
                         _pair(ctx.dilation))
        return n, weight.size(0), ho, wo


modulated_deform_conv = ModulatedDeformConvFunction.apply


def deform_conv(*args, **kwargs):
    raise NotImplementedError('DCNv1 (deform_conv) is not on the MRefSR alignment path and is not built here')


class DeformConv(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('DCNv1 (DeformConv) is not on the MRefSR alignment path and is not built here')


DeformConvPack = DeformConv


class ModulatedDeformConv(nn.Module):
    """Same constructor, parameters (weight, bias) and init as deform_conv.py:289-333."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.with_bias = bias
        self.transposed = False
        self.output_padding = _single(0)
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.init_weights()

    def init_weights(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                                     self.groups, self.deformable_groups)


class ModulatedDeformConvPack(ModulatedDeformConv):
    """deform_conv.py:336-379: adds the zero-initialised conv_offset that predicts offsets and masks."""

    _version = 2

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels,
                                     self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
                                     kernel_size=self.kernel_size, stride=_pair(self.stride),
                                     padding=_pair(self.padding), dilation=_pair(self.dilation), bias=True)
        self.init_weights()

    def init_weights(self):
        super().init_weights()
        if hasattr(self, 'conv_offset'):
            self.conv_offset.weight.data.zero_()
            self.conv_offset.bias.data.zero_()

    def forward(self, x):
        out = self.conv_offset(x)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        mask = torch.sigmoid(mask)
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                                     self.groups, self.deformable_groups)


# ---------------------------------------------------------------------------------------------------
# the five functions the reference's pybind module exports (deform_conv_ext.cpp:150-164), same argument order,
# so that basicsr/ops/dcn/deform_conv.py can bind `deform_conv_ext = mrefsr_b200.dcn.ext` unchanged
# ---------------------------------------------------------------------------------------------------
class _Ext:
    @staticmethod
    def modulated_deform_conv_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w,
                                      stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                                      deformable_group, with_bias):
        if not input.is_cuda:
            raise RuntimeError('modulated deform conv is not implemented on CPU')  # deform_conv_ext.cpp:124
        if not input.is_contiguous():
            raise RuntimeError('input tensor has to be contiguous')                # deform_conv_cuda.cpp:497
        if not weight.is_contiguous():
            raise RuntimeError('weight tensor has to be contiguous')
        if weight.shape[2] != kernel_h or weight.shape[3] != kernel_w:
            raise RuntimeError("Input shape and kernel shape won't match: (%d x %d vs %d x %d)." %
                               (kernel_h, kernel_w, weight.shape[2], weight.shape[3]))
        out = dcn_forward_raw(input.float(), offset.contiguous().float(), mask.contiguous().float(), weight.float(),
                              bias.contiguous().float() if with_bias else None, (stride_h, stride_w), (pad_h, pad_w),
                              (dilation_h, dilation_w), group, deformable_group)
        output.view(out.shape).copy_(out)

    @staticmethod
    def modulated_deform_conv_backward(input, weight, bias, ones, offset, mask, columns, grad_input, grad_weight,
                                       grad_bias, grad_offset, grad_mask, grad_output, kernel_h, kernel_w, stride_h,
                                       stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group,
                                       with_bias):
        if not input.is_cuda:
            raise RuntimeError('modulated deform conv is not implemented on CPU')  # deform_conv_ext.cpp:146
        lib = _lib.lib()
        b, c, h, w = input.shape
        co = weight.shape[0]
        for t in (input, weight, offset, mask, grad_output, grad_input, grad_weight, grad_offset, grad_mask):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError('deform_conv_ext shim: fp32 contiguous tensors only')
        with torch.cuda.device(input.device):
            nbytes = lib.mrefsr_dcn_workspace_bytes(b, c, h, w, co, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                                                    dilation_h, dilation_w, group, deformable_group, _default_mode, 1)
            ws, ws_bytes = _lib.workspace(nbytes, input.device)
            rc = lib.mrefsr_modulated_deform_conv_backward(
                _lib.ptr(input), _lib.ptr(weight), _lib.ptr(offset), _lib.ptr(mask), _lib.ptr(grad_output),
                _lib.ptr(grad_input), _lib.ptr(grad_weight), _lib.ptr(grad_bias) if with_bias else None,
                _lib.ptr(grad_offset), _lib.ptr(grad_mask), b, c, h, w, co, kernel_h, kernel_w, stride_h, stride_w,
                pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, int(with_bias), _default_mode, ws,
                ws_bytes, _lib.stream_ptr(input.device))
        _lib.check(rc, 'mrefsr_modulated_deform_conv_backward')

    @staticmethod
    def deform_conv_forward(*a, **k):
        raise NotImplementedError('DCNv1 is not on the MRefSR alignment path and is not built here')

    deform_conv_backward_input = deform_conv_forward
    deform_conv_backward_parameters = deform_conv_forward


ext = _Ext()
