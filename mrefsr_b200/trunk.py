"""Inference fast path for the plain-convolution trunk around the hot path.

The convolutions stay cuDNN library calls (``F.conv2d`` without bias); what torch runs after each of them as
separate kernels -- a non-vectorised broadcast bias add, the activation, the residual add -- is one streaming
pass of csrc/trunk.cu.  Used only when autograd is off and the tensors are contiguous fp32 CUDA tensors;
otherwise the caller keeps the ordinary ``nn.Module`` path (training, CPU construction, other dtypes).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

ACT_NONE, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2


def layout_of(t):
    """0: dense NCHW, 1: dense channels-last (NHWC), None: neither."""
    if t.is_contiguous():
        return 0
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return 1
    return None


def fast_ok(*tensors):
    """True when the fused inference path applies to these tensors (no autograd, no autocast, dense fp32 CUDA)."""
    return (not torch.is_grad_enabled()) and (not torch.is_autocast_enabled()) and all(
        t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and layout_of(t) is not None for t in tensors)


def dense(t):
    """t if it is dense NCHW or channels-last, else a contiguous copy."""
    return t if layout_of(t) is not None else t.contiguous()


def bias_act_(x, bias=None, act=ACT_NONE, slope=0.0, residual=None, slope_dev=None, scale=1.0, res_div=1,
              res_pre=False):
    """x[b,c] = act(x[b,c] + bias[c]) * scale + residual[b // res_div, c], in place; returns x.
    res_pre: x[b,c] = act(x[b,c] + bias[c] + residual[b // res_div, c]) * scale instead."""
    b, c, h, w = x.shape
    cl = layout_of(x)
    if cl is None:
        raise ValueError('x must be dense (NCHW or channels_last)')
    if residual is not None:
        if b % res_div or tuple(residual.shape) != (b // res_div, c, h, w):
            raise ValueError('residual must have shape [B / res_div, C, H, W] of x')
        if layout_of(residual) != cl:
            residual = to_nhwc(residual) if cl else to_nchw(residual)
    with torch.cuda.device(x.device):
        rc = _lib.lib().mrefsr_bias_act(_lib.ptr(x), _lib.ptr(bias), _lib.ptr(slope_dev),
                                        0 if slope_dev is None else slope_dev.numel(), _lib.ptr(residual),
                                        int(res_div), int(res_pre), b, c, h * w, cl, act, float(slope), float(scale),
                                        _lib.stream_ptr(x.device))
    _lib.check(rc, 'mrefsr_bias_act')
    return x


def _convert(t, to_cl, bias=None):
    b, c, h, w = t.shape
    out = torch.empty(b, c, h, w, dtype=t.dtype, device=t.device,
                      memory_format=torch.channels_last if to_cl else torch.contiguous_format)
    with torch.cuda.device(t.device):
        rc = _lib.lib().mrefsr_layout_convert(_lib.ptr(t), _lib.ptr(out), _lib.ptr(bias), b, c, h * w, int(to_cl),
                                              _lib.stream_ptr(t.device))
    _lib.check(rc, 'mrefsr_layout_convert')
    return out


def _convertible(t):
    # (inside an autograd.Function's forward / backward grad mode is off and the raw kernel is safe on any tensor)
    return (t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] % 4 == 0 and t.shape[0] <= 65535
            and t.numel() > 0 and (not t.requires_grad or not torch.is_grad_enabled()))


def to_nchw(t):
    """Dense NCHW tensor with the values of t; a channels-last fp32 CUDA tensor goes through csrc/trunk.cu's tiled
    transpose (torch's strided copy runs at a fraction of HBM speed)."""
    if t.is_contiguous():
        return t
    if _convertible(t) and layout_of(t) == 1:
        return _convert(t, False)
    return t.contiguous()


def _bf16_convertible(t):
    return (t.is_cuda and t.dim() == 4 and t.shape[1] % 4 == 0 and t.shape[0] <= 65535 and t.numel() > 0 and
            (not t.requires_grad or not torch.is_grad_enabled()))


def to_nchw_f32(t, gate=None, slope=1.0):
    """Dense fp32 NCHW tensor with the values of t (what the alignment kernels read).  A bf16 channels-last tensor --
    a convolution's output under autocast -- is converted in ONE pass (torch: a cast, then a strided copy).
    gate / slope: t * (gate > 0 ? 1 : slope), i.e. the backward of a leaky ReLU whose output is `gate`, on the way."""
    fast = t.dtype == torch.bfloat16 and layout_of(t) == 1 and not t.is_contiguous() and _bf16_convertible(t)
    if fast and (gate is None or (gate.dtype == torch.bfloat16 and gate.shape == t.shape and layout_of(gate) == 1
                                  and not gate.is_contiguous())):
        b, c, h, w = t.shape
        out = torch.empty(b, c, h, w, dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            rc = _lib.lib().mrefsr_layout_convert_bf16_act(_lib.ptr(t), _lib.ptr(out), _lib.ptr(gate), float(slope), b, c,
                                                           h * w, 0, _lib.stream_ptr(t.device))
        _lib.check(rc, 'mrefsr_layout_convert_bf16_act')
        return out
    if gate is not None:
        t = torch.where(gate > 0, t, t * slope)
    return to_nchw(t.float())


def from_nchw_f32(t, dtype, channels_last, slope=1.0):
    """The way back: dense fp32 NCHW `t` as a tensor of `dtype` in the given layout (one pass to bf16 channels-last;
    otherwise torch's conversions).  slope != 1 applies a leaky ReLU on the way."""
    if (dtype == torch.bfloat16 and channels_last and t.dtype == torch.float32 and t.is_contiguous() and
            _bf16_convertible(t)):
        b, c, h, w = t.shape
        out = torch.empty(b, c, h, w, dtype=torch.bfloat16, device=t.device, memory_format=torch.channels_last)
        with torch.cuda.device(t.device):
            rc = _lib.lib().mrefsr_layout_convert_bf16_act(_lib.ptr(t), _lib.ptr(out), None, float(slope), b, c, h * w, 1,
                                                           _lib.stream_ptr(t.device))
        _lib.check(rc, 'mrefsr_layout_convert_bf16_act')
        return out
    if slope != 1.0:
        t = F.leaky_relu(t, slope)
    if channels_last and dtype == torch.float32:
        return to_nhwc(t)
    return t.to(dtype=dtype, memory_format=torch.channels_last if channels_last else torch.contiguous_format)


def is_channels_last(t):
    return t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def conv_bias_to_nchw(x, conv):
    """conv(x) + bias as a dense NCHW tensor.  From a channels-last convolution the bias rides on the layout
    conversion (one pass instead of two); otherwise it is the ordinary conv + bias epilogue."""
    y = conv_raw(x, conv)
    if layout_of(y) == 1 and not y.is_contiguous() and _convertible(y):
        return _convert(y, False, conv.bias)
    return bias_act_(to_nchw(dense(y)), conv.bias)


def to_nhwc(t):
    """torch.channels_last tensor with the values of t."""
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return t
    if _convertible(t) and layout_of(t) == 0:
        return _convert(t, True)
    return t.contiguous(memory_format=torch.channels_last)


def attn_modulate_(refs, attn_mul, attn_add, bias_mul=None, bias_add=None):
    """refs = refs * sigmoid(attn_mul + bias_mul[c]) * 2 + (attn_add + bias_add[c]), in place; returns refs."""
    b, c, h, w = refs.shape
    cl = layout_of(refs)
    if cl is None:
        raise ValueError('refs must be dense (NCHW or channels_last)')
    attn_mul, attn_add = ((to_nhwc(t) if cl else to_nchw(t)) for t in (attn_mul, attn_add))
    with torch.cuda.device(refs.device):
        rc = _lib.lib().mrefsr_attn_modulate(_lib.ptr(refs), _lib.ptr(attn_mul), _lib.ptr(attn_add), _lib.ptr(bias_mul),
                                             _lib.ptr(bias_add), b, c, h * w, cl, _lib.stream_ptr(refs.device))
    _lib.check(rc, 'mrefsr_attn_modulate')
    return refs


def conv_raw(x, conv):
    """The convolution alone (no bias): cuDNN."""
    return F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)


def conv_bias_act(x, conv, act=ACT_NONE, slope=0.0, residual=None, prelu=None, scale=1.0):
    """act(conv(x) + bias) * scale + residual with the epilogue in one pass.  `prelu`: an nn.PReLU module.
    conv + bias + ReLU on a channels-last tensor goes to cuDNN's own fused epilogue (cudnnConvolutionBiasActivation:
    measured 63 vs 103 us for 16x64x160x160 on B200; in NCHW it is no faster than conv + one glue pass)."""
    if (act == ACT_LEAKY and slope == 0.0 and prelu is None and residual is None and scale == 1.0
            and conv.bias is not None and layout_of(x) == 1 and conv.padding_mode == 'zeros'
            and not isinstance(conv.padding, str)):
        return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation,
                                            conv.groups)
    y = dense(conv_raw(x, conv))
    if prelu is not None:
        return bias_act_(y, conv.bias, ACT_LEAKY, 0.0, residual, prelu.weight, scale)
    return bias_act_(y, conv.bias, act, slope, residual, None, scale)


def maxpool2x2(x, m=None):
    """nn.MaxPool2d(2, 2) on a channels-last fp32 CUDA tensor through csrc/trunk.cu; anything else through torch."""
    b, c, h, w = x.shape
    plain = m is None or (m.kernel_size in (2, (2, 2)) and m.stride in (2, (2, 2)) and m.padding in (0, (0, 0))
                          and m.dilation in (1, (1, 1)) and not m.ceil_mode and not m.return_indices)
    if not (plain and layout_of(x) == 1 and x.is_cuda and x.dtype == torch.float32 and c % 4 == 0 and h % 2 == 0
            and w % 2 == 0 and not x.requires_grad):
        return m(x) if m is not None else F.max_pool2d(x, 2, 2)
    out = torch.empty(b, c, h // 2, w // 2, dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        rc = _lib.lib().mrefsr_maxpool2x2_nhwc(_lib.ptr(x), _lib.ptr(out), b, c, h, w, _lib.stream_ptr(x.device))
    _lib.check(rc, 'mrefsr_maxpool2x2_nhwc')
    return out


def run_sequential(seq, x, taps=None, out=None):
    """nn.Sequential of Conv2d / ReLU / LeakyReLU / MaxPool2d (VGG-style): every Conv2d that is followed by an
    activation runs fused.  `taps`: names whose outputs are collected into `out` (a dict), as the reference's
    VGGFeatureExtractor.forward does (vgg_arch.py:141-161)."""
    items = list(seq._modules.items())
    i = 0
    while i < len(items):
        name, m = items[i]
        nxt = items[i + 1] if i + 1 < len(items) else (None, None)
        if isinstance(m, nn.Conv2d) and isinstance(nxt[1], (nn.ReLU, nn.LeakyReLU)):
            slope = nxt[1].negative_slope if isinstance(nxt[1], nn.LeakyReLU) else 0.0
            if taps is not None and name in taps:          # a tap on the pre-activation output: keep it separate
                x = conv_bias_act(x, m)
                out[name] = x.clone()
                x = bias_act_(x, None, ACT_LEAKY, slope)
            else:
                x = conv_bias_act(x, m, ACT_LEAKY, slope)
            if taps is not None and nxt[0] in taps:
                out[nxt[0]] = x                            # later layers never write into x (convs / pools allocate)
            i += 2
            continue
        if isinstance(m, nn.Conv2d):
            x = conv_bias_act(x, m)
        elif isinstance(m, nn.MaxPool2d):
            x = maxpool2x2(x, m)
        else:
            x = m(x)
        if taps is not None and name in taps:
            out[name] = x
        i += 1
    return x


# ---------------------------------------------------------------------------------------------------------------------
# training path of the same epilogue (autograd on): channels-last fp32 / bf16 activations
# ---------------------------------------------------------------------------------------------------------------------
_CL = torch.channels_last
TRAIN_FUSED = True      # False: modules keep their ordinary torch expressions under autograd (cross-check / A-B)


def train_ok(*tensors):
    """True when the fused TRAINING epilogue applies: autograd on, dense channels-last CUDA tensors of one dtype
    (fp32, or bf16 under autocast), channel count the 16-byte channel vectors of csrc/trunk.cu can serve."""
    if not TRAIN_FUSED or not torch.is_grad_enabled() or not tensors:
        return False
    dt = tensors[0].dtype
    if dt not in (torch.float32, torch.bfloat16):
        return False
    for t in tensors:
        if not (t.is_cuda and t.dim() == 4 and t.dtype == dt and layout_of(t) == 1):
            return False
    return True


class BiasActFunction(torch.autograd.Function):
    """y = act(x + bias[c]) * scale + residual on a convolution's bias-free output, in place; backward = ONE pass that
    produces grad_x = grad_y * act'(y) * scale and grad_bias (torch: activation backward, then a reduction over
    grad_output per convolution).  act: ACT_NONE or ACT_LEAKY (ReLU = slope 0)."""

    @staticmethod
    def forward(ctx, x, bias, act, slope, residual, scale):
        lib = _lib.lib()
        b, c, h, w = x.shape
        dt = 1 if x.dtype == torch.bfloat16 else 0
        if act == ACT_LEAKY and (residual is not None or scale <= 0):
            raise ValueError('BiasActFunction: an activation excludes residual / non-positive scale')
        with torch.cuda.device(x.device):
            rc = lib.mrefsr_bias_act_train_forward(_lib.ptr(x), _lib.ptr(bias), _lib.ptr(residual), b * h * w, c, dt, act,
                                                   float(slope), float(scale), _lib.stream_ptr(x.device))
        _lib.check(rc, 'mrefsr_bias_act_train_forward')
        ctx.mark_dirty(x)
        ctx.act, ctx.slope, ctx.scale, ctx.has_res = act, float(slope), float(scale), residual is not None
        ctx.bias_dtype = bias.dtype
        if act == ACT_LEAKY:
            ctx.save_for_backward(x)
        return x

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        lib = _lib.lib()
        if layout_of(g) != 1:
            g = g.contiguous(memory_format=_CL)
        y = ctx.saved_tensors[0] if ctx.act == ACT_LEAKY else None
        if y is not None and y.dtype != g.dtype:
            g = g.to(y.dtype)
        b, c, h, w = g.shape
        dt = 1 if g.dtype == torch.bfloat16 else 0
        plain = ctx.act == ACT_NONE and ctx.scale == 1.0
        gin = g if plain else torch.empty_like(g, memory_format=_CL)
        gb = torch.empty(c, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            part = torch.empty(lib.mrefsr_bias_act_train_blocks() * c, dtype=torch.float32, device=g.device)
            rc = lib.mrefsr_bias_act_train_backward(_lib.ptr(g), _lib.ptr(y), None if plain else _lib.ptr(gin), _lib.ptr(gb),
                                                    _lib.ptr(part), b * h * w, c, dt, ctx.act, ctx.slope, ctx.scale,
                                                    _lib.stream_ptr(g.device))
        _lib.check(rc, 'mrefsr_bias_act_train_backward')
        return gin, gb.to(ctx.bias_dtype), None, None, (g if ctx.has_res else None), None


def conv_bias_act_train(x, conv, act=ACT_NONE, slope=0.0, residual=None, scale=1.0):
    """Training-path conv -> [bias, activation, * scale, + residual]: the convolution runs bias-free (cuDNN, autocast as
    usual) and the epilogue is BiasActFunction.  Returns None when the shapes / layouts are not served (the caller keeps
    its ordinary nn.Module expression)."""
    if conv.bias is None or not train_ok(x):
        return None
    y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
    dt = 1 if y.dtype == torch.bfloat16 else 0
    if not train_ok(y) or not _lib.lib().mrefsr_bias_act_train_supported(y.shape[1], dt):
        y = y + conv.bias.view(1, -1, 1, 1).to(y.dtype)          # same arithmetic through torch
        if act == ACT_LEAKY:
            y = F.leaky_relu(y, slope)
        y = y * scale if scale != 1.0 else y
        return y + residual if residual is not None else y
    if residual is not None and not (train_ok(residual) and residual.dtype == y.dtype and residual.shape == y.shape):
        y = BiasActFunction.apply(y, conv.bias, act, slope, None, scale)
        return y + residual
    return BiasActFunction.apply(y, conv.bias, act, slope, residual, scale)


def conv_act(x, conv, act=ACT_NONE, slope=0.0):
    """act(conv(x)) for the autograd path: the fused training epilogue where it applies, the torch expression otherwise."""
    if train_ok(x):
        y = conv_bias_act_train(x, conv, act, slope)
        if y is not None:
            return y
    y = conv(x)
    return F.leaky_relu(y, slope) if act == ACT_LEAKY else y
