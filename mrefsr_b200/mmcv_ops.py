"""mmcv.ops-compatible shim.

MRefSR's DynAgg does ``from mmcv.ops import ModulatedDeformConv2d, modulated_deform_conv2d``
(basicsr/archs/ref_mrapa_restoration_arch.py:5, ref_restoration_arch.py:5); mmcv is un-vendored and unpinned in
the reference.  ``install()`` registers this module as ``mmcv.ops`` so the reference arch files import our op.
"""
import math
import sys
import types

import torch
from torch import nn
from torch.nn.modules.utils import _pair

from .dcn import modulated_deform_conv


def modulated_deform_conv2d(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                            deform_groups=1):
    """mmcv signature: pair-valued stride / padding / dilation, ``deform_groups`` spelling."""
    return modulated_deform_conv(input, offset, mask, weight, bias, _pair(stride), _pair(padding), _pair(dilation),
                                 groups, deform_groups)


class ModulatedDeformConv2d(nn.Module):
    """Attributes DynAgg relies on (ref_mrapa_restoration_arch.py:27-38, :74-76): in_channels, out_channels,
    kernel_size / stride / padding / dilation (pairs), groups, deform_groups, weight, bias."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deform_groups=1, bias=True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.deform_groups = deform_groups
        self.transposed = False
        self.output_padding = (0,)
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
        return modulated_deform_conv2d(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                       self.dilation, self.groups, self.deform_groups)


def install(force=False):
    """Register this shim as ``mmcv.ops`` (and a bare ``mmcv`` package if mmcv is not installed)."""
    if 'mmcv.ops' in sys.modules and not force:
        return sys.modules['mmcv.ops']
    ops = types.ModuleType('mmcv.ops')
    ops.ModulatedDeformConv2d = ModulatedDeformConv2d
    ops.modulated_deform_conv2d = modulated_deform_conv2d
    mmcv = sys.modules.get('mmcv') or types.ModuleType('mmcv')
    mmcv.ops = ops
    sys.modules['mmcv'] = mmcv
    sys.modules['mmcv.ops'] = ops
    return ops
