"""mrefsr_b200 -- B200-native (sm_100a) implementation of MRefSR's reference-alignment hot path:
correspondence matcher, DCNv2 (modulated deformable convolution) and multi-reference attention fusion,
behind the reference's own operator API.  Hand-written CUDA in csrc/, reached through the C ABI of
include/mrefsr_b200.h; no CPU fallback (ops raise if the library is missing or tensors are not CUDA).
"""
from . import _lib  # noqa: F401
from .matcher import (sample_patches, feature_match_index, feature_match_index_batched, pre_offsets,  # noqa: F401
                      correspondence)
from .dcn import (ModulatedDeformConvFunction, modulated_deform_conv, ModulatedDeformConv,  # noqa: F401
                  ModulatedDeformConvPack, DeformConvFunction, DeformConv, DeformConvPack, deform_conv)
from .mmcv_ops import ModulatedDeformConv2d, modulated_deform_conv2d  # noqa: F401
from .fusion import MRAPAFusion, mrapa_attention  # noqa: F401
from .dynagg import DynAgg  # noqa: F401
from .archs import CorrespondenceGenerationArch, VGGFeatureExtractor  # noqa: F401

__version__ = '0.1.0'
