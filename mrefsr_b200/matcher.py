"""Correspondence matcher: host-side mirror of basicsr/archs/ref_map_util.py and of the matching half of
basicsr/archs/corres_generation_arch.py, running on the sm_100a kernels of csrc/match.cu.

Same names, argument meaning and return contract as the reference:
    sample_patches(inputs, patch_size=3, stride=1)                        ref_map_util.py:4-23
    feature_match_index(feat_input, feat_ref, patch_size=3, input_stride=1, ref_stride=1,
                        is_norm=True, norm_input=False) -> (max_idx, max_val)   ref_map_util.py:26-86
plus the batched entry points the reference does not have (it loops in Python over batch items,
corres_generation_arch.py:53, and over references, multi_ref_restoration_model.py:287).
"""
import torch

from . import _lib

MATCH_AUTO, MATCH_TC_BF16X3, MATCH_TC_BF16, MATCH_FP32 = 0, 1, 2, 3
FLAG_NO_STRIP, FLAG_BASE_OFFSET, FLAG_NO_DIAG, FLAG_NO_BSTRIP = 0x100, 0x200, 0x400, 0x800
_MODES = {'auto': MATCH_AUTO, 'bf16x3': MATCH_TC_BF16X3, 'bf16': MATCH_TC_BF16, 'fp32': MATCH_FP32}


def _mode(mode):
    return _MODES[mode] if isinstance(mode, str) else int(mode)


def sample_patches(inputs, patch_size=3, stride=1):
    """[C,h,w] -> [C,ps,ps,N] row-major patches (a strided view, as in the reference)."""
    c, h, w = inputs.shape
    return inputs.unfold(1, patch_size, stride).unfold(2, patch_size, stride) \
        .reshape(c, -1, patch_size, patch_size).permute(0, 2, 3, 1)


def feature_match_index_batched(feat_input, feat_ref, patch_size=3, input_stride=1, ref_stride=1, is_norm=True,
                                norm_input=False, normalize_pixels=False, in_div=None, mode='auto'):
    """feat_input [n_in,C,h,w], feat_ref [n_pairs,C,h2,w2] -> (max_idx int64 [n_pairs,h',w'], max_val fp32).

    Pair p is matched against input (p // in_div) % n_in; in_div defaults to n_pairs // n_in, i.e. pairs laid
    out [B, R].  normalize_pixels applies F.normalize(x.reshape(C,-1), dim=0) to both first
    (corres_generation_arch.py:57-59)."""
    _lib.require_cuda(feat_input, feat_ref)
    if feat_input.dim() != 4 or feat_ref.dim() != 4 or feat_input.shape[1] != feat_ref.shape[1]:
        raise ValueError('expected [n,C,h,w] features with equal channel counts')
    from .trunk import to_nchw
    fi = to_nchw(feat_input.float())
    fr = to_nchw(feat_ref.float())
    n_in, c, h, w = fi.shape
    n_pairs, _, h2, w2 = fr.shape
    if in_div is None:
        if n_pairs % n_in:
            raise ValueError('n_pairs (%d) is not a multiple of n_in (%d); pass in_div' % (n_pairs, n_in))
        in_div = n_pairs // n_in
    if h < patch_size or w < patch_size or h2 < patch_size or w2 < patch_size:
        raise ValueError('feature map smaller than the patch')
    ho, wo = (h - patch_size) // input_stride + 1, (w - patch_size) // input_stride + 1
    lib = _lib.lib()
    m = _mode(mode)
    nbytes = lib.mrefsr_match_workspace_bytes(n_in, n_pairs, c, h, w, h2, w2, m)
    with torch.cuda.device(fi.device):
        ws, ws_bytes = _lib.workspace(nbytes, fi.device)
        idx = torch.empty(n_pairs, ho, wo, dtype=torch.int64, device=fi.device)
        val = torch.empty(n_pairs, ho, wo, dtype=torch.float32, device=fi.device)
        rc = lib.mrefsr_feature_match_batched(_lib.ptr(fi), _lib.ptr(fr), n_in, n_pairs, in_div, c, h, w, h2, w2,
                                              patch_size, input_stride, ref_stride, int(bool(is_norm)),
                                              int(bool(norm_input)), int(bool(normalize_pixels)), m, _lib.ptr(idx),
                                              _lib.ptr(val), ws, ws_bytes, _lib.stream_ptr(fi.device))
    _lib.check(rc, 'mrefsr_feature_match_batched')
    return idx, val


def feature_match_index(feat_input, feat_ref, patch_size=3, input_stride=1, ref_stride=1, is_norm=True,
                        norm_input=False, mode='auto'):
    """Drop-in for basicsr.archs.ref_map_util.feature_match_index: [C,h,w] x [C,h,w] ->
    (max_idx int64 [h',w'], max_val [h',w'])."""
    idx, val = feature_match_index_batched(feat_input.unsqueeze(0), feat_ref.unsqueeze(0), patch_size, input_stride,
                                           ref_stride, is_norm, norm_input, False, 1, mode)
    return idx[0], val[0].to(feat_input.dtype)


def pre_offsets(max_idx, scales=(1, 2, 4)):
    """max_idx int64 [n,h-2,w-2] -> tuple of [n,9,s*h,s*w,2] fp32 (x, y) for s in scales:
    index_to_flow + the nine zero-filled shifts of corres_generation_arch.py:30-47, :70-105."""
    _lib.require_cuda(max_idx)
    if max_idx.dtype != torch.int64 or max_idx.dim() != 3:
        raise ValueError('max_idx must be int64 [n, h-2, w-2]')
    mi = max_idx.contiguous()
    n, hp, wp = mi.shape
    h, w = hp + 2, wp + 2
    outs = {s: torch.empty(n, 9, s * h, s * w, 2, dtype=torch.float32, device=mi.device) for s in scales}
    for s in scales:
        if s not in (1, 2, 4):
            raise ValueError('scales must be drawn from (1, 2, 4)')
    with torch.cuda.device(mi.device):
        rc = _lib.lib().mrefsr_pre_offsets(_lib.ptr(mi), n, h, w, _lib.ptr(outs.get(1)), _lib.ptr(outs.get(2)),
                                           _lib.ptr(outs.get(4)), _lib.stream_ptr(mi.device))
    _lib.check(rc, 'mrefsr_pre_offsets')
    return tuple(outs[s] for s in scales)


def correspondence(dense_features1, dense_features2, patch_size=3, stride=1, mode='auto'):
    """The matching half of CorrespondenceGenerationArch.forward (corres_generation_arch.py:49-111) for a whole
    batch at once: [B,C,h,w] x [B,C,h,w] -> pre_offset dict {'relu3_1','relu2_1','relu1_1'} of [B,9,s*h,s*w,2]."""
    idx, _ = feature_match_index_batched(dense_features1, dense_features2, patch_size, stride, stride, True, True,
                                         normalize_pixels=True, in_div=1, mode=mode)
    o1, o2, o4 = pre_offsets(idx)
    return {'relu3_1': o1, 'relu2_1': o2, 'relu1_1': o4}
