"""Oracle for the correspondence matcher (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates basicsr/archs/ref_map_util.py:
  * sample_patches        :4-23   -> sample_patches_oracle
  * feature_match_index   :26-86  -> feature_match_index_oracle

The reference evaluates sim[q, p] with F.conv2d using the (normalised) reference
patches as filters, takes max/argmax over q (first maximal index wins, :69; the
chunk merge at :71-76 uses strict '>' so the lowest q still wins) and finally
divides by the input-patch norm (:78-84).  Here the same quantity is written as
an explicit [N_ref, N_in] matrix product over unfolded patches so that tests can
also read the top-2 gap; ``use_conv=True`` evaluates it with F.conv2d exactly as
the reference does (bit-identical to the reference on the same CPU).
"""
import torch
import torch.nn.functional as F


def sample_patches_oracle(feat, patch_size=3, stride=1):
    """[C,h,w] -> [C,ps,ps,N], row-major patch order (ref_map_util.py:18-23)."""
    c, h, w = feat.shape
    cols = F.unfold(feat.unsqueeze(0), kernel_size=patch_size, stride=stride)  # [1, C*ps*ps, N]
    return cols.view(c, patch_size, patch_size, -1)


def similarity_volume(feat_input, feat_ref, patch_size=3, input_stride=1, ref_stride=1,
                      is_norm=True, norm_input=False, dtype=None):
    """Full similarity matrix sim[q, p] (q: ref patch, p: input patch), row-major patches.

    sim = <P_in(p), P_ref(q)> / ((||P_ref(q)|| + 1e-5) * (||P_in(p)|| + 1e-5))
    with the two divisions switched by is_norm / norm_input (ref_map_util.py:62-63, 78-84).
    """
    dtype = dtype or feat_input.dtype
    fi = feat_input.to(dtype)
    fr = feat_ref.to(dtype)
    c = fi.shape[0]
    k = c * patch_size * patch_size
    p_ref = sample_patches_oracle(fr, patch_size, ref_stride).reshape(k, -1)       # [K, N_ref]
    p_in = sample_patches_oracle(fi, patch_size, input_stride).reshape(k, -1)      # [K, N_in]
    if is_norm:
        p_ref = p_ref / (p_ref.norm(p=2, dim=0) + 1e-5)
    sim = p_ref.t() @ p_in                                                         # [N_ref, N_in]
    if norm_input:
        sim = sim / (p_in.norm(p=2, dim=0) + 1e-5)
    return sim


def feature_match_index_oracle(feat_input, feat_ref, patch_size=3, input_stride=1, ref_stride=1,
                               is_norm=True, norm_input=False, dtype=None, use_conv=False,
                               return_gap=False, chunk=None):
    """CPU restatement of feature_match_index (ref_map_util.py:26-86).

    Returns (max_idx int64 [h', w'], max_val [h', w']) and, if return_gap, the
    top-1 minus top-2 similarity per input position (same scaling as max_val)
    which the parity tests use for the "gap < 1e-5" exclusion the north star names.
    """
    dtype = dtype or feat_input.dtype
    c, h, w = feat_input.shape
    ho = (h - patch_size) // input_stride + 1
    wo = (w - patch_size) // input_stride + 1
    if chunk and not use_conv:
        # same quantity, evaluated for `chunk` input positions at a time so that the [N_ref, N_in] matrix of the
        # native validation shapes (125^2 / 128^2 feature grids: 2 GB in fp64) never exists in full
        k = c * patch_size * patch_size
        p_ref = sample_patches_oracle(feat_ref.to(dtype), patch_size, ref_stride).reshape(k, -1)
        p_in = sample_patches_oracle(feat_input.to(dtype), patch_size, input_stride).reshape(k, -1)
        if is_norm:
            p_ref = p_ref / (p_ref.norm(p=2, dim=0) + 1e-5)
        p_ref_t = p_ref.t().contiguous()
        idxs, vals, gaps = [], [], []
        for lo in range(0, p_in.shape[1], chunk):
            blk = p_in[:, lo:lo + chunk]
            sim = p_ref_t @ blk
            if norm_input:
                sim = sim / (blk.norm(p=2, dim=0) + 1e-5)
            v, i = sim.max(dim=0)
            idxs.append(i)
            vals.append(v)
            if return_gap:
                if sim.shape[0] > 1:
                    t2 = sim.topk(2, dim=0).values
                    gaps.append(t2[0] - t2[1])
                else:
                    gaps.append(torch.full_like(v, float('inf')))
        out = (torch.cat(idxs).view(ho, wo), torch.cat(vals).view(ho, wo))
        if return_gap:
            out = out + (torch.cat(gaps).view(ho, wo),)
        return out
    if use_conv:
        # literally the reference's evaluation order: conv2d with ref patches as filters
        fr = feat_ref.to(dtype)
        fi = feat_input.to(dtype)
        filt = sample_patches_oracle(fr, patch_size, ref_stride)                   # [C,ps,ps,N]
        if is_norm:
            filt = filt / (filt.norm(p=2, dim=(0, 1, 2)) + 1e-5)
        sim = F.conv2d(fi.unsqueeze(0), filt.permute(3, 0, 1, 2), stride=input_stride)[0]
        sim = sim.reshape(sim.shape[0], -1)
        if norm_input:
            p_in = sample_patches_oracle(fi, patch_size, input_stride)
            sim = sim / (p_in.norm(p=2, dim=(0, 1, 2)) + 1e-5)
    else:
        sim = similarity_volume(feat_input, feat_ref, patch_size, input_stride, ref_stride,
                                is_norm, norm_input, dtype)
    max_val, max_idx = sim.max(dim=0)            # first maximal index on ties
    out = (max_idx.view(ho, wo), max_val.view(ho, wo))
    if return_gap:
        if sim.shape[0] > 1:
            top2 = sim.topk(2, dim=0).values
            gap = (top2[0] - top2[1]).view(ho, wo)
        else:
            gap = torch.full((ho, wo), float('inf'), dtype=sim.dtype)
        out = out + (gap,)
    return out
