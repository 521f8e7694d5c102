"""Shared helpers for the parity tests."""
import torch
import torch.nn.functional as F

import oracle


def unit_features(n, c, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, h * w, generator=g)
    return F.normalize(x, dim=1).view(n, c, h, w)


def match_parity(idx, val, fi, fr, kw, val_rtol=1e-3, gap_tol=1e-5, dtype=torch.float64):
    """Compare one pair's CUDA result with the oracle.  Returns dict(n_bad, n_lowgap, val_err).

    North-star contract: indices must match exactly except where the top-2 similarity gap is below 1e-5
    (those are counted); max similarity within 1e-3 relative."""
    o_idx, o_val, gap = oracle.feature_match_index_oracle(fi, fr, return_gap=True, dtype=dtype, **kw)
    idx, val = idx.cpu(), val.cpu().to(dtype)
    scale = max(1.0, float(o_val.abs().max()))
    diff = idx != o_idx
    lowgap = gap < gap_tol * scale
    n_bad = int((diff & ~lowgap).sum())
    n_low = int((diff & lowgap).sum())
    val_err = float(((val - o_val).abs() / (o_val.abs() + 1e-3 * scale)).max())
    return dict(n_bad=n_bad, n_lowgap_mismatch=n_low, n_lowgap=int(lowgap.sum()), val_err=val_err, n=idx.numel())


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
