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


def refill_parameters(module, seed=0):
    """Deterministic, init-order-independent weights: every state-dict entry (except the fixed mean/std buffers) is
    re-drawn from a generator seeded by its KEY, so a reference module and its mirror with equal key names get equal
    weights.  Used by tests/golden/make_golden.py (reference side) and the full-model parity test (our side)."""
    import zlib
    sd = module.state_dict()
    with torch.no_grad():
        for key in sorted(sd):
            t = sd[key]
            if key.endswith('mean') or key.endswith('std') or not t.dtype.is_floating_point:
                continue
            g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + seed) & 0x7fffffff)
            if t.dim() == 4:
                fan_in = t.shape[1] * t.shape[2] * t.shape[3]
                scale = (0.25 if 'conv_offset_mask' in key else 0.7) / fan_in ** 0.5
                t.copy_(torch.randn(t.shape, generator=g) * scale)
            elif t.numel() == 1:
                t.fill_(0.25)                       # PReLU slope
            else:
                t.copy_(torch.randn(t.shape, generator=g) * 0.02)
    return module
