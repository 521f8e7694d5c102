"""Oracle for the multi-reference attention core (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates basicsr/archs/ref_mrapa_restoration_arch.py:321-335.  With
  q = conv_emb1(target) * C**-0.5          [n, C, h, w]      (:321, already scaled)
  k = conv_emb2(refs)                      [n*t, C, h, w]    (:324)
  v = conv_ass(refs)                       [n*t, 2C, h, w]   (:327)
the reference permutes to per-pixel matrices and computes
  p[n, y, x, :] = softmax_t( sum_c q[n, c, y, x] * k[n, t, c, y, x] )            (:331-332)
  out[n, :, y, x] = sum_t p[n, y, x, t] * v[n, t, :, y, x]                       (:333-335)
"""
import torch


def mrapa_attention_oracle(emb_t, emb, ass, t, dtype=None, return_prob=False):
    dtype = dtype or emb_t.dtype
    n, c, h, w = emb_t.shape
    q = emb_t.to(dtype)
    k = emb.to(dtype).view(n, t, c, h, w)
    v = ass.to(dtype).view(n, t, -1, h, w)
    logits = torch.einsum('nchw,ntchw->nthw', q, k)
    prob = torch.softmax(logits, dim=1)
    out = torch.einsum('nthw,ntchw->nchw', prob, v)
    return (out, prob) if return_prob else out
