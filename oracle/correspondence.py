"""Oracle for idx -> flow -> shifted pre-offsets (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates basicsr/archs/corres_generation_arch.py:30-47 (index_to_flow) and :70-105
(nine zero-filled shifts at three scales), with tensor_shift from
basicsr/archs/arch_util.py:386-410, in closed form:

  flow[y, x]  = (idx % w' - x, idx // w' - y)        for y < h-2, x < w-2   (w' = w-2)
              = (0, 0)                               on the 2-wide bottom/right border (:45)
  pre_s[k=3i+j, Y, X] = s * flow[Y//s - i, X//s - j]  if Y//s >= i and X//s >= j else 0

for s in {1, 2, 4}: repeat_interleave by s then shifting by (s*i, s*j) (:82-98) equals
shifting the coarse grid by (i, j) then repeating.  Last dim order is (x, y) (:41,100).
"""
import torch


def index_to_flow_oracle(max_idx):
    """max_idx int64 [h', w'] -> flow float32 [1, h'+2, w'+2, 2] (x, y), zero padded."""
    hp, wp = max_idx.shape
    fx = (max_idx % wp).to(torch.float32)
    fy = torch.div(max_idx, wp, rounding_mode='floor').to(torch.float32)
    gy, gx = torch.meshgrid(torch.arange(hp), torch.arange(wp), indexing='ij')
    flow = torch.stack((fx - gx.float(), fy - gy.float()), dim=2)
    out = torch.zeros(1, hp + 2, wp + 2, 2, dtype=torch.float32)
    out[0, :hp, :wp] = flow
    return out


def pre_offsets_oracle(max_idx):
    """max_idx int64 [h-2, w-2] -> dict of float32 [9, s*h, s*w, 2] for relu3_1/2_1/1_1."""
    flow = index_to_flow_oracle(max_idx)[0]                  # [h, w, 2]
    h, w, _ = flow.shape
    out = {}
    for name, s in (('relu3_1', 1), ('relu2_1', 2), ('relu1_1', 4)):
        res = torch.zeros(9, s * h, s * w, 2, dtype=torch.float32)
        big = flow.repeat_interleave(s, 0).repeat_interleave(s, 1) * s
        for i in range(3):
            for j in range(3):
                si, sj = s * i, s * j
                res[3 * i + j, si:, sj:] = big[:s * h - si, :s * w - sj]
        out[name] = res
    return out
