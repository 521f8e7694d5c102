"""Tiny driver for ncu: one DCN forward per scale (80 samples), tcgen05 path."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrefsr_b200 import dcn as D
DEV = 'cuda:0'
which = sys.argv[1:] or ['256', '128', '64']
for c, hw in [(256, 40), (128, 80), (64, 160)]:
    if str(c) not in which:
        continue
    n = 80
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, c, hw, hw, generator=g).to(DEV)
    off = (torch.randn(n, 144, hw, hw, generator=g) * 3).to(DEV)
    mask = torch.rand(n, 72, hw, hw, generator=g).to(DEV)
    wgt = (torch.randn(c, c, 3, 3, generator=g) * 0.02).to(DEV)
    bias = torch.zeros(c).to(DEV)
    for _ in range(3):
        D.dcn_forward_raw(x, off, mask, wgt, bias, (1, 1), (1, 1), (1, 1), 1, 8, mode='tf32')
    torch.cuda.synchronize()
