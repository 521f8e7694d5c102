#!/bin/bash
# ncu --set full of the DCN launch at the large scale (bench inputs, fused mode, coherent flows)
mkdir -p gpurun_out
T=${1:-r02c}
cat > /tmp/prof_win.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import bench, mrefsr_b200 as M
from mrefsr_b200.dcn import dynagg_dcn_forward
d = bench.make_inputs(16, 5, 1234, 'cuda:0')
ys, xs = torch.meshgrid(torch.arange(38), torch.arange(38), indexing='ij')
mi = torch.stack([(ys + 2).clamp(0, 37) * 38 + (xs - 3).clamp(0, 37)] * 80).cuda()
for c, hw in ((64, 160),):
    for _ in range(2):
        dynagg_dcn_forward(d[f'x{c}'], d[f'conv_out{c}'], mi, hw // 40, d[f'w{c}'], d[f'b{c}'], 8)
torch.cuda.synchronize()
PY
ncu --set full --import-source on --clock-control none -k regex:dcn_win_kernel -s 1 -c 1 -o gpurun_out/${T}_win_full -f python /tmp/prof_win.py > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
ls -la gpurun_out/${T}_win_full.ncu-rep
