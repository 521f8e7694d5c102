"""One channels-last forward of the whole network for ncu (trunk glue kernels):
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:'bias_act|nhwc|maxpool|attn_modulate' \
    --csv --log-file gpurun_out/trunk_kernels.csv python tools/prof_trunk.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrefsr_b200.models import MRefSRPipeline
dev = 'cuda:0'
b, r = 16, 5
torch.manual_seed(10)
torch.backends.cudnn.benchmark = True
net = MRefSRPipeline().eval().to(dev).channels_last_()
g = torch.Generator().manual_seed(99)
lq = torch.rand(b, 3, 40, 40, generator=g).to(dev)
up = torch.nn.functional.interpolate(lq, scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)
refs = torch.rand(b, r, 3, 160, 160, generator=g).to(dev)
for _ in range(int(os.environ.get('WARM', 2))):
    net(lq, up, refs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
net(lq, up, refs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
