"""Full-model throughput variants (cudnn.benchmark, channels_last) -- exploratory."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrefsr_b200.models import MRefSRPipeline
dev = 'cuda:0'
b, r = 16, 5
torch.manual_seed(10)
net = MRefSRPipeline().eval().to(dev)
g = torch.Generator().manual_seed(99)
lq = torch.rand(b, 3, 40, 40, generator=g).to(dev)
up = torch.nn.functional.interpolate(lq, scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)
refs = torch.rand(b, r, 3, 160, 160, generator=g).to(dev)
def run(tag):
    for _ in range(3): net(lq, up, refs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): net(lq, up, refs)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f'{tag}: {dt*1e3:.1f} ms/step  {b/dt:.0f} img/s', flush=True)
run('default')
torch.backends.cudnn.benchmark = True
run('cudnn.benchmark')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    net(lq, up, refs); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=70))
