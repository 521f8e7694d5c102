import sys, runpy, torch
sys.argv = ['tools/train_step_bench.py', '--batch', '12', '--steps', '2'] + sys.argv[1:]
ns = runpy.run_path('tools/train_step_bench.py', run_name='__main__')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    ns['step'](); torch.cuda.synchronize()
print(p.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=90))
