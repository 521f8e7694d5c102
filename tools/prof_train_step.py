"""torch-profiler tables of one config-5 training step: by CUDA kernel and by aten operator (with input shapes).
    python tools/prof_train_step.py [--bf16] [--channels-last]"""
import sys, runpy, torch
sys.argv = ['tools/train_step_bench.py', '--batch', '12', '--steps', '2'] + sys.argv[1:]
ns = runpy.run_path('tools/train_step_bench.py', run_name='__main__')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as p:
    ns['step'](); torch.cuda.synchronize()
print(p.key_averages().table(sort_by='self_cuda_time_total', row_limit=40, max_name_column_width=70))
print(p.key_averages(group_by_input_shape=True).table(sort_by='self_cuda_time_total', row_limit=50, max_name_column_width=50,
                                                      max_shapes_column_width=90))
