"""Full-model throughput variants of the plain-convolution trunk (memory format, cuDNN autotune, bf16 autocast)
with the SR-output deviation of each variant from the fp32/NCHW default -- exploratory, feeds models.py."""
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
ref_out = [None]


def run(tag, fn=None, prof=False):
    fn = fn or (lambda: net(lq, up, refs))
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        out = fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    out = out.float()
    if ref_out[0] is None:
        ref_out[0] = out
    err = float((out - ref_out[0]).abs().max() / ref_out[0].abs().max())
    print(f'{tag}: {dt*1e3:.1f} ms/step  {b/dt:.0f} img/s   rel dev from default {err:.2e}', flush=True)
    if prof:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as p:
            fn(); torch.cuda.synchronize()
        print(p.key_averages().table(sort_by='cuda_time_total', row_limit=22, max_name_column_width=80))


which = sys.argv[1:] or ['default', 'notf32', 'bench', 'cl', 'bf16']
if 'default' in which:
    run('default (NCHW, cuDNN TF32)', prof='prof' in which)
if 'notf32' in which:
    torch.backends.cudnn.allow_tf32 = False
    run('NCHW, cuDNN fp32 (allow_tf32 off)')
    torch.backends.cudnn.allow_tf32 = True
if 'bench' in which:
    torch.backends.cudnn.benchmark = True
    run('NCHW + cudnn.benchmark')
if 'cl' in which:
    torch.backends.cudnn.benchmark = True
    net.channels_last_()
    run('channels_last + cudnn.benchmark', prof='prof' in which)
if 'bf16' in which:
    def f():
        with torch.autocast('cuda', dtype=torch.bfloat16):
            return net(lq, up, refs)
    run('bf16 autocast (plain convs), current layout', f)
if 'graph' in which:
    # CUDA-graph replay of the whole forward (static input buffers)
    static = [t.clone() for t in (lq, up, refs)]
    s_ = torch.cuda.Stream()
    s_.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_):
        for _ in range(3):
            net(*static)
    torch.cuda.current_stream().wait_stream(s_)
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        out_static = net(*static)

    def replay():
        g_.replay()
        return out_static
    run('CUDA graph replay of the current configuration', replay)
if 'fusedconv' in which:
    # micro: conv + bias + ReLU as F.conv2d + bias_act_ vs torch.cudnn_convolution_relu (cuDNN fused epilogue)
    from mrefsr_b200 import trunk as T
    for cl in (False, True):
        for (n, c, hw) in ((16, 64, 160), (96, 64, 160), (16, 64, 40)):
            x = torch.randn(n, c, hw, hw, device=dev)
            conv = torch.nn.Conv2d(c, c, 3, 1, 1).to(dev)
            if cl:
                x = x.contiguous(memory_format=torch.channels_last); conv = conv.to(memory_format=torch.channels_last)
            def a():
                with torch.no_grad():
                    return T.conv_bias_act(x, conv, T.ACT_LEAKY, 0.0)
            def b_():
                with torch.no_grad():
                    return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, (1, 1), (1, 1), (1, 1), 1)
            for tag, f in (('conv2d + bias_act_', a), ('cudnn_convolution_relu', b_)):
                for _ in range(5): y = f()
                torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20): y = f()
                e1.record(); torch.cuda.synchronize()
                print(f'cl={cl} [{n},{c},{hw},{hw}] {tag}: {e0.elapsed_time(e1)/20*1e3:.0f} us', flush=True)
