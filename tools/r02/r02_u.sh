#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02u}
timeout 600 python -m pytest tests/test_fusion_gpu.py -q -x 2>&1 | tail -3
python - <<'PY' > gpurun_out/r02u_fusion_bf16.json
import json, torch, sys, os
sys.path.insert(0, os.getcwd())
import bench, mrefsr_b200 as M
d = bench.make_inputs(16, 5, 1234, 'cuda:0')
res = {}
for dt in (torch.float32, torch.bfloat16):
    tot = 0.0
    for c, hw in bench.SCALES:
        a, b_, v = (d[f'{k}{c}'].to(dt) for k in ('emb_t', 'emb', 'ass'))
        for _ in range(3): M.mrapa_attention(a, b_, v, 5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): M.mrapa_attention(a, b_, v, 5)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res['%s_C%d' % (str(dt).split('.')[-1], c)] = round(ms, 4)
        tot += ms
    res['%s_total_ms' % str(dt).split('.')[-1]] = round(tot, 4)
print(json.dumps(res))
PY
cat gpurun_out/r02u_fusion_bf16.json
