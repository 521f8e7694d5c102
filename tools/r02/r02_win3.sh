#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02i}
timeout 600 python -m pytest tests/test_dcn_gpu.py -q -x -k "window or variants or fused" 2>&1 | tail -4
timeout 300 python tools/dcn_ab.py win > gpurun_out/${T}_dcn_ab.jsonl 2> gpurun_out/${T}_dcn_ab.err; tail -3 gpurun_out/${T}_dcn_ab.err
python - <<PY
import json
for l in open('gpurun_out/${T}_dcn_ab.jsonl'):
    j = json.loads(l)
    print(j.get('tag'), j.get('flow'), j.get('C', 'total'), j.get('ms', j.get('total_ms')))
PY
