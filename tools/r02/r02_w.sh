#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02w}
timeout 600 python -m pytest tests/test_dcn_gpu.py -q -x -k "variants or fused_dynagg or golden" 2>&1 | tail -3
: > gpurun_out/${T}.jsonl
timeout 200 python tools/dcn_ab.py default >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
MREFSR_DCN_ALT=1 timeout 200 python tools/dcn_ab.py alt >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
python - <<PY
import json
for l in open('gpurun_out/${T}.jsonl'):
    j = json.loads(l)
    print(j.get('tag'), j.get('flow'), j.get('C', 'total'), j.get('ms', j.get('total_ms')), j.get('checksum', ''))
PY
tail -2 gpurun_out/${T}.err
