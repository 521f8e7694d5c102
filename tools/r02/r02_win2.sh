#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02f}
timeout 900 python -m pytest tests/test_dcn_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/${T}_pytest_dcn.log; tail -5 gpurun_out/${T}_pytest_dcn.log
timeout 300 python tools/dcn_ab.py win > gpurun_out/${T}_dcn_ab.jsonl 2> gpurun_out/${T}_dcn_ab.err; tail -3 gpurun_out/${T}_dcn_ab.err
python - <<PY
import json
for l in open('gpurun_out/${T}_dcn_ab.jsonl'):
    j = json.loads(l)
    print(j.get('tag'), j.get('flow'), j.get('C', 'total'), j.get('ms', j.get('total_ms')), j.get('checksum', ''))
PY
: > gpurun_out/${T}_ablate.jsonl
export MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/windbg.so
for dbg in 0 1 2 4 32 33 37 127; do
  MREFSR_DCN_DBG=$dbg timeout 120 python tools/dcn_ablate.py $T 64 256 >> gpurun_out/${T}_ablate.jsonl 2>> gpurun_out/${T}_ablate.err
done
cat gpurun_out/${T}_ablate.jsonl
