#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02z}
: > gpurun_out/${T}.jsonl
timeout 200 python tools/dcn_ab.py default >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
export MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/windbg.so
for kb in 205 222; do
  MREFSR_DCN_SMEM_KB=$kb timeout 120 python tools/dcn_ablate.py smem$kb 256 128 >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
done
for dbg in 6 16 1; do
  MREFSR_DCN_DBG=$dbg timeout 120 python tools/dcn_ablate.py dbg$dbg 64 >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
done
python - <<PY
import json
for l in open('gpurun_out/${T}.jsonl'):
    j = json.loads(l)
    if j.get('flow') in ('coherent', 'random') and 'C' in j: continue
    print(j.get('tag'), j.get('flow', j.get('dbg')), j.get('C', 'total'), j.get('ms', j.get('total_ms')), j.get('checksum', ''))
PY
tail -2 gpurun_out/${T}.err
