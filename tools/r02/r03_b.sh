#!/bin/bash
mkdir -p gpurun_out
T=${1:-r03b}
: > gpurun_out/${T}.jsonl
for lib in default spin20 spin50 spin100; do
if [ $lib = default ]; then unset MREFSR_LIB; else export MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/$lib.so; fi
timeout 200 python tools/dcn_ab.py $lib >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
done
python - <<PY
import json
for l in open('gpurun_out/${T}.jsonl'):
    j = json.loads(l)
    if j.get('flow') != 'bench': continue
    print(j.get('tag'), j.get('flow'), j.get('C', 'total'), j.get('ms', j.get('total_ms')), j.get('checksum', ''))
PY
tail -2 gpurun_out/${T}.err
