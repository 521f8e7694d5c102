#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02p}
timeout 900 python -m pytest tests/test_dcn_gpu.py tests/test_refshard_gpu.py -q -x 2>&1 | tail -3
for f in "" "--no-overlap"; do
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-full-model --no-extras --steps 20 $f > gpurun_out/${T}_bench$f.json 2>> gpurun_out/${T}_bench.err
python - <<PY
import json
j = json.loads(open('gpurun_out/${T}_bench$f.json').read().strip().split('\n')[-1])
print('$f', round(j['value'], 1), round(j['ms_per_step'], 4), {k: round(v['ms_per_step'], 3) for k, v in j['kernel_ms_per_step'].items()})
PY
done
tail -2 gpurun_out/${T}_bench.err
