#!/bin/bash
# r04a: matcher in diagonal form: tests + bench quick line
T=${1:-r04a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_matcher_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-full-model --no-extras > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
j = json.load(open('gpurun_out/${T}_bench.json'))
print(j['value'], j['ms_per_step'], json.dumps(j['kernel_ms_per_step']))
PY
tail -3 gpurun_out/${T}_bench.err
if [ -n "$NCU_MATCH" ]; then
ncu --set full --import-source on --clock-control none -k regex:match_diag -s 2 -c 1 -o gpurun_out/${T}_match_full -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-full-model --no-extras > gpurun_out/${T}_ncu_match.log 2>&1
ls -la gpurun_out/${T}_match_full*
fi
