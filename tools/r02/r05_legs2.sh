#!/bin/bash
# r05: which earlier part of bench.py slows the train_step leg down
T=${1:-r05c}
mkdir -p gpurun_out
for flags in "--no-e2e --no-cpu-baseline --no-full-model" "--no-cpu-baseline --no-full-model" "--no-e2e --no-cpu-baseline"; do
  timeout 600 python bench.py $flags > gpurun_out/${T}_b.json 2> gpurun_out/${T}_b.err
  python - <<PY
import json
j = json.load(open('gpurun_out/${T}_b.json'))
print("$flags", '| train', round(j['train_step']['ms_per_step'], 1), '| ragged', round(j['ragged']['ms_per_step'], 1), '| value', round(j['value']))
PY
done
