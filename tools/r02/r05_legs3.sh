#!/bin/bash
T=${1:-r05d}
mkdir -p gpurun_out
for i in 1 2; do
  SECONDS=0; timeout 900 python bench.py > gpurun_out/${T}_b$i.json 2> gpurun_out/${T}_b$i.err
  echo "wall ${SECONDS}s"
  python - <<PY
import json
j = json.load(open('gpurun_out/${T}_b$i.json'))
print('run $i | train', round(j['train_step']['ms_per_step'], 1), '| ragged', round(j['ragged']['ms_per_step'], 1), '| value', round(j['value']), '| e2e', round(j['e2e']['value'], 1), '| full', round(j['full_model']['value']))
PY
done
