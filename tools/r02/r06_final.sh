#!/bin/bash
# r06: last verification pass of the round (no --set full captures: the hot-path kernels are those of profiles/r04_*)
TAG=${1:-r06}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest.log
tail -2 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
j = json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', round(j['value']), 'ms', round(j['ms_per_step'], 3), 'e2e', round(j['e2e']['value'], 1), 'full', round(j['full_model']['value']),
      'train', round(j['train_step']['ms_per_step'], 1), 'ragged', round(j['ragged']['ms_per_step'], 1),
      'refshard', j['refshard']['ms_per_image'], 'gpu_ref x', j['gpu_reference'].get('speedup'))
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
head -c 200 gpurun_out/${TAG}_bench_reference.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-full-model --no-extras > gpurun_out/${TAG}_ncu_bench.log 2>&1
ls gpurun_out/${TAG}_* | wc -l
