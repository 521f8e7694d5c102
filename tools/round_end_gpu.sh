#!/bin/bash
# One gpurun call that refreshes everything measured for a round (TAG = file prefix under gpurun_out/):
#   GPU tests, reference-route comparison, bench.py (N=1), ncu launch list of the bench, ncu --set full of the DCN launches.
TAG=${1:-r04}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tests/perf_reference_gpu.py > gpurun_out/${TAG}_gpu_reference.jsonl 2> gpurun_out/${TAG}_gpu_reference.err
tail -1 gpurun_out/${TAG}_gpu_reference.jsonl
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 400 gpurun_out/${TAG}_bench.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
head -c 300 gpurun_out/${TAG}_bench_reference.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-full-model > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:dcn_tc_ -s 3 -c 3 -o gpurun_out/${TAG}_dcn_full -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-full-model > gpurun_out/${TAG}_ncu_dcn.log 2>&1
ls -la gpurun_out/${TAG}_* | head -20
ncu --set full --import-source on --clock-control none -k regex:match_diag -s 1 -c 1 -o gpurun_out/${TAG}_match_full -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-full-model --no-extras > gpurun_out/${TAG}_ncu_match.log 2>&1
ls -la gpurun_out/${TAG}_match_full* | head -3
