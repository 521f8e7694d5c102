#!/bin/bash
# round 2, first GPU call: sanity of the restored tree + pending round-1 hardware runs
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02a_pytest.log; tail -3 gpurun_out/r02a_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; head -c 600 gpurun_out/r02a_bench.json; echo
MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/fp16split.so python -m pytest tests/test_matcher_gpu.py -q 2>&1 | tail -5 > gpurun_out/r02a_fp16split_pytest.log; tail -2 gpurun_out/r02a_fp16split_pytest.log
MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/fp16split.so python bench.py --no-cpu-baseline --no-e2e --no-full-model > gpurun_out/r02a_bench_fp16split.json 2>> gpurun_out/r02a_bench.err; head -c 300 gpurun_out/r02a_bench_fp16split.json; echo
timeout 600 python tools/validate.py --synthetic 3 > gpurun_out/r02a_validate.log 2>&1; tail -6 gpurun_out/r02a_validate.log
timeout 300 python tools/dcn_ab.py base > gpurun_out/r02a_dcn_ab.jsonl 2> gpurun_out/r02a_dcn_ab.err; grep total_ms gpurun_out/r02a_dcn_ab.jsonl
