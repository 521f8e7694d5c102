#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02q}
: > gpurun_out/${T}_ablate_default.jsonl
export MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/windbg.so
for dbg in 0 1 2 4 8 16 33 6 39 63; do
  MREFSR_DCN_WIN=0 MREFSR_DCN_DBG=$dbg timeout 120 python tools/dcn_ablate.py default 64 >> gpurun_out/${T}_ablate_default.jsonl 2>> gpurun_out/${T}_ablate.err
done
cat gpurun_out/${T}_ablate_default.jsonl; tail -2 gpurun_out/${T}_ablate.err
