#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02e}
: > gpurun_out/${T}_ablate.jsonl
export MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/windbg.so
for dbg in 0 1 2 3 4 8 16 32 33 37 45 61 63; do
  MREFSR_DCN_DBG=$dbg timeout 120 python tools/dcn_ablate.py $T 64 256 >> gpurun_out/${T}_ablate.jsonl 2>> gpurun_out/${T}_ablate.err
done
cat gpurun_out/${T}_ablate.jsonl
tail -3 gpurun_out/${T}_ablate.err
