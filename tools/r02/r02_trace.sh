#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02g}
export MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/windbg.so
for dbg in 0 127; do
MREFSR_DCN_TRACE=1 MREFSR_DCN_DBG=$dbg timeout 120 python tools/dcn_ablate.py $T 64 > gpurun_out/${T}_trace_$dbg.log 2>&1
done
unset MREFSR_LIB
timeout 300 python tools/dcn_ab.py win > gpurun_out/${T}_dcn_ab.jsonl 2> gpurun_out/${T}_dcn_ab.err; grep total gpurun_out/${T}_dcn_ab.jsonl
