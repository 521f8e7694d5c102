#!/bin/bash
# gpurun helper: DCN tuning-knob / variant A/B (tools/dcn_ab.py) -> gpurun_out/dcn_ab.jsonl
# knobs: MREFSR_DCN_TILE=linear|2d, MREFSR_DCN_SPLIT=0|1; variants: tools/build_variant.sh NAME "-D..."
mkdir -p gpurun_out
: > gpurun_out/dcn_ab.jsonl
for lib in default $(ls mrefsr_b200/lib/variants/*.so 2>/dev/null); do
  for tile in ${TILES:-linear 2d}; do
    for split in ${SPLITS:-0 1}; do
      if [ "$lib" = default ]; then unset MREFSR_LIB; else export MREFSR_LIB=$PWD/$lib; fi
      MREFSR_DCN_TILE=$tile MREFSR_DCN_SPLIT=$split timeout 300 python tools/dcn_ab.py "$(basename $lib .so)/$tile/split$split" >> gpurun_out/dcn_ab.jsonl 2>> gpurun_out/dcn_ab.err
    done
  done
done
unset MREFSR_LIB
grep total_ms gpurun_out/dcn_ab.jsonl
