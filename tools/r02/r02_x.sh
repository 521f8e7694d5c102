#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02x}
: > gpurun_out/${T}.jsonl
export MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/windbg.so
for alt in 0 1; do
for kb in 150 175 205; do
  MREFSR_DCN_ALT=$alt MREFSR_DCN_SMEM_KB=$kb timeout 120 python tools/dcn_ablate.py alt$alt-$kb 64 128 256 >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
done
done
for dbg in 1 6 16 63; do
  MREFSR_DCN_ALT=1 MREFSR_DCN_DBG=$dbg timeout 120 python tools/dcn_ablate.py alt1-dbg$dbg 64 >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
done
python - <<PY
import sys, json, collections
d = collections.OrderedDict()
for l in open('gpurun_out/${T}.jsonl'):
    j = json.loads(l); d.setdefault(j['tag'], {})[j['C']] = j['ms']
for k, v in d.items(): print(k, v, round(sum(v.values()), 3))
PY
tail -2 gpurun_out/${T}.err
