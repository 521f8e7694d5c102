#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02s}
: > gpurun_out/${T}_smem.jsonl
for lib in windbg windbg_ntab2; do
export MREFSR_LIB=$PWD/mrefsr_b200/lib/variants/$lib.so
for kb in 150 175 205 222; do
  MREFSR_DCN_WIN=0 MREFSR_DCN_SMEM_KB=$kb timeout 120 python tools/dcn_ablate.py $lib-$kb 64 128 256 >> gpurun_out/${T}_smem.jsonl 2>> gpurun_out/${T}_smem.err
done
done
cat gpurun_out/${T}_smem.jsonl | python -c "
import sys, json, collections
d = collections.OrderedDict()
for l in sys.stdin:
    j = json.loads(l); d.setdefault(j['tag'], {})[j['C']] = j['ms']
for k, v in d.items(): print(k, v, round(sum(v.values()), 3))
"
tail -2 gpurun_out/${T}_smem.err
