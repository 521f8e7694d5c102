#!/bin/bash
# window-kernel bring-up: DCN tests, A/B of the three flow kinds with and without the window kernel
mkdir -p gpurun_out
T=${1:-r02b}
timeout 900 python -m pytest tests/test_dcn_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/${T}_pytest_dcn.log; tail -5 gpurun_out/${T}_pytest_dcn.log
timeout 300 python tools/dcn_ab.py win > gpurun_out/${T}_dcn_ab.jsonl 2> gpurun_out/${T}_dcn_ab.err; tail -3 gpurun_out/${T}_dcn_ab.err
MREFSR_DCN_WIN=0 timeout 300 python tools/dcn_ab.py nowin >> gpurun_out/${T}_dcn_ab.jsonl 2>> gpurun_out/${T}_dcn_ab.err
cat gpurun_out/${T}_dcn_ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l)
    print(j.get('tag'), j.get('flow'), j.get('C', 'total'), j.get('ms', j.get('total_ms')), j.get('checksum', ''))
"
