#!/bin/bash
# full GPU check: test suite, bench with the extra legs, reference arm
mkdir -p gpurun_out
T=${1:-r02j}
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; head -c 300 gpurun_out/${T}_bench.json; echo; tail -3 gpurun_out/${T}_bench.err
