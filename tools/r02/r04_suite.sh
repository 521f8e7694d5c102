#!/bin/bash
# r04: whole GPU suite + smoke
T=${1:-r04f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
