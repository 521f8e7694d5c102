#!/bin/bash
# r04: graph-replayed forward_ragged: tests + the ragged leg alone
T=${1:-r04p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -6
timeout 600 python - <<PY | tee gpurun_out/${T}_ragged.json
import json, sys, torch
sys.path.insert(0, 'tools')
import bench_legs as L
print(json.dumps(L.ragged_leg(torch.device('cuda', 0), 0, 1, None)))
PY
