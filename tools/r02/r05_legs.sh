#!/bin/bash
# r05: why the train_step leg is slower inside bench.py than alone: the leg alone, then after the ragged leg
T=${1:-r05b}
mkdir -p gpurun_out
timeout 900 python - <<PY | tee gpurun_out/${T}_legs.jsonl
import json, sys, torch
sys.path.insert(0, 'tools')
import bench_legs as L
dev = torch.device('cuda', 0)
r = L.train_step_leg(dev, 0, 1, None); print(json.dumps({'alone': r['ms_per_step']}), flush=True)
r = L.train_step_leg(dev, 0, 1, None, steps=6); print(json.dumps({'alone_6_steps': r['ms_per_step']}), flush=True)
torch.cuda.empty_cache()
r = L.ragged_leg(dev, 0, 1, None); print(json.dumps({'ragged': r['ms_per_step']}), flush=True)
torch.cuda.empty_cache()
r = L.train_step_leg(dev, 0, 1, None); print(json.dumps({'after_ragged': r['ms_per_step']}), flush=True)
PY
