#!/bin/bash
# r04: training path (fused DynAgg node, batched references, fused bias / activation epilogues): tests + training step
T=${1:-r04j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trunk_gpu.py tests/test_dcn_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "training or fused_autograd or model or resblock" --tb=short 2>&1 | tail -12
timeout 300 python tools/train_step_bench.py --batch 12 --bf16 --channels-last 2>&1 | tail -1 | tee gpurun_out/${T}_train.json
timeout 600 python tools/prof_train_step.py --bf16 --channels-last > gpurun_out/${T}_train_prof.txt 2>&1
grep -E "Self CUDA time total|Self CPU time total" gpurun_out/${T}_train_prof.txt | head -2
