#!/bin/bash
# r04: backward-side kernels: tests + training step + profile
T=${1:-r04l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dcn_gpu.py tests/test_model_gpu.py tests/test_fusion_gpu.py tests/test_trunk_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.log
timeout 300 python tools/train_step_bench.py --batch 12 --bf16 --channels-last 2>&1 | tail -1 | tee gpurun_out/${T}_train.json
timeout 600 python tools/prof_train_step.py --bf16 --channels-last > gpurun_out/${T}_train_prof.txt 2>&1
grep -E "Self CUDA time total|mrapa_bwd|dynagg_offsets_vec4|bias_grad_sum" gpurun_out/${T}_train_prof.txt | cut -c1-60,120-200 | head -8
