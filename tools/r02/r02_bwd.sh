#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02l}
timeout 900 python -m pytest tests/test_dcn_gpu.py -q -x -k "gemm or backward or golden" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_fusion_gpu.py -q -x 2>&1 | tail -3
timeout 300 python tools/train_step_bench.py --batch 12 --bf16 --channels-last 2>&1 | tail -2
timeout 300 python tools/train_step_bench.py --batch 12 --bf16 2>&1 | tail -1
