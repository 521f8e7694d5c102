#!/bin/bash
# r04: DCN backward kernels: tests + training step + profile
T=${1:-r04l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dcn_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "backward or golden or fused_autograd or training or reference_cuda" --tb=short 2>&1 | tail -6
timeout 300 python tools/train_step_bench.py --batch 12 --bf16 --channels-last 2>&1 | tail -1 | tee gpurun_out/${T}_train.json
timeout 600 python tools/prof_train_step.py --bf16 --channels-last > gpurun_out/${T}_train_prof.txt 2>&1
grep -E "Self CUDA time total|dcn_bwd_coord|dcn_im2col_planes|gemm_tf32|DynAggDCNFunctionBackward" gpurun_out/${T}_train_prof.txt | cut -c1-60,120-200 | head -8
