#!/bin/bash
# r06: GEMM epilogue stores staged through shared memory: GEMM / backward tests + training step + profile rows
T=${1:-r06a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dcn_gpu.py -m gpu -q -x --tb=short -k "gemm or backward or golden or reference_cuda or fused_autograd" 2>&1 | tail -3
timeout 300 python tools/train_step_bench.py --batch 12 --bf16 --channels-last 2>&1 | tail -1 | cut -c1-200
timeout 600 python tools/prof_train_step.py --bf16 --channels-last > gpurun_out/${T}_train_prof.txt 2>&1
grep -E "Self CUDA time total|gemm_tf32_nt|dcn_bwd_coord_cols" gpurun_out/${T}_train_prof.txt | cut -c1-60,120-200 | head -4
