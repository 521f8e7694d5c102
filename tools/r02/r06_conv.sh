#!/bin/bash
# r06c: channel-tile choice of the bf16 <-> fp32 conversion kernels: tests + training step
T=${1:-r06c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trunk_gpu.py tests/test_fusion_gpu.py tests/test_dcn_gpu.py -m gpu -q -x --tb=short -k "layout_convert or bf16 or fused_autograd or folded or training" 2>&1 | tail -3
timeout 300 python tools/train_step_bench.py --batch 12 --bf16 --channels-last 2>&1 | tail -1 | cut -c1-200
