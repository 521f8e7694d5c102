#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02o}
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err
echo "stdout lines: $(wc -l < gpurun_out/${T}_bench_n8.json)"; head -c 250 gpurun_out/${T}_bench_n8.json; echo
grep -c "NCCL INFO" gpurun_out/${T}_bench_n8.err; grep -m2 "nranks 8" gpurun_out/${T}_bench_n8.err | cut -c1-160
grep -i "error\|Traceback" gpurun_out/${T}_bench_n8.err | head -5
