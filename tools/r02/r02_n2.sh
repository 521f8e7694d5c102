#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02k}
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
echo "stdout lines: $(wc -l < gpurun_out/${T}_bench_n2.json)"; head -c 200 gpurun_out/${T}_bench_n2.json; echo
grep -c "NCCL INFO" gpurun_out/${T}_bench_n2.err; grep -m3 "nranks\|Connected all rings\|NVLS" gpurun_out/${T}_bench_n2.err | cut -c1-200
tail -3 gpurun_out/${T}_bench_n2.err | cut -c1-300
