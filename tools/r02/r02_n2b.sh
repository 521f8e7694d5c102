#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02n}
timeout 600 python -m pytest tests/test_refshard_gpu.py tests/test_dcn_gpu.py -q -x -k "refshard or reference_sharded or slabs or several_buffers" 2>&1 | tail -4
cat > /tmp/leg.py <<'PY'
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import bench_legs
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
out = bench_legs.refshard_leg(dev, rank, world, dist)
if rank == 0: print(json.dumps(out))
dist.destroy_process_group()
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 /tmp/leg.py > gpurun_out/${T}_refshard_n2.json 2> gpurun_out/${T}_refshard_n2.err; cat gpurun_out/${T}_refshard_n2.json; tail -3 gpurun_out/${T}_refshard_n2.err | cut -c1-300
