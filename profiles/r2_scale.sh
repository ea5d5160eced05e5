#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash profiles/r2_scale.sh N'
N=$1
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/nccl_check.py > gpurun_out/r2_nccl_check_${N}gpu.log 2>&1; echo "nccl_check rc=$?"
tail -5 gpurun_out/r2_nccl_check_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu > gpurun_out/r2_scale_c4_${N}gpu.json 2> gpurun_out/r2_scale_c4_${N}gpu.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_scale_c4_${N}gpu.json; tail -5 gpurun_out/r2_scale_c4_${N}gpu.err
