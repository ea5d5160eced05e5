#!/bin/bash
# gpurun --gpus N -- 'bash profiles/r2_scale_debug.sh N [extra env]'
N=$1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu --no-extras --steps 3 > gpurun_out/r2_dbg_c4_${N}gpu.json 2> gpurun_out/r2_dbg_c4_${N}gpu.err; echo "bench rc=$?"
grep -a "nem_b200\|NemError" gpurun_out/r2_dbg_c4_${N}gpu.err | cut -c1-400 | head -40
head -c 400 gpurun_out/r2_dbg_c4_${N}gpu.json
