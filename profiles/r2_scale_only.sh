#!/bin/bash
# gpurun --gpus N -- 'bash profiles/r2_scale_only.sh N'   (the driver's command at N GPUs)
N=$1
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_c4_${N}gpu.json 2> gpurun_out/r2_scale_c4_${N}gpu.err; echo "bench rc=$?"
grep -a "NemError\|illegal" gpurun_out/r2_scale_c4_${N}gpu.err | grep -a -v "alloc.h" | head -5 | cut -c1-300
head -c 300 gpurun_out/r2_scale_c4_${N}gpu.json
