#!/bin/bash
# ncu launch list of the default C4 command, the engine's kernels only (torch's data generation is
# ~530 launches before the first of them); the gpurun ncu wrapper runs the command plain first
mkdir -p gpurun_out; O=gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file $O/r2_c4_launches.csv python bench.py --no-cpu --no-extras --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1; echo "ncu rc=$?"
wc -l $O/r2_c4_launches.csv
