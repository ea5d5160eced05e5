#!/bin/bash
# On the GPU box (one gpurun call): the GPU test suite on the in-tree library, then the default
# workload once per variant, all on the same box so that the fits compare.
#   gpurun -- 'bash profiles/ab/run_ab.sh base:base new:new base:base_again'
# Each argument is VARIANT:TAG; prints  tag  fit_ms  sweep_ms  mstep_delta_ms  e2e_fit_ms  iters  repeatable
mkdir -p gpurun_out
cp pangenomenem_b200/libnem_b200.so /tmp/libnem_b200.tree.so
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -c 120 gpurun_out/pytest_gpu.log
run() { v=$1; tag=$2; cp scratch/variants/libnem_b200.$v.so pangenomenem_b200/libnem_b200.so
  timeout 100 python bench.py --no-cpu --steps 10 --warmup 3 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python -c "
import json;d=json.load(open('gpurun_out/ab_$tag.json'));r=d['roofline'];print('$tag',round(d['ms_per_step'],4),round(r['sweep_avg_ms'],4),round(r['mstep_delta']['avg_ms'],4),round(d['e2e']['fit_ms'],3),d['config']['em_iterations_per_fit'],d['full_size_properties']['labels_repeatable_across_fits'])"; }
for v in "$@"; do run ${v%%:*} ${v##*:}; done
cp /tmp/libnem_b200.tree.so pangenomenem_b200/libnem_b200.so
