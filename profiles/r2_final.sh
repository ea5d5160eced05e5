#!/bin/bash
# final single-GPU pass of round 2: GPU suite, smoke, the default bench line, C5 and the skd line
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2f_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2f_smoke.log
timeout 600 python bench.py > $O/r2f_bench_c4.json 2> $O/r2f_bench_c4.err; echo "bench rc=$?"
timeout 300 python bench.py --workload c5 --no-cpu > $O/r2f_bench_c5.json 2> $O/r2f_bench_c5.err; echo "c5 rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_bench_ref.json 2> $O/r2f_bench_ref.err; echo "ref rc=$?"
python - <<'P'
import json
j=json.load(open('gpurun_out/r2f_bench_c4.json')); print('c4 value %.4g ms %.3f e2e %.4g frac %.3f launches %d'%(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['gpu_launches']))
j=json.load(open('gpurun_out/r2f_bench_c5.json')); print('c5 value %.4g'%j['value'])
j=json.load(open('gpurun_out/r2f_bench_ref.json')); print('ref', j.get('value'), j.get('unavailable'))
P
