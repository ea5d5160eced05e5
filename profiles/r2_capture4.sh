#!/bin/bash
# builder (gather / edges) + fuzzy M-step rewrite: GPU suite, C5 probe, C3 nem line
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r2d_pytest_gpu.log
WORKERS=1,4,8 timeout 300 python profiles/c5_probe.py > $O/r2d_c5_probe.txt 2> $O/r2d_c5_probe.err; echo "probe rc=$?"; cat $O/r2d_c5_probe.txt; tail -3 $O/r2d_c5_probe.err
timeout 300 python bench.py --workload c3 --algo nem --update para --no-cpu > $O/r2d_bench_c3_nem.json 2> $O/r2d_bench_c3_nem.err; echo "c3 nem rc=$?"
timeout 400 python bench.py --workload c5 --no-cpu > $O/r2d_bench_c5.json 2> $O/r2d_bench_c5.err; echo "c5 rc=$?"
python - <<'P'
import json
j=json.load(open('gpurun_out/r2d_bench_c3_nem.json')); r=j['roofline']
print('c3 nem value %.3g ms %.3f iters %s mstep %.4f ms'%(j['value'], j['ms_per_step'], j['config']['em_iterations_per_fit'], r['mstep']['avg_ms']))
j=json.load(open('gpurun_out/r2d_bench_c5.json')); print('c5 value %.3g ms/step %.1f launches %d'%(j['value'], j['ms_per_step'], j['gpu_launches']))
P
