mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c_pytest_gpu.log
timeout 300 python bench.py --workload c3 --disp skd --no-cpu > $O/r2c_bench_c3_skd.json 2> $O/r2c_bench_c3_skd.err; echo "c3 skd rc=$?"
NEM_B200_DENSITY_WARP=1 timeout 300 python bench.py --workload c3 --disp skd --no-cpu > $O/r2c_bench_c3_skd_warp.json 2> $O/r2c_bench_c3_skd_warp.err; echo "c3 skd warp rc=$?"
timeout 300 python bench.py --workload c4 --disp skd --no-cpu --steps 3 > $O/r2c_bench_c4_skd.json 2> $O/r2c_bench_c4_skd.err; echo "c4 skd rc=$?"
python - <<'P'
import json
for f in ['c3_skd','c3_skd_warp','c4_skd']:
    try:
        j=json.load(open(f'gpurun_out/r2c_bench_{f}.json')); r=j['roofline']
        print(f, 'value %.3g ms %.3f iters %s | density %.4f ms x%d frac %.3f'%(j['value'], j['ms_per_step'], j['config']['em_iterations_per_fit'], r['avg_launch_ms'], r['launches_timed'], r['frac']))
    except Exception as e: print(f, 'ERR', e)
P
