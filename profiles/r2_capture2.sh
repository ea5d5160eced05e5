#!/bin/bash
# Round 2, second evidence pass on the shipped binary:
#   gpurun --timeout 1500 -- 'bash profiles/r2_capture2.sh'
# GPU suite, default bench line, the off-default lines (skd / fuzzy nem on C3), C5 on one GPU, and
# compute-sanitizer memcheck + synccheck of the smoke fit and of one sharded (local-comm) fit.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }

timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2b_pytest_gpu.log 2>&1; stamp "pytest rc=$?"
tail -3 $O/r2b_pytest_gpu.log
timeout 600 python bench.py > $O/r2b_bench_c4.json 2> $O/r2b_bench_c4.err; stamp "bench c4 rc=$?"
timeout 300 python bench.py --workload c3 --disp skd --no-cpu > $O/r2b_bench_c3_skd.json 2> $O/r2b_bench_c3_skd.err; stamp "c3 skd rc=$?"
timeout 300 python bench.py --workload c3 --algo nem --update para --no-cpu > $O/r2b_bench_c3_nem.json 2> $O/r2b_bench_c3_nem.err; stamp "c3 nem rc=$?"
timeout 400 python bench.py --workload c5 --no-cpu > $O/r2b_bench_c5.json 2> $O/r2b_bench_c5.err; stamp "c5 rc=$?"
timeout 300 compute-sanitizer --tool memcheck --log-file $O/r2b_sanitizer_memcheck_smoke.log python __graft_entry__.py smoke > $O/r2b_san_mem.out 2>&1; stamp "memcheck smoke rc=$?"
timeout 300 compute-sanitizer --tool synccheck --log-file $O/r2b_sanitizer_synccheck_smoke.log python __graft_entry__.py smoke > $O/r2b_san_sync.out 2>&1; stamp "synccheck smoke rc=$?"
timeout 400 compute-sanitizer --tool memcheck --log-file $O/r2b_sanitizer_memcheck_sharded.log python -m pytest tests/test_gpu_sharded.py -x -q -k "needs_cross_rank_rounds" > $O/r2b_san_mem_sh.out 2>&1; stamp "memcheck sharded rc=$?"
timeout 400 compute-sanitizer --tool synccheck --log-file $O/r2b_sanitizer_synccheck_sharded.log python -m pytest tests/test_gpu_sharded.py -x -q -k "needs_cross_rank_rounds" > $O/r2b_san_sync_sh.out 2>&1; stamp "synccheck sharded rc=$?"
tail -2 $O/r2b_san_*.out
ls -la $O
