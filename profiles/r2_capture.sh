#!/bin/bash
# Round 2 evidence for the binary that ships (persistent EM kernel, default build):
#   gpurun --timeout 1500 -- 'bash profiles/r2_capture.sh'
# Every ncu pass runs only after the same command exited 0 without ncu; .ncu-rep files are turned
# into CSV/text pages on the box (gpurun_out/ is capped at 64 MiB) and removed.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }

timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1; stamp "pytest rc=$?"
tail -3 $O/r2_pytest_gpu.log
timeout 600 python bench.py > $O/r2_bench_c4.json 2> $O/r2_bench_c4.err; stamp "bench c4 rc=$?"
timeout 300 python profiles/pk_phase_probe.py > $O/r2_pk_phases.txt 2> $O/r2_pk_phases.err; stamp "phase probe rc=$?"

# launch list (share of the step per kernel)
timeout 200 python bench.py --no-cpu --no-extras --steps 2 --warmup 1 > $O/r2_plain.json 2> $O/r2_plain.err; stamp "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c4.csv \
    python bench.py --no-cpu --no-extras --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1; stamp "ncu launches rc=$?"

# full captures of the three kernels a C4 fit spends its time in
for k in k_em_persist k_density_tma k_mstep_ncem; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -o /tmp/r2_$k -f \
    python bench.py --no-cpu --no-extras --steps 1 --warmup 1 > $O/ncu_$k.log 2>&1; stamp "ncu $k rc=$?"
  if [ -f /tmp/r2_$k.ncu-rep ]; then
    python profiles/summarize.py full /tmp/r2_$k.ncu-rep $O/r2_c4_${k}_full.csv
    python profiles/summarize.py stalls /tmp/r2_$k.ncu-rep $O/r2_c4_${k}_stalls.txt
    ncu -i /tmp/r2_$k.ncu-rep --page raw --csv > $O/r2_c4_${k}_raw.csv 2>/dev/null
  fi
done
ls -la $O
