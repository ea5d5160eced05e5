#!/bin/bash
# Round 2, first GPU call: evidence for the binary that ships (JAC_HUB_GUARD=1 default build).
#   profiles/ab/build_variant.sh guard0 -DJAC_HUB_GUARD=0; profiles/ab/build_variant.sh guard1
#   profiles/ab/build_variant.sh acq1 -DFIXUP_ACQ=1
#   gpurun --timeout 900 -- 'bash profiles/r2_capture_baseline.sh'
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python profiles/ab/hub_race_probe.py --seeds 3 > $O/r2_hub_probe_guard1.log 2>&1; echo "probe guard1 rc=$?" )
cp pangenomenem_b200/libnem_b200.so /tmp/tree.so
cp scratch/variants/libnem_b200.guard0.so pangenomenem_b200/libnem_b200.so
( timeout 300 python profiles/ab/hub_race_probe.py --seeds 3 > $O/r2_hub_probe_guard0.log 2>&1; echo "probe guard0 rc=$?" )
cp /tmp/tree.so pangenomenem_b200/libnem_b200.so
bash profiles/ab/run_ab.sh guard0:guard0 guard1:guard1 acq1:acq1 guard0:guard0_b guard1:guard1_b acq1:acq1_b > $O/r2_ab.log 2>&1
cat $O/r2_ab.log
for w in c1 c2 c3; do timeout 200 python bench.py --workload $w --no-cpu --steps 20 --warmup 5 > $O/r2_base_bench_$w.json 2> $O/r2_base_bench_$w.err; echo "$w rc=$?"; done
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2_base_bench_c4.json 2> $O/r2_base_bench_c4.err; echo "c4 rc=$?"
# launch list (share of the step per kernel), default build
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_base_launches_c4.csv \
    python bench.py --no-cpu --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# full captures: the fix-up tail, the grid-wide fix-up round, the dense round, the X^T recount, criteria
for k in k_sweep_ncem_fixup k_sweep_ncem_jacobi k_mstep_ncem k_criteria_partial k_mstep_finalize_tables; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 4 -o $O/r2_base_$k -f \
    python bench.py --no-cpu --steps 1 --warmup 1 > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la $O
