mkdir -p gpurun_out; O=gpurun_out
timeout 300 python profiles/c5_probe.py > $O/r2_c5_probe.txt 2> $O/r2_c5_probe.err; echo "probe rc=$?"; cat $O/r2_c5_probe.txt; tail -3 $O/r2_c5_probe.err
RUNS=4 WORKERS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_c5_launches.csv python profiles/c5_probe.py > $O/ncu_c5.log 2>&1; echo "ncu rc=$?"
python profiles/summarize.py 2>&1 | head -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_density_general_tiled -s 3 -c 1 -o /tmp/r2_dgt -f python bench.py --workload c3 --disp skd --no-cpu --steps 1 --warmup 1 > $O/ncu_dgt.log 2>&1; echo "ncu dgt rc=$?"
if [ -f /tmp/r2_dgt.ncu-rep ]; then
  python profiles/summarize.py full /tmp/r2_dgt.ncu-rep $O/r2_c3_k_density_general_tiled_full.csv
  python profiles/summarize.py stalls /tmp/r2_dgt.ncu-rep $O/r2_c3_k_density_general_tiled_stalls.txt
fi
ls -la $O | tail -8
