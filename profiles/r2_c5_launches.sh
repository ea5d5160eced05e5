mkdir -p gpurun_out; O=gpurun_out
RUNS=8 WORKERS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_c5_launches.csv python profiles/c5_probe.py > $O/ncu_c5.log 2>&1; echo "ncu rc=$?"
tail -3 $O/ncu_c5.log
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_c5_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print('%-62s %5d %10.1f us total %8.2f us avg'%(k,c,t/1e3,t/1e3/c))
P
