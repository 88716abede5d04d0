out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py tests/test_gpu_gicp.py tests/test_scanmatcher.py tests/test_gpu_baseline_sizes.py -m gpu -q > $out/pytest_r2o.log 2>&1; tail -3 $out/pytest_r2o.log
timeout 300 python tools/diag_c4.py 8 > $out/diag_c4_r2o.log 2>&1; tail -4 $out/diag_c4_r2o.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_c4_r2o.csv python tools/profile_c4.py > $out/prof_c4_r2o.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r2o.csv python tools/profile_step.py headline 1 0 > $out/prof_step_r2o.log 2>&1
python - <<PY
import csv, collections
for f in ["$out/launches_c4_r2o.csv", "$out/launches_r2o.csv"]:
    rows=list(csv.reader(open(f)))
    for i,r in enumerate(rows):
        if 'Kernel Name' in r: hdr=r; start=i+1; break
    ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
    agg=collections.OrderedDict(); tot=0
    for r in rows[start:]:
        if len(r)<=vi: continue
        name=r[ki].split('(')[0].replace('void ','').replace('b200::','').replace('<unnamed>::','').replace('unnamed>::','')
        v=float(r[vi].replace(',',''))
        if r[ui]=='ns': v/=1000
        elif r[ui]=='ms': v*=1000
        a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
    print(f, 'total', round(tot,1))
    for k,(n,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]:
        print(f"  {k[:50]:50s} {n:4d} {v:10.1f} us  avg {v/n:8.1f}")
PY
timeout 300 python bench.py --workload c3 --no-cpu-baseline > $out/bench_c3_r2o.json 2> $out/bench_c3_r2o.err; python -c "
import json; l=json.loads(open('$out/bench_c3_r2o.json').read().strip().splitlines()[-1]); print('c3', l['value'], l['ms_per_step'])"
