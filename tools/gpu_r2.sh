#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    name=r[hdr.index('Kernel Name')]; v=float(r[hdr.index('Metric Value')].replace(',','')); unit=r[hdr.index('Metric Unit')]
    if unit in('us','usecond'): v*=1000
    elif unit in('ms','msecond'): v*=1e6
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for n,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print(n[:90], c, int(t/c))
PY
