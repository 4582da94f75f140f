#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dsd_launches.csv python tools/probe_dsd.py 2>&1 | tail -1
python - <<'PY'
import csv,collections
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/dsd_launches.csv") if not l.startswith("=="))]
d=collections.OrderedDict()
for r in rows:
    if r["Metric Name"]=="gpu__time_duration.sum":
        k=r["Kernel Name"][:70]; v=float(r["Metric Value"].replace(",",""))
        v*= {"ns":1e-6,"us":1e-3,"ms":1}.get(r["Metric Unit"],1e-6)
        d.setdefault(k,[]).append(v)
for k,v in d.items(): print("%-72s n=%d last=%.3f ms"%(k,len(v),v[-1]))
PY
