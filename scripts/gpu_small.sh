#!/bin/bash
python scripts/gpu_tune.py 128 200 0,0,-1,0 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches128.csv python scripts/gpu_tune.py 128 3 0,0,-1,0 > /dev/null 2>&1
python - <<'PY'
import csv
from collections import defaultdict
d=defaultdict(list)
rr=[x for x in csv.reader(open('gpurun_out/launches128.csv')) if x and not x[0].startswith('==')]
h=rr[0]; ki,vi=h.index('Kernel Name'),h.index('Metric Value')
for x in rr[1:]:
    if len(x)>vi: d[x[ki]].append(float(x[vi].replace(',','')))
for k,v in d.items(): print(f"{k[:70]:70s} n={len(v):3d} mean={sum(v)/len(v)/1e3:8.1f} us")
PY
