#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_periodic.py tests/test_gpu_api.py -x -q 2>&1 | tail -3
python tools/periodic_bench.py 2048 3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_periodic.csv python tools/periodic_bench.py 512 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_launches_periodic.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
k = rows[hdr].index("Kernel Name"); v = rows[hdr].index("Metric Value")
c = collections.Counter(); t = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) > v:
        name = r[k].split("(")[0]; c[name] += 1; t[name] += float(r[v].replace(",", ""))
for n, cnt in c.most_common(): print(f"{cnt:4d} x {n[:70]:70s} {t[n]/1e6:9.3f} ms")
PY
