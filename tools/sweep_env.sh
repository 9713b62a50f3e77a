#!/bin/bash
# sweep PWB_OPT_CTAS_PER_SM x --chunks on the quick bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ch in 1 2 4 8; do for k in 3 4 5 6 8; do
  PWB_OPT_CTAS_PER_SM=$k python bench.py --steps 10 --warmup 3 --skip-cpu --chunks $ch > gpurun_out/sweep.json 2> gpurun_out/sweep.err || { echo "chunks $ch k $k FAILED"; continue; }
  python - $ch $k <<'PY'
import json, sys
d = json.load(open('gpurun_out/sweep.json'))
print('chunks', sys.argv[1], 'opt_ctas/sm', sys.argv[2], round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 3))
PY
done; done
