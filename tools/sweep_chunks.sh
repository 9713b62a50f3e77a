#!/bin/bash
# step time of the default bench against the number of pipelined sub-batches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ch in 1 2 3 4 6 8; do
  python bench.py --steps 20 --warmup 3 --skip-cpu --chunks $ch > gpurun_out/sweep.json 2> gpurun_out/sweep.err || { echo "chunks $ch FAILED"; continue; }
  python - $ch <<'PY'
import json, sys
d = json.load(open('gpurun_out/sweep.json'))
print('chunks', sys.argv[1], round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 3))
PY
done
