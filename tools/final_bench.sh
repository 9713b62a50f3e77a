#!/bin/bash
# Round-end evidence, part B: the un-profiled default bench (with the CPU baseline leg) and the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_final.json'))
print(d['value'], d['e2e'], d['ms_per_step'], d['clocks'], d['cpu_baseline'])
print({k: round(v['ms'], 3) for k, v in d['roofline']['stages'].items()}, d['roofline']['whole_step'])
r = json.load(open('gpurun_out/bench_reference.json'))
print('reference arm', r.get('value'), r.get('cpu_baseline'))
PY
