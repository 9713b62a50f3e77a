#!/bin/bash
# quick GPU check: parity tests then a short bench (no CPU baseline)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/quick_tests.log
python bench.py --steps 10 --warmup 3 --skip-cpu --chunks 4 > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
cat gpurun_out/quick_tests.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'])
print({k:v.get('ms') for k,v in d['roofline']['stages'].items()})
PY
