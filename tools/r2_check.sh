#!/bin/bash
# Round-2 GPU check: parity tests, a small bench (all legs), the default bench, the reference arm (short).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests.log
cat gpurun_out/r2_tests.log
timeout 600 python bench.py --frames 20000 --steps 5 --warmup 3 --api-frames 20000 --cell-frames 512 --skip-cpu \
  > gpurun_out/r2_small.json 2> gpurun_out/r2_small.err
tail -5 gpurun_out/r2_small.err
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -5 gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2_reference.json 2> gpurun_out/r2_reference.err
python - <<'PY'
import json
for f in ("r2_small", "r2_bench"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.4g e2e %.4g ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"])
    print({k: round(v["ms"], 3) for k, v in d["roofline"]["stages"].items()}, d["roofline"]["kernel"], d["roofline"]["frac"])
    print("api", json.dumps(d["api"])[:900])
    print("cpu", d["cpu_baseline"])
    print("secondary", json.dumps(d["secondary"])[:3000])
try:
    r = json.load(open("gpurun_out/r2_reference.json")); print("reference arm", r.get("value"), r.get("cpu_baseline"))
except Exception as e:
    print("reference arm unreadable", e)
PY
