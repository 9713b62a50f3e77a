#!/bin/bash
# GPU tests + variant stage timings + a short bench (api leg at full size)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|noise|free-running|gpu |well-cond|Error|assert" | tail -15
bash tools/variants.sh run 100000 | sed 's#/root/repo/pywindow_b200/csrc/variants/##' | cut -c1-170
for c in ${CHUNKS:-4}; do
timeout 600 python bench.py --frames ${FRAMES:-100000} --steps 5 --warmup 3 --skip-cpu --skip-secondary --chunks $c > gpurun_out/r2_quick_c$c.json 2> gpurun_out/r2_quick_c$c.err
python - $c <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2_quick_c{sys.argv[1]}.json"))
print("chunks", sys.argv[1], "value %.4g e2e %.4g ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
      {k[:7]: round(v["ms"], 2) for k, v in d["roofline"]["stages"].items()})
a = d["api"]
print("api %.4g frames/s" % a["value"], a["seconds"], "decode %.4g" % a["host_decode_frames_per_s"])
PY
done
