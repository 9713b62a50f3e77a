#!/bin/bash
# Scaling run on one box: bench.py at N = 1 and N = $1 GPUs (torchrun, as the driver launches it); JSON lines under
# gpurun_out/.  usage: gpurun --gpus N -- bash tools/r2_scale.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python bench.py --skip-cpu > gpurun_out/r2_scale_1.json 2> gpurun_out/r2_scale_1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --skip-cpu > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
tail -3 gpurun_out/r2_scale_$N.err
python - $N <<'PY'
import json, sys
for n in ("1", sys.argv[1]):
    try:
        d = json.loads(open(f"gpurun_out/r2_scale_{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "unreadable", e); continue
    c4 = (d.get("secondary") or {}).get("config4_periodic_trajectory") or {}
    print(f"N={n}: value {d['value']:.4g} e2e {d['e2e']['value']:.4g} api {d['api']['value']:.4g} "
          f"(decode {d['api'].get('host_decode_frames_per_s', 0):.3g}/s on {d['api'].get('host_threads')} threads) "
          f"cfg4 {c4.get('cell_frames_per_s', 0):.4g} cell-frames/s  ms/step {d['ms_per_step']:.2f} clocks {d['clocks']['sm_mhz']}")
PY
