#!/bin/bash
# ncu --set full capture of k_opt_pore (one launch, 30 000 frames) for the default build and for the variants
# named on the command line; reports come back under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
F=${F:-30000}
run() {   # name, library ('' = default)
  PWB_LIBRARY=$2 ncu --set full --import-source on --clock-control none -k regex:"${K:-k_opt_pore}" --launch-skip ${SKIP:-1} -c 1 -f \
      -o gpurun_out/prof_r2_$1 python tools/stage_bench.py $F 1 > gpurun_out/ncu_r2_$1.log 2>&1
  tail -1 gpurun_out/ncu_r2_$1.log | cut -c1-160
}
run default ""
for v in "$@"; do run $v $PWD/pywindow_b200/csrc/variants/libpwb_$v.so; done
