#!/bin/bash
# Experiment harness. Build side (here):  tools/variants.sh build name "-DFOO -DBAR=1" ...
# GPU side (gpurun):                      tools/variants.sh run [frames] -> tools/stage_bench.py for every built
#                                         variant (stage times per 10 000 frames + a fingerprint of the records)
set -e
cd "$(dirname "$0")/.."
V=pywindow_b200/csrc/variants
if [ "$1" = build ]; then
  shift; mkdir -p $V; rm -f $V/libpwb_*.so
  while [ $# -gt 1 ]; do
    name=$1; flags=$2; shift 2
    (cd pywindow_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
       $flags -o variants/libpwb_$name.so pwb_kernels.cu pwb_periodic.cu pwb_hostio.cpp) &
  done
  wait; ls -la $V
else
  mkdir -p gpurun_out
  python tools/stage_bench.py ${2:-100000} 3 2>&1 | tail -1 | tee gpurun_out/variants.log
  for f in $V/libpwb_*.so; do
    PWB_LIBRARY=$PWD/$f timeout 300 python tools/stage_bench.py ${2:-100000} 3 2>&1 | tail -1 | tee -a gpurun_out/variants.log
  done
  if [ -n "$TESTS" ]; then PWB_LIBRARY=$PWD/$V/libpwb_$TESTS.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
fi
