#!/bin/bash
# Experiment harness. Build side (here):  tools/variants.sh build name "-DFOO -DBAR=1" ...
# GPU side (gpurun):                      tools/variants.sh run   -> quick bench of every built variant
set -e
cd "$(dirname "$0")/.."
V=pywindow_b200/csrc/variants
if [ "$1" = build ]; then
  shift; mkdir -p $V
  while [ $# -gt 1 ]; do
    name=$1; flags=$2; shift 2
    (cd pywindow_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
       $flags -o variants/libpwb_$name.so pwb_kernels.cu pwb_periodic.cu pwb_hostio.cpp) &
  done
  wait; ls -la $V
else
  mkdir -p gpurun_out
  cp pywindow_b200/csrc/libpwb.so /tmp/libpwb_orig.so
  for f in $V/libpwb_*.so; do
    cp $f pywindow_b200/csrc/libpwb.so
    n=$(basename $f .so)
    if python bench.py --steps ${STEPS:-10} --warmup 3 --skip-cpu --chunks 4 > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err; then
      python - $n <<'PY'
import json, sys
d = json.load(open('gpurun_out/var_%s.json' % sys.argv[1]))
print(sys.argv[1], round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 3),
      {k[:7]: round(v.get('ms'), 3) for k, v in d['roofline']['stages'].items()})
PY
    else echo "$n FAILED"; tail -3 gpurun_out/var_$n.err; fi
  done
  cp /tmp/libpwb_orig.so pywindow_b200/csrc/libpwb.so
  if [ -n "$TESTS" ]; then cp $V/libpwb_$TESTS.so pywindow_b200/csrc/libpwb.so; python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
fi
