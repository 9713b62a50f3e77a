#!/bin/bash
# Round-end evidence, part A (profilers): launch list of a short bench run + one ncu --set full capture of the
# five pipeline kernels on 10 000-frame launches.  Part B (tools/final_bench.sh) is the un-profiled bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/final_launches.log 2>&1
# chunks=1 so that every kernel launch covers the whole 10 000-frame batch; the first analysis of the run is
# the engine's own warm-up, skipped per kernel name below
ncu --set full --import-source on --clock-control none \
    -k regex:"k_prepare|k_opt_pore|k_rays|k_cluster|k_windows" --launch-skip 24 -c 6 -f -o gpurun_out/prof_final \
    python bench.py --steps 2 --warmup 0 --skip-cpu --chunks 1 > gpurun_out/final_ncu.log 2>&1
tail -2 gpurun_out/final_ncu.log | cut -c1-200
