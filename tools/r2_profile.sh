#!/bin/bash
# Round-2 evidence (profilers): launch list of a short bench run and one ncu --set full capture of the pipeline
# kernels on 100 000-frame launches (the bench's step size).  Summaries are made here afterwards:
#   python tools/ncu_summary.py gpurun_out/prof_r2_final.ncu-rep > profiles/r2_ncu_full_kernels.txt
#   python tools/make_traffic.py gpurun_out/prof_r2_final.ncu-rep 100000
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu --skip-api --skip-secondary > gpurun_out/r2_launches.log 2>&1
ncu --set full --import-source on --clock-control none \
    -k regex:"k_prepare|k_opt_pore|k_rays|k_cluster|k_windows" --launch-skip 7 -c 7 -f -o gpurun_out/prof_r2_final \
    python tools/stage_bench.py 100000 1 > gpurun_out/r2_ncu.log 2>&1
tail -1 gpurun_out/r2_ncu.log | cut -c1-200
