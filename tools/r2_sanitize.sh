#!/bin/bash
# compute-sanitizer evidence: memcheck over every single-GPU test, racecheck over the periodic tests and a few parity
# tests (racecheck is ~100x slower).  Logs under gpurun_out/; the summaries go to profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_memcheck.log \
  python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py 2>&1 | tail -3 > gpurun_out/r2_memcheck_tests.txt
echo "memcheck exit code $?" >> gpurun_out/r2_memcheck_tests.txt
grep -E "ERROR SUMMARY|Invalid|Error" gpurun_out/r2_memcheck.log | sort | uniq -c | head -20 >> gpurun_out/r2_memcheck_tests.txt
cat gpurun_out/r2_memcheck_tests.txt
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis --log-file gpurun_out/r2_racecheck.log \
  python -m pytest tests/test_gpu_periodic.py tests/test_gpu_parity.py -m gpu -x -q -k "periodic or discrete or cc3 or edge or dense or pipeline_kernels or modular" 2>&1 | tail -3 > gpurun_out/r2_racecheck_tests.txt
grep -E "RACECHECK SUMMARY" gpurun_out/r2_racecheck.log >> gpurun_out/r2_racecheck_tests.txt
grep -oE "pwb::[A-Za-z_:0-9<>]+" gpurun_out/r2_racecheck.log | sort | uniq -c | sort -rn | head -30 >> gpurun_out/r2_racecheck_tests.txt
cat gpurun_out/r2_racecheck_tests.txt
