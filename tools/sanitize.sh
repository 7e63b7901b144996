#!/bin/bash
# compute-sanitizer over the smoke-size cases of every product kernel family:
#   memcheck + racecheck + initcheck on tools/sanitize_case.py, which runs one small mesh task per (order, dim, preconditioner,
#   SpMM kind) through the C ABI and checks Ra against the oracle.  Logs: gpurun_out/sanitize_*.log (summaries are copied to
#   profiles/ by hand).  Usage on a GPU box:  bash tools/sanitize.sh
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck initcheck; do
  timeout 1500 $S --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case " gpurun_out/sanitize_$tool.log | tail -20 >> gpurun_out/sanitize_summary.log
done
cat gpurun_out/sanitize_summary.log
