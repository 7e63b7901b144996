#!/bin/bash
# compute-sanitizer over the smoke-size cases of every product kernel family:
#   memcheck + racecheck + initcheck on tests/sanitize_cases.py, which runs one small mesh task per (order, dim, preconditioner,
#   SpMM kind) through the C ABI and checks Ra against the oracle.  Logs: gpurun_out/sanitize_*.log (summaries are copied to
#   profiles/ by hand).  Usage on a GPU box:  bash tools/sanitize.sh
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
# Pools that close compute-sanitizer print a notice and exit at once: then the cases still run, plainly, with the library's own
# table validator on (remo_set_option("ebe_check", 1)) and every result compared with the CPU oracle.
timeout 900 python tests/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain (ebe_check=1, oracle comparison) rc=$?" | tee gpurun_out/sanitize_summary.log
grep -E "case |all cases" gpurun_out/sanitize_plain.log >> gpurun_out/sanitize_summary.log
for tool in memcheck racecheck initcheck; do
  timeout 1500 $S --tool $tool --error-exitcode 9 --print-limit 20 python tests/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case |closed on this pool" gpurun_out/sanitize_$tool.log | tail -20 >> gpurun_out/sanitize_summary.log
done
cat gpurun_out/sanitize_summary.log
