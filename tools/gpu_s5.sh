#!/bin/bash
# round-2 session 5: reference-log tests after the far-field fix of the 2D mesher
mkdir -p gpurun_out
L=gpurun_out/s5.log
: > $L
rm -f gpurun_out/golden_stats.json
REMO_GOLDEN_STATS=gpurun_out/golden_stats.json timeout 2400 python -m pytest tests -m gpu -q -k "reference_logs or golden_example01" > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -12 gpurun_out/s5_pytest.log >> $L
cat $L
