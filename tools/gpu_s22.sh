#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python bench.py --mode example01 > gpurun_out/s22_example01_$i.json 2> /dev/null; python - $i <<PY
import json, sys
d = json.load(open('gpurun_out/s22_example01_%s.json' % sys.argv[1]))
print('example01 run', sys.argv[1], 'value', round(d['value'], 1), 'wall ms', round(d['ms_per_step']), 'busy', round(d['config']['gpu_busy_fraction'], 3), 'parity', d['parity']['max_rel_err_ra'], 'vs_baseline', round(d['vs_baseline'], 1))
PY
done
timeout 900 python -m pytest tests/test_gpu_reference_logs.py -q 2>&1 | tail -2
