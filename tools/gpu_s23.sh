#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s23_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s23_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-companions 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3))"
