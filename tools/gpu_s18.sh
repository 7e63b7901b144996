#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size.py -q -s > gpurun_out/s18_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "size|passed|failed|^E " gpurun_out/s18_pytest.log | tail -12
