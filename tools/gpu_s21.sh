#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_c3.py tests/test_gpu_full_size.py -q > gpurun_out/s21_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s21_pytest.log
