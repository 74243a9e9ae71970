#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2k_pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -12 gpurun_out/r2k_pytest_kernels.log
timeout 900 python tools/tune_sellw.py > gpurun_out/r2k_tune_sellw.json 2> gpurun_out/r2k_tune_sellw.err; echo "tune rc=$?"; cat gpurun_out/r2k_tune_sellw.json; tail -3 gpurun_out/r2k_tune_sellw.err
