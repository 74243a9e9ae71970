#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --maxfail=25 --timeout 600 -k "field_window" > gpurun_out/r2i_pytest_fw.log 2>&1; echo "pytest fw rc=$?"; tail -4 gpurun_out/r2i_pytest_fw.log
timeout 600 python tools/tune_fw.py > gpurun_out/r2i_tune_fw.json 2> gpurun_out/r2i_tune_fw.err; echo "tune rc=$?"; cat gpurun_out/r2i_tune_fw.json; tail -3 gpurun_out/r2i_tune_fw.err
