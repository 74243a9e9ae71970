#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tune_fw.py > gpurun_out/r2l_tune_fw.json 2> gpurun_out/r2l_tune_fw.err; echo "tune rc=$?"; cat gpurun_out/r2l_tune_fw.json; tail -3 gpurun_out/r2l_tune_fw.err
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -k "fw or field or window" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest.log
