#!/bin/bash
# round 2, first GPU call: parity suite + the default workload on the device-resident and the host-driven loop
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras > gpurun_out/r2a_bench_pipe.json 2> gpurun_out/r2a_bench_pipe.err
echo "bench pipe rc=$?"; head -c 600 gpurun_out/r2a_bench_pipe.json
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras --host-loop > gpurun_out/r2a_bench_host.json 2> gpurun_out/r2a_bench_host.err
echo "bench host rc=$?"; head -c 600 gpurun_out/r2a_bench_host.json
