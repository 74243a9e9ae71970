#!/bin/bash
# 2-GPU call: new reduction tail + LL cross-GPU protocol -- parity, soak, N=1 and N=2 timings
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -3 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench n1 rc=$?"; head -c 200 gpurun_out/r2f_bench_n1.json; echo
timeout 600 $TR tools/dist_gpu_check.py 30000 auto lkdv soak 300 > gpurun_out/r2f_soak.log 2>&1; echo "soak rc=$?"; grep -E "soak|Error|error" gpurun_out/r2f_soak.log | tail -3
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --skip-extras > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 rc=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2f_bench_n2.json | head -1
