#!/bin/bash
# round 2, 2-GPU call: sharded parity (persistent communicator, per-call sessions, zero-v slice), soak, strong-scaling bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log; tail -4 gpurun_out/r2b_pytest.log
timeout 600 $TR tools/dist_gpu_check.py 300000 auto lkdv > gpurun_out/r2b_dist_lkdv.log 2>&1; echo "dist lkdv rc=$?"; grep -E "OK|FAIL|Error|error" gpurun_out/r2b_dist_lkdv.log | tail -5
timeout 600 $TR tools/dist_gpu_check.py 1000000 auto swe > gpurun_out/r2b_dist_swe.log 2>&1; echo "dist swe rc=$?"; grep -E "OK|FAIL|Error|error" gpurun_out/r2b_dist_swe.log | tail -5
timeout 600 $TR tools/dist_gpu_check.py 30000 auto lkdv soak 200 > gpurun_out/r2b_soak.log 2>&1; echo "soak rc=$?"; grep -E "soak|Error|error" gpurun_out/r2b_soak.log | tail -3
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --skip-extras > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench n2 rc=$?"; head -c 300 gpurun_out/r2b_bench_n2.json; tail -3 gpurun_out/r2b_bench_n2.err
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --skip-extras --skip-e2e --skip-parity --host-loop > gpurun_out/r2b_bench_n2_host.json 2> gpurun_out/r2b_bench_n2_host.err; echo "bench n2 host rc=$?"; head -c 300 gpurun_out/r2b_bench_n2_host.json
