#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras --skip-e2e --skip-parity-mode > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "bench n1 rc=$?"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --skip-extras --skip-e2e --skip-parity > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 $TR tools/dist_gpu_check.py 300000 auto lkdv > gpurun_out/r2g_dist_lkdv.log 2>&1; echo "dist lkdv rc=$?"; grep -E "OK|FAIL" gpurun_out/r2g_dist_lkdv.log | tail -3
