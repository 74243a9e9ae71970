#!/bin/bash
# 2 GPUs: the whole GPU suite (distributed tests included), the distributed checks, the N=2 bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 --timeout 900 > gpurun_out/r2_2gpu_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_2gpu_pytest.log
timeout 600 $TR tools/dist_gpu_check.py 300000 auto lkdv > gpurun_out/r2_2gpu_dist_lkdv.log 2>&1; echo "dist lkdv rc=$?"; grep -E "OK|FAIL|rel" gpurun_out/r2_2gpu_dist_lkdv.log | tail -6
timeout 600 $TR tools/dist_gpu_check.py 300000 auto swe > gpurun_out/r2_2gpu_dist_swe.log 2>&1; echo "dist swe rc=$?"; grep -E "OK|FAIL|rel" gpurun_out/r2_2gpu_dist_swe.log | tail -6
bash tools/gpu_r2_scale.sh 2 r2_2gpu
