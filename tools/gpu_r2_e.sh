#!/bin/bash
# round 2, 4-GPU call: strong scaling of the default workload + swe at 1e8 (extra), sharded parity at 4 ranks
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
S=$(date +%s)
timeout 1200 $TR bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n4.json 2> gpurun_out/r2e_bench_n4.err; echo "bench n4 rc=$? wall $(( $(date +%s) - S )) s"; head -c 300 gpurun_out/r2e_bench_n4.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2e_bench_n4.err | tail -5
timeout 600 $TR tools/dist_gpu_check.py 300000 auto lkdv > gpurun_out/r2e_dist_lkdv.log 2>&1; echo "dist lkdv rc=$?"; grep -E "OK|FAIL|Error|error" gpurun_out/r2e_dist_lkdv.log | tail -5
