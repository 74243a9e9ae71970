#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus 4 --steps 3 --warmup 3 --skip-extras --skip-parity --skip-e2e --trace-file gpurun_out/r2p_trace_n4.json > gpurun_out/r2p_n4.json 2> gpurun_out/r2p_n4.err; echo "rc=$?"
tail -3 gpurun_out/r2p_n4.err
timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extras --skip-e2e --skip-parity-mode --trace-file gpurun_out/r2p_trace_n1.json > gpurun_out/r2p_n1.json 2> gpurun_out/r2p_n1.err; echo "rc=$?"
