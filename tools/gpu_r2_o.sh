#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
SPIS_TRACE=1 timeout 900 $TR bench.py --gpus 4 --steps 3 --warmup 3 --skip-extras --skip-parity --skip-e2e > gpurun_out/r2o_n4.json 2> gpurun_out/r2o_n4.err; echo "rc=$?"
grep -c "spis trace" gpurun_out/r2o_n4.err
