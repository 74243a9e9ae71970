#!/bin/bash
mkdir -p gpurun_out
SPIS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extras --skip-parity-mode --skip-e2e > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "rc=$?"
grep "spis trace" gpurun_out/r2w_bench.err | tail -12
