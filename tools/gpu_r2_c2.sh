#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; head -c 300 gpurun_out/r2c_bench_default.json; tail -3 gpurun_out/r2c_bench_default.err
