#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python bench.py > gpurun_out/bench_r2_lkdv_n1.json 2> gpurun_out/bench_r2_lkdv_n1.err ) 2> gpurun_out/bench_r2_lkdv_n1.time; echo "bench rc=$?"; cat gpurun_out/bench_r2_lkdv_n1.time
( time timeout 1500 python bench.py --impl reference > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err ) 2> gpurun_out/bench_r2_reference.time; echo "ref rc=$?"; cat gpurun_out/bench_r2_reference.time
tail -c 600 gpurun_out/bench_r2_reference.json
