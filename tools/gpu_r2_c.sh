#!/bin/bash
# round 2, 1-GPU call: parity suite again, the FULL default bench line (oracle parity, extras), the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -4 gpurun_out/r2c_pytest.log
nproc; free -g | head -2
/usr/bin/time -v timeout 1500 python bench.py > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2c_bench_default.json; grep -E "Elapsed|Maximum resident" gpurun_out/r2c_bench_default.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err; echo "ref rc=$?"; head -c 1200 gpurun_out/r2c_bench_reference.json
