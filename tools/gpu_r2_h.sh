#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --maxfail=25 --timeout 600 -k "field_window or spmv" > gpurun_out/r2h_pytest_fw.log 2>&1; echo "pytest fw rc=$?"; tail -5 gpurun_out/r2h_pytest_fw.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras --skip-e2e --skip-parity-mode > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2h_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
print(d['ms_per_step'], 'kernel ms', d['kernel_ms_per_step'])
for k,v in d['kernels'].items(): print('   ',k, round(v['ms_per_step'],3), v['launches_per_step'], round(v['frac_of_peak'] or 0,3), 'idle before', round(v['idle_before_ms_per_step'],3))
PY
