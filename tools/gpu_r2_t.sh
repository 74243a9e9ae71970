#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -k "one_pass or ride_in" > gpurun_out/r2t_pytest_gram.log 2>&1; echo "pytest gram rc=$?"; tail -30 gpurun_out/r2t_pytest_gram.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-extras --skip-parity-mode --e2e-steps 3 --trace-file gpurun_out/r2t_trace.json > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python - gpurun_out/r2t_bench.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'],3), 'kernel ms', d.get('kernel_ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'))
for k,v in d.get('kernels',{}).items(): print('   ',k, round(v['ms_per_step'],3), v['launches_per_step'], round(v['frac_of_peak'] or 0,3), 'idle before', round(v['idle_before_ms_per_step'],3))
PY
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2t_pytest.log
