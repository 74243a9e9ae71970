#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -k "one_pass_constraint" > gpurun_out/r2q_pytest_gram.log 2>&1; echo "pytest gram rc=$?"; tail -30 gpurun_out/r2q_pytest_gram.log
for g in 1 0; do
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-extras --skip-e2e --skip-parity-mode --ctx-option gram=$g --trace-file gpurun_out/r2q_trace_g$g.json > gpurun_out/r2q_bench_g$g.json 2> gpurun_out/r2q_bench_g$g.err; echo "bench gram=$g rc=$?"
python - gpurun_out/r2q_bench_g$g.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], d['ms_per_step'], 'kernel ms', d.get('kernel_ms_per_step'), 'parity', d.get('parity'))
for k,v in d.get('kernels',{}).items(): print('   ',k, round(v['ms_per_step'],3), v['launches_per_step'], round(v['frac_of_peak'] or 0,3), 'idle before', round(v['idle_before_ms_per_step'],3))
PY
done
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2q_pytest.log
