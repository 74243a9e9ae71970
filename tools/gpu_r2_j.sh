#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2j_pytest.log
timeout 600 python tools/tune_fw.py > gpurun_out/r2j_tune_fw.json 2> gpurun_out/r2j_tune_fw.err; echo "tune rc=$?"; cat gpurun_out/r2j_tune_fw.json
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-extras --skip-e2e --skip-parity-mode > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
print(d['ms_per_step'], 'kernel ms', d['kernel_ms_per_step'])
for k,v in d['kernels'].items(): print('   ',k, round(v['ms_per_step'],3), v['launches_per_step'], round(v['frac_of_peak'] or 0,3), 'idle before', round(v['idle_before_ms_per_step'],3))
PY
