#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/final_smoke.log
( time timeout 1500 python bench.py > gpurun_out/bench_r2_lkdv_n1.json 2> gpurun_out/bench_r2_lkdv_n1.err ) 2>&1 | grep real
python - gpurun_out/bench_r2_lkdv_n1.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', d['e2e'], 'parity', d['parity']['rel_diff'], d['parity']['steps'], 'roofline frac', round(d['roofline']['frac'],3), 'gpu_launches', d['gpu_launches'], 'clocks', d['clocks'])
print('cpu_baseline', d['cpu_baseline']['value'])
for k,v in d['extra'].items(): print('  ', k, round(v.get('ms_per_step',0),3))
PY
