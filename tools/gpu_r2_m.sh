#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], d['ms_per_step'], 'kernel ms', d.get('kernel_ms_per_step'))
for k,v in d.get('kernels',{}).items(): print('   ',k, round(v['ms_per_step'],3), v['launches_per_step'], round(v['frac_of_peak'] or 0,3), 'idle before', round(v['idle_before_ms_per_step'],3))
PY
}
for h in 1 0 1 0; do
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-extras --skip-e2e --skip-parity-mode --ctx-option hess_async=$h > gpurun_out/r2m_bench_h$h.json 2> gpurun_out/r2m_bench_h$h.err; echo "bench h=$h rc=$?"; show gpurun_out/r2m_bench_h$h.json | head -1
done
show gpurun_out/r2m_bench_h1.json
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2m_pytest.log
