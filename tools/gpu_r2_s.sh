#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'],3), 'kernel ms', d.get('kernel_ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'))
PY
}
for v in early noearly early noearly; do
F=""; [ $v = noearly ] && F="--no-early-download"
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-extras --skip-parity-mode --e2e-steps 3 $F > gpurun_out/r2s_bench_$v.json 2> gpurun_out/r2s_bench_$v.err; echo "bench $v rc=$?"; show gpurun_out/r2s_bench_$v.json
done
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s_pytest.log
