#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 3 4 6 8 2 4; do
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-extras --skip-parity-mode --skip-e2e --early-download-chunks $c > gpurun_out/r2u_bench_c$c.json 2> gpurun_out/r2u_bench_c$c.err; echo "bench chunks=$c rc=$?"
python - gpurun_out/r2u_bench_c$c.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'],3), 'kernel ms', round(d.get('kernel_ms_per_step'),3))
PY
done
