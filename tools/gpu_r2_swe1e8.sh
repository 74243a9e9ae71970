#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python bench.py --workload swe --dofs 100000000 --steps 3 --warmup 3 --skip-e2e --skip-cpu --skip-extras --skip-parity-mode > gpurun_out/bench_r2_swe1e8_n1.json 2> gpurun_out/bench_r2_swe1e8_n1.err ) 2>&1 | tail -4
python - gpurun_out/bench_r2_swe1e8_n1.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'kernel ms', round(d.get('kernel_ms_per_step'),3))
for k,v in d.get('kernels',{}).items(): print('   ',k, round(v['ms_per_step'],3), v['launches_per_step'], round(v['frac_of_peak'] or 0,3))
PY
tail -3 gpurun_out/bench_r2_swe1e8_n1.err
