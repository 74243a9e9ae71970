#!/bin/bash
mkdir -p gpurun_out
for v in 3 4 5 6 8; do
timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-parity-mode --skip-e2e --extras-budget-s 12 --ctx-option spmv_dual_ctas_per_sm=$v > gpurun_out/r2gm_$v.json 2> gpurun_out/r2gm_$v.err
python - gpurun_out/r2gm_$v.json $v <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
e=d['extra']['general_matrix']
print('dual ctas/SM', sys.argv[2], 'lkdv', round(d['ms_per_step'],3), 'general_matrix', round(e['ms_per_step'],3), 'spmv', round(e['kernels']['spmv']['ms_per_step'],3), 'frac', round(e['spmv_frac_dram'],3))
PY
done
