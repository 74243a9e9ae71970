#!/bin/bash
# usage: gpu_r2_scale.sh N TAG [extra bench flags]   -- one multi-rank bench line on N GPUs of the box
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err; echo "bench n$N rc=$?"
python - gpurun_out/${TAG}_n$N.json <<'PY'
import json,sys
ls=[l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')]
if not ls: print("no json line"); sys.exit(0)
d=json.loads(ls[-1])
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'kernel ms', d.get('kernel_ms_per_step'), 'e2e', d.get('e2e'))
print('parity_vs_single', d.get('parity_vs_single')); print('xcomm', d.get('xcomm'))
for k,v in d.get('kernels',{}).items(): print('   ',k, round(v['ms_per_step'],3), v['launches_per_step'], round(v['frac_of_peak'] or 0,3), 'idle before', round(v['idle_before_ms_per_step'],3))
for k,v in (d.get('extra') or {}).items(): print('   extra', k, v if not isinstance(v, dict) else {a:b for a,b in v.items() if a in ('ms_per_step','value','steps','parity','n','skipped')})
PY
tail -3 gpurun_out/${TAG}_n$N.err
