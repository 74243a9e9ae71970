#!/bin/bash
mkdir -p gpurun_out
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
for v in early noearly; do
F=""; [ $v = noearly ] && F="--no-early-download"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --skip-extras --skip-parity --skip-e2e $F --trace-file gpurun_out/r2z_trace_$v.json > gpurun_out/r2z_$v.json 2> gpurun_out/r2z_$v.err; echo "rc=$?"
python - gpurun_out/r2z_$v.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'kernel ms', round(d.get('kernel_ms_per_step'),3), d['ms_each_step'] if 'ms_each_step' in d else '')
PY
python tools/show_trace.py gpurun_out/r2z_trace_$v.json.rank0 | tail -5
done
