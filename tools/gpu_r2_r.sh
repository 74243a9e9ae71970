#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_mdot_insitu.py > gpurun_out/r2r_mdot_insitu.json 2> gpurun_out/r2r_mdot_insitu.err; echo "rc=$?"; tail -3 gpurun_out/r2r_mdot_insitu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_mdot_insitu.json').read().strip().splitlines()[-1])
for k in ('ride','plain'):
    print(k); 
    for c,v in d[k].items(): print('  ',c, v)
print(d['mdot_alone_us']); print(d['orthmid_alone_us'])
PY
