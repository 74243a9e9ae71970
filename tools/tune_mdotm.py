#!/usr/bin/env python
"""mdotm_kernel<4> (constraint catch-up: 4 right-hand sides per pass over Z) vs CTAs per SM."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.device import KrylovContext
n = 10_000_050
with KrylovContext(n, 50) as ctx:
    ctx.set_option("bench_mdotm_nw", 4)
    for m in (4, 8, 12, 16, 21, 32, 50):
        for ctas in (1, 2, 3, 4, 6, 8):
            ctx.set_option("mdotm_ctas_per_sm", ctas)
            ms, by = ctx.bench_kernel(nat.PROF_MDOT, m, reps=10)
            print(json.dumps(dict(m=m, ctas=ctas, us=round(ms * 1e3, 1), gbs=round(by / ms * 1e-6))), flush=True)
