#!/usr/bin/env python
"""mdot: per-tile warp reductions (mdot_kernel, auto variant) against per-thread register sums (mdot_reg_kernel),
n = 1e7, resident data, rows 1..40."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.device import KrylovContext

n = int(os.environ.get("SPIS_N", 10_000_050))
rows = []
with KrylovContext(n, 44) as ctx:
    for m in (1, 2, 3, 4, 5, 6, 8, 10, 12, 14, 16, 18, 20, 24, 28, 32, 36, 40):
        ctx.set_option("mdot_variant", 0); ctx.set_option("mdot_reg_auto", 0)
        ms, by = ctx.bench_kernel(nat.PROF_MDOT, m, reps=20)
        rec = dict(m=m, legacy_us=ms * 1e3, legacy_gbs=by / ms * 1e-6)
        ctx.set_option("mdot_variant", 1)
        for ctas in (0, 2, 3, 4):
            ctx.set_option("mdot_reg_ctas_per_sm", ctas)
            ms, by = ctx.bench_kernel(nat.PROF_MDOT, m, reps=20)
            rec[f"reg_c{ctas}_us"] = ms * 1e3
        rows.append(rec)
        print(json.dumps({k: round(v, 1) if isinstance(v, float) else v for k, v in rec.items()}), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_mdot_reg.json"), "w"), indent=1)
