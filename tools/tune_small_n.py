"""Fixed cost of the reducing sweeps at row-sharded sizes: mdot / orth_mid / lincomb alone on ONE GPU at n = 1.25M
and 2.5M rows (an eighth / a quarter of the 1e7 system), per grid size."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.device import KrylovContext
out = []
for n in (1_250_050, 2_500_050, 10_000_050):
    with KrylovContext(n, 44) as ctx:
        for m in (1, 4, 11, 21):
            rec = {"n": n, "m": m}
            for per in (1, 2, 4):
                ctx.set_option("mdot_reg_ctas_per_sm", per)
                ms, by = ctx.bench_kernel(nat.PROF_MDOT, m, reps=50)
                rec[f"mdot_c{per}_us"] = round(ms * 1e3, 1)
            ms, by = ctx.bench_kernel(nat.PROF_ORTHMID, m, reps=50)
            rec["orthmid_us"] = round(ms * 1e3, 1)
            ms, by = ctx.bench_kernel(nat.PROF_LINCOMB, m, reps=50)
            rec["lincomb_us"] = round(ms * 1e3, 1)
            rec["ideal_mdot_us"] = round((m + 1) * 8 * n / 6.5e12 * 1e6, 1)
            out.append(rec)
            print(json.dumps(rec), flush=True)
