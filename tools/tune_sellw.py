"""SELLW (staged x windows + 16-bit columns) against the plain SELL / SELLD kernels on the 1e7 swe and lkdv operators."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from structurepreservingiterativesolvers_b200 import _native as nat, solvers, wrappers
from structurepreservingiterativesolvers_b200.problems import lkdv, swe
warnings.simplefilter("ignore")
out = {}
for wl in ("swe", "lkdv_sell"):
    if wl == "swe":
        M = swe.benchmark_size(10_000_000)
        d, _ = swe.linforms(M=M, mlength=0.8 * M, sort=False)
        x0 = np.zeros(d["b"].size); cl = wrappers.swe.conlist(d, x0); tol = 1e-7; fmt = "auto"
    else:
        M = lkdv.benchmark_size(10_000_000)
        d, _ = lkdv.linforms(space="CG", M=M, mlength=0.8 * M)
        x0 = np.zeros(d["b"].size); full = wrappers.lkdv.conlist(d, x0); cl = [full[0], full[2]]; tol = 1e-6; fmt = "sell"
    for sw in (7, 1, 0):
        sess = solvers.DeviceSession(d["A"], d["b"], x0, 50, conlist=cl, profile=True, spmv_format=fmt)
        ctx = sess.ctx
        ctx.set_option("spmv_sellw", sw)
        r = {"cap": ctx.info("sellw_cap:0"), "fmt": ctx.info("fmt:0")}
        for mode in (0, 1, 2):
            ms, by = ctx.bench_kernel(nat.PROF_SPMV, mode, 30)
            r[f"mode{mode}_us"] = round(ms * 1e3, 1)
        for rep in range(3):
            ctx.reset_profile()
            x, info = solvers.cgmres(d["A"], d["b"], x0, 50, tol=tol, contol=10, conlist=cl, timing=True, small_solver="kkt", session=sess)
            ctx.sync()
        p = ctx.profile()
        r["solve_spmv_ms"] = round(p["spmv"]["ms"], 3); r["solve_spmv_launches"] = p["spmv"]["launches"]
        r["solve_spmv_aux_ms"] = round(p["spmv_aux"]["ms"], 3); r["steps"] = info["steps"]
        r["spmv_gbs_moved"] = round(p["spmv"]["gbs_moved"], 1); r["kernel_ms"] = round(sum(v["ms"] for v in p.values()), 3)
        out[f"{wl}:{ {7: 'sellw_all', 1: 'sellw_dual_only', 0: 'plain'}[sw]}"] = r
        sess.close()
print(json.dumps(out, indent=1))
