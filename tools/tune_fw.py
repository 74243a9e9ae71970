"""Field-window SpMV (x staged in shared memory) against the L1-gather row-pattern kernels on the 1e7 lkdv operator."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from structurepreservingiterativesolvers_b200 import _native as nat, solvers, wrappers
from structurepreservingiterativesolvers_b200.problems import lkdv
warnings.simplefilter("ignore")
M = lkdv.benchmark_size(int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000)
d, _ = lkdv.linforms(space="CG", M=M, mlength=0.8 * M)
x0 = np.zeros(d["b"].size)
full = wrappers.lkdv.conlist(d, x0)
cl = [full[0], full[2]]
out = {}
for fw, rows in ((7, 4), (7, 8), (0, 8)):
    sess = solvers.DeviceSession(d["A"], d["b"], x0, 50, conlist=cl, profile=True)
    ctx = sess.ctx
    ctx.set_option("spmv_fw", fw)
    ctx.set_option("spmv_fw_rows", rows)
    r = {"fw_fields": ctx.info("fw_fields:0")}
    for mode in (0, 1, 2):
        ms, by = ctx.bench_kernel(nat.PROF_SPMV, mode, 50)
        r[f"mode{mode}_us"] = round(ms * 1e3, 1)
    for rep in range(3):
        ctx.reset_profile()
        x, info = solvers.cgmres(d["A"], d["b"], x0, 50, tol=1e-6, contol=10, conlist=cl, timing=True, small_solver="kkt", session=sess)
        ctx.sync()
    p = ctx.profile()
    r["solve_spmv_ms"] = round(p["spmv"]["ms"], 3); r["solve_spmv_launches"] = p["spmv"]["launches"]
    r["solve_spmv_aux_ms"] = round(p["spmv_aux"]["ms"], 3); r["steps"] = info["steps"]
    r["spmv_gbs_moved"] = round(p["spmv"]["gbs_moved"], 1)
    out[{7: f"fw_all_rows{rows}", 0: "gather"}[fw]] = r
    sess.close()
print(json.dumps(out, indent=1))
