"""Where does a pipelined solve on a SELLD-format system spend its wall time?  (SPIS_TRACE=1)"""
import os, sys, time, warnings
os.environ["SPIS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from structurepreservingiterativesolvers_b200 import solvers, wrappers, _native as nat
from structurepreservingiterativesolvers_b200.problems import swe
warnings.simplefilter("ignore")
M = swe.benchmark_size(10_000_000)
d, _ = swe.linforms(M=M, mlength=0.8 * M, sort=False)
x0 = np.zeros(d["b"].size)
cl = wrappers.swe.conlist(d, x0)
for pipe in (True, False):
    solvers.configure(pipeline=pipe)
    sess = solvers.DeviceSession(d["A"], d["b"], x0, 50, conlist=cl)
    for rep in range(3):
        t0 = time.perf_counter()
        x, info = solvers.cgmres(d["A"], d["b"], x0, 50, tol=1e-7, contol=10, conlist=cl, timing=True, small_solver="kkt", session=sess)
        sess.ctx.sync()
        print("pipeline", pipe, "rep", rep, "%.2f ms" % (1e3 * (time.perf_counter() - t0)), "steps", info["steps"], "pinned out", nat.pinned_outstanding(), type(x.base).__name__, flush=True)
    sess.close()
