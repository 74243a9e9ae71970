#!/usr/bin/env python
"""Wall-clock breakdown of one end-to-end solve (where do the non-kernel milliseconds go?)."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from structurepreservingiterativesolvers_b200 import _native as nat, solvers
from structurepreservingiterativesolvers_b200.device import KrylovContext

def T(label, fn, *a, **k):
    t0 = time.perf_counter(); out = fn(*a, **k); dt = time.perf_counter() - t0
    print(f"{label:40s} {dt*1e3:9.2f} ms", flush=True)
    return out

dic, x0, conlist, _ = bench.build_system(int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000)
A, b = dic["A"], dic["b"]; n = b.size
pinned = bench.pin_inputs(dic, x0, conlist)
for label, (Ax, bx, x0x, cl) in (("pageable", (A, b, x0, conlist)), ("pinned", pinned)):
    print("==", label)
    ctx = T("ctx create", KrylovContext, n, 50)
    T("upload A (+SELL conversion)", ctx.upload_matrix, nat.SLOT_A, Ax)
    T("upload b", ctx.upload_vec, nat.VEC_B, bx)
    T("upload x0", ctx.upload_vec, nat.VEC_X0, x0x)
    T("upload energy M", ctx.upload_matrix, nat.SLOT_CON0 + 1, cl[1].M)
    T("ctx close", ctx.close)
    sess = T("DeviceSession(...)", solvers.DeviceSession, Ax, bx, x0x, 50, conlist=cl)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for rep in range(3):
            x, info = T(f"cgmres(session) #{rep}", solvers.cgmres, Ax, bx, x0x, 50, tol=1e-6, conlist=cl, timing=True, small_solver="kkt", session=sess)
        print("steps", info["steps"], "timings", {k: (round(v, 5) if isinstance(v, float) else v) for k, v in info["timings"].items()})
        T("download x again", sess.ctx.download, nat.VEC_X)
        T("sync", sess.ctx.sync)
        T("arnoldi_step(0) alone", sess.ctx.arnoldi_step, 0)
        T("session close", sess.close)
        for rep in range(2):
            T(f"cgmres end-to-end #{rep}", solvers.cgmres, Ax, bx, x0x, 50, tol=1e-6, conlist=cl, timing=True, small_solver="kkt")
