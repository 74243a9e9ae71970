#!/usr/bin/env python
"""Short profiling target: device-resident CGMRES solves of a bench workload (for ncu).

    SPIS_WORKLOAD=lkdv|swe|lkdvRK|jacobi|csrpre python tools/ncu_target.py [n] [solves]

jacobi / csrpre: the lkdv workload with the point-Jacobi preconditioner on the device (fused into the last sweep) / the
same diagonal handed over as a sparse matrix (SPIS_PRE_CSR path); lkdvRK brings the 6x6 block-diagonal kernel."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scipy.sparse as sps
import bench
from structurepreservingiterativesolvers_b200 import solvers
from structurepreservingiterativesolvers_b200.preconditioners import JacobiPreconditioner

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
workload = os.environ.get("SPIS_WORKLOAD", "lkdv")
base = workload if workload in bench.WORKLOADS else "lkdv"
dic, x0, conlist, _, pre, _ = bench.build_system(n, base)
if workload == "jacobi":
    pre = JacobiPreconditioner(dic["A"])
elif workload == "csrpre":
    pre = (sps.diags(1.0 / dic["A"].diagonal()) + 0.0 * dic["A"]).tocsr()       # a sparse P with A's pattern: z = P q by SpMV
tol = bench.workload_tol(base, dic)
sess = solvers.DeviceSession(dic["A"], dic["b"], x0, bench.K_KRYLOV, conlist=conlist, pre=pre)
warnings.simplefilter("ignore")
for _ in range(solves):
    x, info = solvers.cgmres(dic["A"], dic["b"], x0, bench.K_KRYLOV, tol=tol, contol=bench.CONTOL,
                             conlist=conlist, pre=pre, timing=True, small_solver="kkt", session=sess)
print("steps", info["steps"], "res", info["res"][-1])
