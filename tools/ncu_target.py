#!/usr/bin/env python
"""Short profiling target: one device-resident CGMRES solve of the bench workload (for ncu)."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from structurepreservingiterativesolvers_b200 import solvers

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
workload = os.environ.get("SPIS_WORKLOAD", "lkdv")
dic, x0, conlist, _ = bench.build_system(n, workload)
sess = solvers.DeviceSession(dic["A"], dic["b"], x0, bench.K_KRYLOV, conlist=conlist)
warnings.simplefilter("ignore")
for _ in range(solves):
    x, info = solvers.cgmres(dic["A"], dic["b"], x0, bench.K_KRYLOV, tol=bench.WORKLOADS[workload]["tol"], contol=bench.CONTOL,
                             conlist=conlist, timing=True, small_solver="kkt", session=sess)
print("steps", info["steps"], "res", info["res"][-1])
