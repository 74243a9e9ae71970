"""The signed 1e-12 acceptance test of the reference (solvers.py:266, quirk Q4) in non-timing mode is decided by the last
bits of the reduced constraint terms: the same small lkdv solve with the two summation orders of the constraint
reduction (gram = 1: one-pass tensor-core Gram kernel; gram = 0: four columns per pass) against the numpy oracle."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cgmres_oracle as orc
from structurepreservingiterativesolvers_b200 import solvers, wrappers
from structurepreservingiterativesolvers_b200.problems import lkdv
warnings.simplefilter("ignore")
for M in (10_000, 12_000, 15_000):
    dic, _ = lkdv.linforms(space="CG", M=M, mlength=0.8 * M)
    A, b = dic["A"], dic["b"]
    x0 = np.zeros(b.size)
    cl = wrappers.lkdv.conlist(dic, x0)
    tol = 1e-6 * np.sqrt(b.size / 150)
    xr, ir = orc.cgmres(A, b, x0, 50, tol=tol, contol=10, conlist=cl)
    for gram in (1, 0):
        solvers.configure(ctx_options={"gram": gram})
        for eng in ("slsqp", "kkt"):
            x, info = solvers.cgmres(A, b, x0, 50, tol=tol, contol=10, conlist=cl, device=0, small_solver=eng)
            print(f"n={b.size} gram={gram} engine={eng}: steps {info['steps']} (oracle {ir['steps']}) rel.diff {np.linalg.norm(x - xr) / np.linalg.norm(xr):.2e}", flush=True)
