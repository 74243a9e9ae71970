"""lkdv/LinearSolver.py mirror: mass, momentum, energy (3 class-form constraints, :28-47)."""
from __future__ import annotations

import warnings
import numpy as np

from .. import solvers
from ._common import QuadraticInvariant, direct_solve


def conlist(dic, x0):
    A, M, L = dic["A"], dic["M"], dic["L"]
    zeros = np.zeros_like(x0)
    return [QuadraticInvariant(0 * A, np.transpose(dic["omega"]), -dic["m0"], "mass"),      # :28-32
            QuadraticInvariant(M, zeros, -dic["mo0"], "momentum"),                          # :34-38
            QuadraticInvariant(L - M, zeros, -dic["e0"], "energy")]                         # :40-44


def cgmresWrapper(dic, x0, k, tol=1e-50, contol=10, timing=None, **ext):
    cl = conlist(dic, x0)
    if tol > 1e-20:                                                                          # :50-52
        return solvers.cgmres(A=dic["A"], b=dic["b"], x0=x0, k=k, tol=tol, contol=contol,
                              conlist=cl, timing=timing, **ext)
    if timing is not None:
        raise NotImplementedError("Timings are not available for prototypical solver")
    return solvers.cgmres_p(A=dic["A"], b=dic["b"], x0=x0, k=k, conlist=cl, **ext)           # :58-59


def gmresWrapper(dic, x0, k, tol=1e-50, contol=None, **ext):
    if contol is not None:
        warnings.warn("Contol is ignored as not used in GMRES")
    return solvers.gmres(A=dic["A"], b=dic["b"], x0=x0, k=k, tol=tol, **ext)


def exact(dic, x0=None, k=None, tol=None, prob=None, contol=None):
    return direct_solve(dic)
