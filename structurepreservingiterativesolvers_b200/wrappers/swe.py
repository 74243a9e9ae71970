"""swe/LinearSolver.py mirror: mass + energy (2 class-form constraints, :23-36), optional `pre`."""
from __future__ import annotations

import numpy as np

from .. import solvers
from ._common import QuadraticInvariant


def conlist(dic, x0):
    return [QuadraticInvariant(0 * dic["A"], np.transpose(dic["omega"]), -dic["m0"], "mass"),
            QuadraticInvariant(dic["L"], np.zeros_like(x0), -dic["e0"], "energy")]


def cgmresWrapper(dic, x0, k, tol=1e-50, pre=None, timing=None, **ext):
    cl = conlist(dic, x0)
    if tol < 1e-20:                                                                          # :40-43
        return solvers.cgmres_p(A=dic["A"], b=dic["b"], x0=x0, k=k, conlist=cl, pre=pre, **ext)
    return solvers.cgmres(A=dic["A"], b=dic["b"], x0=x0, k=k, tol=tol, conlist=cl,
                          timing=timing, pre=pre, **ext)


def gmresWrapper(dic, x0, k, tol=1e-50, pre=None, **ext):
    return solvers.gmres(A=dic["A"], b=dic["b"], x0=x0, k=k, tol=tol, pre=pre, **ext)
