"""heat/LinearSolver.py mirror: mass + energy dissipation (:26-38); the energy constraint has a
non-zero linear term v = dt/2 L z0 and M + dt/2 L as its quadratic form."""
from __future__ import annotations

import numpy as np

from .. import solvers
from ._common import QuadraticInvariant


def conlist(dic, x0):
    dt = dic["dt"]
    return [QuadraticInvariant(0 * dic["A"], np.transpose(dic["omega"]), -dic["m0"], "mass"),
            QuadraticInvariant(dic["M"] + 0.5 * dt * dic["L"], 0.5 * dt * dic["Lz0"],
                               -dic["old_energy"], "energy")]


def cgmresWrapper(dic, x0, k, tol=1e-50, pre=None, timing=None, **ext):
    cl = conlist(dic, x0)
    if tol < 1e-20:                                                                          # :43-46
        return solvers.cgmres_p(A=dic["A"], b=dic["b"], x0=x0, k=k, conlist=cl, pre=pre, **ext)
    return solvers.cgmres(A=dic["A"], b=dic["b"], x0=x0, k=k, tol=tol, conlist=cl, pre=pre,
                          timing=timing, **ext)


def gmresWrapper(dic, x0, k, tol=1e-50, pre=None, **ext):
    return solvers.gmres(A=dic["A"], b=dic["b"], x0=x0, k=k, tol=tol, pre=pre, **ext)
