"""lkdvRK/LinearSolver.py mirror: three DICT-form (opaque callback) constraints on the
Runge-Kutta reconstruction z1 = z0 + dt sum_s b_s X_s of the stage vector X = x0 + Q y (:29-79)."""
from __future__ import annotations

import numpy as np

from .. import solvers
from ..problems import lkdvRK as _rk
from ._common import direct_solve


def conlist(dic, x0, prob):
    M, L, omega = dic["M"], dic["L"], dic["omega"]
    m0, mo0, e0, z0 = dic["m0"], dic["mo0"], dic["e0"], dic["z0"]

    def z1_of(y, x0_, Q):
        return _rk.z1calc(prob, x0_ + Q @ y, z0)

    def mass(y, x0_, Q):
        return omega @ z1_of(y, x0_, Q) - m0

    def momentum(y, x0_, Q):
        X = z1_of(y, x0_, Q)
        return 0.5 * X @ (M @ X) - mo0

    def energy(y, x0_, Q):
        X = z1_of(y, x0_, Q)
        return 0.5 * X @ (L @ X) - 0.5 * X @ (M @ X) - e0

    def d_mass(y, x0_, Q):
        return omega @ _rk.dz1calc(prob, Q, z0)

    def d_momentum(y, x0_, Q):
        return z1_of(y, x0_, Q) @ (M @ _rk.dz1calc(prob, Q, z0))

    def d_energy(y, x0_, Q):
        X = z1_of(y, x0_, Q)
        dX = _rk.dz1calc(prob, Q, z0)
        return X @ (L @ dX) - X @ (M @ dX)

    return [{"func": mass, "jac": d_mass},
            {"func": momentum, "jac": d_momentum},
            {"func": energy, "jac": d_energy}]


def cgmresWrapper(dic, x0, k, prob=None, pre=None, tol=1e-50, contol=10, **ext):
    cl = conlist(dic, x0, prob)
    if tol > 1e-20:                                                                          # :82-85
        return solvers.cgmres(A=dic["A"], b=dic["b"], x0=x0, k=k, pre=pre, tol=tol,
                              contol=contol, conlist=cl, **ext)
    return solvers.cgmres_p(A=dic["A"], b=dic["b"], x0=x0, k=k, pre=pre, conlist=cl, **ext)


def gmresWrapper(dic, x0, k, tol=1e-50, pre=None, contol=None, prob=None, **ext):
    return solvers.gmres(A=dic["A"], b=dic["b"], x0=x0, k=k, tol=tol, pre=pre, **ext)


def exact(dic, x0, k=None, tol=None, prob=None, pre=None, contol=None):
    return direct_solve(dic)
