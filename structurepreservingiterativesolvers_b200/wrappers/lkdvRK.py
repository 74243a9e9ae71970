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


def conlist_structured(dic, x0, prob):
    """The same three constraints as CLASS-form quadratics in the stage vector X (device-resident path).

    The callbacks above see the Krylov basis only through z1 = z0 + B X with the stage-weight map
    B = dt [b_1 I ... b_s I] (z1calc, lkdvRK/lkdvRK.py:162-174), and every invariant is a quadratic
    1/2 z1' S z1 + w' z1 + c0 (mass: S = 0, w = omega; momentum: S = M; energy: S = L - M).  Substituting,
        1/2 X' (B'SB) X + (B'(S z0 + w))' X + (1/2 z0'S z0 + w'z0 + c0) = 0,
    i.e. an object with attributes M, v, c that constraint_container reduces on the device (solvers.py:33-36):
    no n x m block of Z is copied to the host per constrained iteration (the price of opaque callbacks)."""
    import scipy.sparse as sps
    from ._common import QuadraticInvariant
    M, L, omega = dic["M"], dic["L"], np.asarray(dic["omega"], dtype=np.float64)
    z0 = np.asarray(dic["z0"], dtype=np.float64)
    wts = prob.dt * np.asarray(prob.butcher_tableau.b, dtype=np.float64)          # B = kron(wts', I)
    outer = sps.csr_matrix(np.outer(wts, wts))

    def lift(S, w, c0, name):
        MM = sps.kron(outer, S, format="csr") if S is not None else sps.csr_matrix((wts.size * z0.size,) * 2)
        g = (S @ z0 if S is not None else 0.0) + w
        vv = np.kron(wts, g)
        cc = (0.5 * z0 @ (S @ z0) if S is not None else 0.0) + w @ z0 + c0
        return QuadraticInvariant(MM, vv, float(cc), name)

    zeros = np.zeros_like(z0)
    return [lift(None, omega, -dic["m0"], "mass"),
            lift(M, zeros, -dic["mo0"], "momentum"),
            lift((L - M).tocsr(), zeros, -dic["e0"], "energy")]


def cgmresWrapper(dic, x0, k, prob=None, pre=None, tol=1e-50, contol=10, structured=False, **ext):
    cl = conlist_structured(dic, x0, prob) if structured else conlist(dic, x0, prob)
    if tol > 1e-20:                                                                          # :82-85
        return solvers.cgmres(A=dic["A"], b=dic["b"], x0=x0, k=k, pre=pre, tol=tol,
                              contol=contol, conlist=cl, **ext)
    return solvers.cgmres_p(A=dic["A"], b=dic["b"], x0=x0, k=k, pre=pre, conlist=cl, **ext)


def gmresWrapper(dic, x0, k, tol=1e-50, pre=None, contol=None, prob=None, **ext):
    return solvers.gmres(A=dic["A"], b=dic["b"], x0=x0, k=k, tol=tol, pre=pre, **ext)


def exact(dic, x0, k=None, tol=None, prob=None, pre=None, contol=None):
    return direct_solve(dic)
