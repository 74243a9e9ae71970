"""numpy re-assembly of the Gauss-Legendre stage system of lkdvRK/lkdvRK.py (no Irksome needed).

Unknowns are the stage derivatives K = [K_1; ...; K_ns], each field-blocked [u; v; w]
(Irksome's stage-derivative formulation, lkdvRK/lkdvRK.py:113-116), with stage values
Z_s = z0 + dt sum_j a_sj K_j and the update z1 = z0 + dt sum_s b_s K_s (z1calc, :162-174):

    A = I_ns (x) blkdiag(Mm, 0, 0) + dt A_GL (x) [[0, G, 0], [-Mm, Mm, -G], [-G, 0, Mm]]
    b_s = -[G v0; Mm (v0 - u0) - G w0; Mm w0 - G u0]

Derived, not cross-checked against Firedrake/Irksome output.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

from . import lkdv as _lkdv


class ButcherTableau:
    def __init__(self, A, b, c):
        self.A, self.b, self.c = np.asarray(A), np.asarray(b), np.asarray(c)
        self.num_stages = len(self.b)


def gauss_legendre(ns):
    """Collocation Runge-Kutta tableau at the Gauss-Legendre nodes (what irk.GaussLegendre builds)."""
    x, _ = np.polynomial.legendre.leggauss(ns)
    c = 0.5 * (x + 1.0)
    V = np.vander(c, ns, increasing=True)                  # V[i,k] = c_i^k
    # sum_j a_ij c_j^k = c_i^(k+1)/(k+1);  sum_j b_j c_j^k = 1/(k+1)
    rhsA = np.stack([c ** (k + 1) / (k + 1) for k in range(ns)], axis=1)
    A = np.linalg.solve(V.T, rhsA.T).T
    b = np.linalg.solve(V.T, 1.0 / np.arange(1, ns + 1))
    return ButcherTableau(A, b, c)


class Problem:
    """Mirror of lkdvRK.problem (lkdvRK/lkdvRK.py:15-31)."""

    def __init__(self, N, M, degree, tstages, space="DG", T=1, mlength=40.0):
        self.mlength, self.degree, self.tstages = mlength, degree, tstages
        self.dim = 3
        self.N, self.M = N, M
        self.dt = float(T) / N
        self.space = space
        self.h = mlength / M
        self.butcher_tableau = gauss_legendre(tstages)
        self.ns = self.butcher_tableau.num_stages
        self.nf = self.dim

    def exact(self, x, t=0.0):
        beta = _lkdv.ALPHA * 2 * np.pi / _lkdv.REF_LENGTH
        return np.sin(beta * (x - (1 - beta ** 2) * t)) + 1


def linforms(N=100, M=50, degree=1, tstages=2, T=10, zinit=None, space="DG", mlength=None):
    """`mlength=None` keeps the reference's domain length 40 (lkdvRK/lkdvRK.py:17); pass mlength=0.8*M to hold the
    mesh width fixed when scaling the stage system up."""
    if degree != 1:
        raise NotImplementedError("only degree 1 is re-assembled")
    prob = Problem(N=N, M=M, T=T, degree=degree, tstages=tstages, space=space,
                   mlength=(40.0 if mlength is None else float(mlength)))
    h, dt = prob.h, prob.dt
    if space == "DG":
        Mm, G, omega_u = _lkdv._field_matrices_dg1(M, h)
        nd = 2 * M
    else:
        Mm, G, omega_u = _lkdv._field_matrices_cg(M, h)
        nd = M
    if zinit is None:
        u0 = _lkdv._project_dg1(prob, M, h) if space == "DG" else prob.exact(np.arange(M) * h)
    else:
        u0 = np.asarray(zinit, dtype=np.float64)[:nd].copy()
    w0 = _lkdv._solve_mass(Mm, G @ u0, space, h) if zinit is None else np.asarray(zinit)[2 * nd:].copy()
    v0 = u0 + _lkdv._solve_mass(Mm, G @ w0, space, h)      # v_finder, lkdvRK/lkdvRK.py:64-72
    zero = sps.csr_matrix((nd, nd))
    J0 = sps.block_diag([Mm, zero, zero], format="csr")
    J1 = sps.bmat([[None, G, None], [-Mm, Mm, -G], [-G, None, Mm]], format="csr")
    bt = prob.butcher_tableau
    ns = prob.ns
    A = (sps.kron(sps.identity(ns), J0) + dt * sps.kron(sps.csr_matrix(bt.A), J1)).tocsr()
    A.sort_indices()
    b_stage = -np.concatenate([G @ v0, Mm @ (v0 - u0) - G @ w0, Mm @ w0 - G @ u0])
    b = np.tile(b_stage, ns)
    Mmat = sps.block_diag([Mm, zero, zero], format="csr")
    Lmat = sps.block_diag([zero, zero, Mm], format="csr")
    omega = np.concatenate([omega_u, np.zeros(nd), np.zeros(nd)])
    z0 = np.concatenate([u0, v0, w0])
    out = {"A": A, "b": b, "M": Mmat, "L": Lmat, "omega": omega,
           "m0": float(omega_u @ u0), "mo0": float(0.5 * u0 @ (Mm @ u0)),
           "e0": float(0.5 * (w0 @ (Mm @ w0) - u0 @ (Mm @ u0))), "T": T, "z0": z0}
    return out, prob


def z1calc(prob, zbig, z0):
    """RK update z1 = z0 + dt sum_s b_s K_s from the stacked stage vector (lkdvRK/lkdvRK.py:162-174)."""
    dof = len(z0)
    K = np.asarray(zbig).reshape(prob.ns, dof)
    return np.asarray(z0, dtype=np.float64) + prob.dt * (prob.butcher_tableau.b @ K)


def dz1calc(prob, Q, z0):
    """d z1 / d y for zbig = x0 + Q y (lkdvRK/lkdvRK.py:178-189): dt sum_s b_s Q[s-block, :]."""
    dof = len(z0)
    Q = np.asarray(Q)
    out = np.zeros((dof, Q.shape[-1]))
    for s in range(prob.ns):
        out += prob.dt * prob.butcher_tableau.b[s] * Q[s * dof:(s + 1) * dof, :]
    return out
