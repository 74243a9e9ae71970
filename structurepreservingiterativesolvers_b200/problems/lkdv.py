"""numpy re-assembly of the linear-KdV midpoint step of lkdv/lkdv.py (no Firedrake needed).

`linforms` returns the same dictionary keys as the reference's `lkdv.linforms`
(lkdv/lkdv.py:135-146): A, b, z0, M, L, omega, m0, mo0, e0, T.  Unknown ordering is
field-blocked [u; v; w] (lkdv/refd.py:17).

Weak form (lkdv/lkdv.py:96-105), with Mm the per-field mass matrix and
G[test, trial] = int trial_x test dx - sum_facets jump(trial, n) avg(test)   (lkdv.py:59-61):

    Mm u/dt + G v               = Mm u0/dt
    -1/2 Mm u + Mm v - 1/2 G w  = 1/2 Mm u0 + 1/2 G w0
    -G u           + Mm w       = 0

space='CG' : periodic P1, Mm = h/6 circ[1,4,1], G = 1/2 circ[-1,0,1] (facet terms vanish).
space='DG' : periodic DG1, Mm = blockdiag(h/6 [[2,1],[1,2]]), G = cell term + central facet flux.
The DG1/M=50 case is the reference's SingleSolve default (n = 300); the CG case
with h = 0.8 held fixed is the 1e7-DOF benchmark system (cfg2).  This is a derived restatement:
Firedrake is not installable here, so entries are not cross-checked against PETSc output.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

ALPHA = 4.0
REF_LENGTH = 40.0            # lkdv/lkdv.py:17


class Problem:
    """Mirror of lkdv.problem (lkdv/lkdv.py:15-37): parameters only."""

    def __init__(self, N, M, degree, space, mlength, T):
        self.N, self.M, self.degree, self.space = N, M, degree, space
        self.mlength = mlength
        self.dim = 3
        self.T = T
        self.dt = float(T) / N
        self.h = mlength / M

    def exact(self, x, t=0.0):
        beta = ALPHA * 2 * np.pi / REF_LENGTH            # lkdv/lkdv.py:33-36
        return np.sin(beta * (x - (1 - beta ** 2) * t)) + 1


def _cyclic_shift(M, dtype=np.float64):
    """(S u)_i = u_{i+1 mod M}"""
    idx = np.arange(M)
    return sps.csr_matrix((np.ones(M, dtype=dtype), (idx, (idx + 1) % M)), shape=(M, M))


def _field_matrices_cg(M, h):
    S = _cyclic_shift(M)
    Mm = (h / 6.0) * (4.0 * sps.identity(M, format="csr") + S + S.T)
    G = 0.5 * (S - S.T)
    omega_u = np.full(M, h)
    return Mm.tocsr(), G.tocsr(), omega_u


def _field_matrices_dg1(M, h):
    nd = 2 * M
    cell = np.arange(M)
    l, r = 2 * cell, 2 * cell + 1                         # left / right dof of each cell
    # mass
    rows = np.concatenate([l, l, r, r])
    cols = np.concatenate([l, r, l, r])
    vals = np.concatenate([np.full(M, 2.0), np.full(M, 1.0), np.full(M, 1.0), np.full(M, 2.0)]) * h / 6.0
    Mm = sps.csr_matrix((vals, (rows, cols)), shape=(nd, nd))
    # cell part of G: rows (test) l and r both get 1/2 (u_r - u_l)
    rows = np.concatenate([l, l, r, r])
    cols = np.concatenate([l, r, l, r])
    vals = np.concatenate([np.full(M, -0.5), np.full(M, 0.5), np.full(M, -0.5), np.full(M, 0.5)])
    Gc = sps.csr_matrix((vals, (rows, cols)), shape=(nd, nd))
    # facet between cell c (left, dof L = r[c]) and cell c+1 (right, dof R = l[c+1]):
    # -jump(u,n) avg(v) = (u_R - u_L) * 1/2 (v_L + v_R)
    L = r
    R = l[(cell + 1) % M]
    rows = np.concatenate([L, L, R, R])
    cols = np.concatenate([R, L, R, L])
    vals = np.concatenate([np.full(M, 0.5), np.full(M, -0.5), np.full(M, 0.5), np.full(M, -0.5)])
    Gf = sps.csr_matrix((vals, (rows, cols)), shape=(nd, nd))
    G = (Gc + Gf).tocsr()
    G.sum_duplicates()
    omega_u = np.full(nd, h / 2.0)
    return Mm, G, omega_u


def _project_dg1(prob, M, h):
    """L2 projection of the initial condition onto DG1 (lkdv/lkdv.py:78), 8-point Gauss rule."""
    xg, wg = np.polynomial.legendre.leggauss(8)
    xi = 0.5 * (xg + 1.0)                                 # reference coordinates in (0,1)
    w = 0.5 * wg
    x = (np.arange(M)[:, None] + xi[None, :]) * h
    f = prob.exact(x)
    b0 = (f * (1 - xi)[None, :] * w[None, :]).sum(axis=1) * h
    b1 = (f * xi[None, :] * w[None, :]).sum(axis=1) * h
    # invert h/6 [[2,1],[1,2]] cell by cell
    det = (h / 6.0) ** 2 * 3.0
    u_l = (h / 6.0) * (2 * b0 - b1) / det
    u_r = (h / 6.0) * (2 * b1 - b0) / det
    return np.stack([u_l, u_r], axis=1).reshape(-1)


def _solve_mass(Mm, rhs, space, h):
    if space == "DG":
        a = rhs.reshape(-1, 2)
        det = (h / 6.0) ** 2 * 3.0
        out = np.empty_like(a)
        out[:, 0] = (h / 6.0) * (2 * a[:, 0] - a[:, 1]) / det
        out[:, 1] = (h / 6.0) * (2 * a[:, 1] - a[:, 0]) / det
        return out.reshape(-1)
    n = rhs.size
    if n <= 200_000:
        return spsla.spsolve(Mm.tocsc(), rhs)
    # periodic tridiagonal h/6 circ[1,4,1]: diagonally dominant, a few dozen Jacobi-preconditioned
    # CG sweeps reach round-off (condition number 3) without a factorisation of a 3e6 system
    x = rhs / (4.0 * h / 6.0)
    r = rhs - Mm @ x
    p = r.copy()
    rs = r @ r
    for _ in range(200):
        Ap = Mm @ p
        a = rs / (p @ Ap)
        x += a * p
        r -= a * Ap
        rs_new = r @ r
        if rs_new <= 1e-32 * (rhs @ rhs):
            break
        p = r + (rs_new / rs) * p
        rs = rs_new
    return x


def linforms(N=100, M=50, degree=1, T=1, space="DG", zinit=None, mlength=None):
    """Same keys as lkdv.linforms (lkdv/lkdv.py:46-148).  `mlength=None` keeps the reference's
    domain length 40; pass mlength=0.8*M to hold h fixed when scaling up."""
    if degree != 1:
        raise NotImplementedError("only degree 1 (P1 / DG1) is re-assembled")
    if mlength is None:
        mlength = REF_LENGTH
    prob = Problem(N, M, degree, space, float(mlength), T)
    h, dt = prob.h, prob.dt
    if space == "CG":
        Mm, G, omega_u = _field_matrices_cg(M, h)
        nd = M
    elif space == "DG":
        Mm, G, omega_u = _field_matrices_dg1(M, h)
        nd = 2 * M
    else:
        raise ValueError("space must be 'CG' or 'DG'")
    if zinit is None:
        u0 = prob.exact(np.arange(M) * h) if space == "CG" else _project_dg1(prob, M, h)
        w0 = _solve_mass(Mm, G @ u0, space, h)            # gfuncproject, lkdv/lkdv.py:62-69,79
        v0 = np.zeros(nd)
    else:
        zinit = np.asarray(zinit, dtype=np.float64)
        u0, v0, w0 = zinit[:nd].copy(), np.zeros(nd), zinit[2 * nd:].copy()
    Z = None
    A = sps.bmat([[Mm / dt, G, Z],
                  [-0.5 * Mm, Mm, -0.5 * G],
                  [-G, Z, Mm]], format="csr")
    A.sort_indices()
    b = np.concatenate([Mm @ u0 / dt, 0.5 * (Mm @ u0) + 0.5 * (G @ w0), np.zeros(nd)])
    zero = sps.csr_matrix((nd, nd))
    Mmat = sps.block_diag([Mm, zero, zero], format="csr")          # lkdv/lkdv.py:114-116
    Lmat = sps.block_diag([zero, zero, Mm], format="csr")          # lkdv/lkdv.py:118-120
    omega = np.concatenate([omega_u, np.zeros(nd), np.zeros(nd)])  # lkdv/lkdv.py:122
    z0 = np.concatenate([u0, v0, w0])
    m0 = float(omega_u @ u0)                                       # lkdv/lkdv.py:125-127
    mo0 = float(0.5 * u0 @ (Mm @ u0))
    e0 = float(0.5 * w0 @ (Mm @ w0) - mo0)
    out = {"A": A, "b": b, "z0": z0, "M": Mmat, "L": Lmat, "omega": omega,
           "m0": m0, "mo0": mo0, "e0": e0, "T": T}
    return out, prob


def compute_invariants(params, uvec):
    """mass, momentum, energy of a solution vector (lkdv/lkdv.py:154-166) from the assembled forms."""
    uvec = np.asarray(uvec)
    return {"mass": float(params["omega"] @ uvec),
            "momentum": float(0.5 * uvec @ (params["M"] @ uvec)),
            "energy": float(0.5 * uvec @ (params["L"] @ uvec) - 0.5 * uvec @ (params["M"] @ uvec))}


def benchmark_size(n_target=10_000_000):
    """Elements per field for the cfg2 benchmark: multiple of 25 so that sin(beta x) is periodic
    on a domain of length 0.8*M (wavelength 10 = 12.5 h)."""
    M = int(np.ceil(n_target / 3 / 25.0)) * 25
    return M
