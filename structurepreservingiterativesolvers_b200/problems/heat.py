"""numpy re-assembly of the heat-equation midpoint step of heat/heat.py (CG1 on UnitSquareMesh).

Weak form (heat/heat.py:68-72):  (z - z0)/dt phi + grad((z + z0)/2).grad(phi) = 0, natural BCs:
    A = Mm/dt + L/2,   b = Mm z0/dt - L z0/2.
Returned keys follow heat.linforms (heat/heat.py:105-117): A, b, M, Lz0, old_energy, omega, L,
m0, e0, z0, dt.  Deviation (documented): the initial condition is interpolated at the vertices
instead of L2-projected (heat/heat.py:52); it only changes the data, not the operators.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps


class Problem:
    def __init__(self, N, M, degree, T):
        self.N, self.M, self.degree, self.T = N, M, degree, T
        self.dt = float(T) / N                             # heat/heat.py:21

    @staticmethod
    def ic(x, y):
        return 1e3 * ((x * (x - 1)) ** 5 + (y * (y - 1)) ** 6)   # heat/heat.py:30-32


def _p1_unit_square(M):
    """Mass, stiffness and load vector of P1 elements on the (M x M, right-diagonal) unit square."""
    h = 1.0 / M
    nv = M + 1
    ii, jj = np.meshgrid(np.arange(M), np.arange(M), indexing="ij")
    v00 = (ii * nv + jj).reshape(-1)
    v10 = ((ii + 1) * nv + jj).reshape(-1)
    v01 = (ii * nv + jj + 1).reshape(-1)
    v11 = ((ii + 1) * nv + jj + 1).reshape(-1)
    tris = np.concatenate([np.stack([v00, v10, v11], axis=1), np.stack([v00, v11, v01], axis=1)])
    xs = (np.arange(nv) * h)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    px, py = X.reshape(-1), Y.reshape(-1)
    x = px[tris]
    y = py[tris]
    # gradients of the barycentric basis functions
    bx = np.stack([y[:, 1] - y[:, 2], y[:, 2] - y[:, 0], y[:, 0] - y[:, 1]], axis=1)
    cx = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], axis=1)
    area2 = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
    area = 0.5 * np.abs(area2)
    Ke = (bx[:, :, None] * bx[:, None, :] + cx[:, :, None] * cx[:, None, :]) / (4.0 * area)[:, None, None]
    Me = area[:, None, None] / 12.0 * (np.ones((3, 3)) + np.eye(3))[None, :, :]
    rows = np.repeat(tris[:, :, None], 3, axis=2).reshape(-1)
    cols = np.repeat(tris[:, None, :], 3, axis=1).reshape(-1)
    n = nv * nv
    L = sps.csr_matrix((Ke.reshape(-1), (rows, cols)), shape=(n, n))
    Mm = sps.csr_matrix((Me.reshape(-1), (rows, cols)), shape=(n, n))
    L.sum_duplicates()
    Mm.sum_duplicates()
    omega = np.bincount(tris.reshape(-1), weights=np.repeat(area / 3.0, 3), minlength=n)
    return Mm, L, omega, px, py


def linforms(N=100, M=50, degree=1, T=10, zinit=None):
    if degree != 1:
        raise NotImplementedError("only CG1 is re-assembled")
    prob = Problem(N, M, degree, T)
    dt = prob.dt
    Mm, L, omega, px, py = _p1_unit_square(M)
    z0 = prob.ic(px, py) if zinit is None else np.asarray(zinit, dtype=np.float64)
    A = (Mm / dt + 0.5 * L).tocsr()
    A.sort_indices()
    b = Mm @ z0 / dt - 0.5 * (L @ z0)
    Lz0 = L @ z0                                                    # heat/heat.py:88-89
    old_energy = 0.5 * z0 @ (Mm @ z0) - 0.25 * dt * z0 @ (L @ z0)   # heat/heat.py:92
    out = {"A": A, "b": b, "M": Mm, "Lz0": Lz0, "old_energy": float(old_energy), "omega": omega,
           "L": L, "m0": float(omega @ z0), "e0": 0, "z0": z0, "dt": dt}
    return out, prob


def compute_invariants(params, uvec, uold):
    """mass and energy-dissipation residual (heat/heat.py:123-149) from the assembled forms."""
    uvec, uold = np.asarray(uvec), np.asarray(uold)
    zmid = 0.5 * (uvec + uold)
    Mm, L, dt = params["M"], params["L"], params["dt"]
    return {"mass": float(params["omega"] @ uvec),
            "energy": float(0.5 * uvec @ (Mm @ uvec) - 0.5 * uold @ (Mm @ uold) + dt * zmid @ (L @ zmid))}
