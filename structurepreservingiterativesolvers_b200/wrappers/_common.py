from __future__ import annotations

import numpy as np
import scipy.sparse.linalg as spsla


class QuadraticInvariant:
    """Class-form constraint 1/2 x^T M x + v^T x + c = 0 (the M / v / c attributes that
    constraint_container reads, solvers.py:33-36)."""

    def __init__(self, M, v, c, name=""):
        self.M, self.v, self.c, self.name = M, v, c, name


def direct_solve(dic):
    """The wrappers' `exact` reference solution (lkdv/LinearSolver.py:76-83)."""
    return spsla.spsolve(dic["A"].tocsc(), dic["b"]), -1
