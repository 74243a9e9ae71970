"""CPU oracle: numpy/scipy restatement of the reference's Krylov solvers.  TEST INFRASTRUCTURE ONLY.

This file restates /root/reference/solvers.py (FGMRES `gmres` :58-127, `cgmres` :131-323,
prototypical `cgmres_p` :328-445, `constraint_container` :21-53, `constraint_checker` :14-18) so
that parity tests and the CPU baseline can run where /root/reference does not exist (the GPU
box).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import it; the product package never does.

Pinning: tests/golden/*.npz hold outputs of the UNMODIFIED reference (imported from
/root/reference with stub firedrake/matplotlib modules by tests/golden/make_golden.py, numpy
2.3.5 / scipy 1.18.1); tests/test_oracle_golden.py checks this restatement against them.

Third-party arithmetic the reference delegates to (not vendored there): numpy dot/norm/lstsq,
scipy.sparse CSR products, scipy.optimize SLSQP.  The same library calls are made here, in the
same order, so results agree with the reference to round-off.

Behavioural quirks kept on purpose:
  Q1 history[0] is r0;  Q2 `res` omits the initial residual (gmres/cgmres);  Q3 breakdown
  breaks before the iterate is updated;  Q4 a violated constraint lands in the unconstrained
  fallback (the reference touches a missing attribute inside its try block);  Q5 once the
  constrained phase has started it never reverts;  Q6 timing mode skips the NaN / violation
  checks;  Q7 the invariant is 1/2 x^T M x + v^T x + c;  Q8 modified Gram-Schmidt, one pass.
"""
from __future__ import annotations

import warnings
from time import time

import numpy as np
import scipy.optimize as spo
import scipy.sparse as sps

_OK_MESSAGES = ("Optimization terminated successfully",
                "`xtol` termination condition is satisfied.",
                "`gtol` termination condition is satisfied.")
_BREAKDOWN = ("GMRES broke down, either initial guess is exact or , more likely, "
              "something has gone wrong.")


def max_signed_violation(y, scipy_constraints):
    """solvers.py:14-18 (no abs: only positive violations count)."""
    worst = 0
    for con in scipy_constraints:
        worst = max(worst, con["fun"](y))
    return worst


class ReducedInvariant:
    """solvers.py:21-53.  Class form -> term0/term1/term2; dict form -> stored callbacks."""

    def __init__(self, const, x0, Z):
        if hasattr(const, "__dict__"):
            self.quadratic = True
        elif type(const) is dict:
            self.quadratic = False
        else:
            raise NotImplementedError("Constraints must be either dictionaries or classes")
        if self.quadratic:
            self.MZ = const.M @ Z                                       # :33
            self.term0 = 0.5 * x0 @ const.M @ x0 + const.c + const.v @ x0   # :34
            self.term1 = const.v @ Z + x0 @ self.MZ                     # :35
            self.term2 = 0.5 * Z.T @ self.MZ                            # :36
        else:
            self.const, self.x0, self.Z = const, x0, Z

    def fun(self, y):
        if self.quadratic:
            return self.term0 + self.term1 @ y + y @ self.term2 @ y     # :44
        return self.const["func"](y, self.x0, self.Z)

    def jac(self, y):
        if self.quadratic:
            return self.term1 + 2 * y @ self.term2                       # :50
        return self.const["jac"](y, self.x0, self.Z)

    def as_scipy(self):
        return {"type": "eq", "fun": self.fun, "jac": self.jac}


def _prefunc(pre, n):
    """solvers.py:149-161."""
    if pre is None:
        pre = sps.identity(n)
    if hasattr(pre, "solve"):
        return lambda vec: pre.solve(vec)

    def apply(vec):
        try:
            return pre @ vec
        except Exception:
            raise ValueError("Preconditioner not supported")
    return apply


class _Krylov:
    """State shared by the three solvers: q, z, h and the modified Gram-Schmidt step."""

    def __init__(self, A, b, x0, k, pre):
        self.A, self.b, self.x0, self.k = A, b, x0, k
        self.pre = _prefunc(pre, len(b))
        self.r0 = b - A.dot(x0)                                          # :167
        n = np.size(self.r0)
        self.q = np.zeros((k + 1, n))
        self.z = np.zeros((k + 1, n))
        self.h = np.zeros((k + 1, k))
        self.beta = np.linalg.norm(self.r0)
        self.q[0] = self.r0 / self.beta                                  # :177

    def step(self, j):
        """solvers.py:190-198; returns False on breakdown (h[j+1,j] == 0)."""
        q, z, h = self.q, self.z, self.h
        z[j] = np.asarray(self.pre(q[j]))
        y = np.asarray(self.A @ z[j])
        for i in range(j + 1):
            h[i, j] = np.dot(q[i], y)
            y = y - h[i, j] * q[i]
        h[j + 1, j] = np.linalg.norm(y)
        if h[j + 1, j] != 0:
            q[j + 1] = y / h[j + 1, j]
            return True
        return False

    def small_problem(self, j):
        """Objective and gradient of |beta e1 - H y|^2 (solvers.py:204-219)."""
        rhs = np.zeros(j + 2)
        rhs[0] = self.beta
        Hj = self.h[: j + 2, : j + 1]

        def func(y):
            F = rhs - Hj @ y
            return np.inner(F, F)

        def jac(y):
            F = rhs - Hj @ y
            return -2 * np.transpose(Hj) @ F

        return rhs, Hj, func, jac

    def Z(self, j):
        return np.transpose(self.z[: j + 1, :])                          # :207

    def iterate(self, j, y):
        return self.Z(j) @ y + self.x0                                   # :287

    def true_residual(self, x):
        return np.linalg.norm(self.A.dot(x) - self.b)                    # :290


def _slsqp(func, jac, y0, constraints, ftol, tol=None):
    return spo.minimize(func, y0, tol=tol, jac=jac, constraints=constraints, method="SLSQP",
                        options={"ftol": ftol, "maxiter": 1e3})


def _complain(j, result):
    if result.message not in _OK_MESSAGES:                               # :280-284
        warnings.warn("Iteration %d failed with message '%s'" % (j, result.message), RuntimeWarning)


def fgmres(A, b, x0, k, tol=1e-50, pre=None):
    """solvers.py:58-127."""
    kr = _Krylov(A, b, x0, k, pre)
    xs = [kr.r0]
    res = [kr.beta]
    steps = 0
    for j in range(k):
        steps = j + 1
        if not kr.step(j):
            warnings.warn(_BREAKDOWN)
            break
        rhs, Hj, _, _ = kr.small_problem(j)
        yk = np.linalg.lstsq(Hj, rhs, rcond=None)[0]                     # :113
        xs.append(kr.iterate(j, yk))
        res.append(kr.true_residual(xs[-1]))
        if res[-1] < tol:
            break
    return xs[-1], {"name": "gmres", "x": xs, "res": res[1:], "steps": steps}


def cgmres(A, b, x0, k, tol=1e-8, contol=10, conlist=(), pre=None, timing=None, _record=None):
    """solvers.py:131-323.  `_record`, if a dict, receives the Hessenberg matrix and the y history."""
    ctol = 1e-12
    if timing:
        marks = {"start": time(), "start_iter": [], "end_iter": [], "start_con": [], "end_con": []}
    kr = _Krylov(A, b, x0, k, pre)
    safety = None
    xs = [kr.r0]
    res = [kr.beta]
    constrained_steps = 0
    steps = 0
    yk = None
    ys = []
    for j in range(k):
        if timing:
            marks["start_iter"].append(time())
        steps = j + 1
        if not kr.step(j):
            warnings.warn(_BREAKDOWN)
            break
        rhs, Hj, func, jac = kr.small_problem(j)
        Z = kr.Z(j)
        y0 = np.zeros(j + 1)
        if j != 0:
            y0[:-1] = yk
        if res[-1] > contol * tol and j < k - 1 and safety is None:     # :230
            sol = _slsqp(func, jac, y0, [], ctol ** 2)
        else:
            try:
                if timing:
                    constrained_steps += 1
                    marks["start_con"].append(time())
                clist = [ReducedInvariant(c, x0, Z).as_scipy() for c in conlist]     # :242-247
                if timing:
                    marks["end_con"].append(time())
                sol = _slsqp(func, jac, y0, clist[:], ctol ** 2)                     # :251-255
                if not timing and np.isnan(max(sol.x)):
                    raise ValueError
                safety = True
                if not timing and max_signed_violation(sol.x, clist) > ctol:
                    safety = False
                    # the reference formats `solve.constr_violation` here, which SLSQP results do
                    # not carry -> AttributeError -> the bare except below (quirk Q4)
                    raise AttributeError("constr_violation")
            except Exception:
                warnings.warn("Constrained solve failed, defaulted to standard solve for iteration %d."
                              " Problem likely overconstrained, a smaller solver tolerance may be "
                              "required." % j, RuntimeWarning)
                if timing and len(marks["end_con"]) < len(marks["start_con"]):
                    marks["end_con"].append(time())
                sol = _slsqp(func, jac, y0, [], ctol ** 2)                           # :274-278
        _complain(j, sol)
        yk = sol.x
        ys.append(np.array(yk))
        xs.append(kr.iterate(j, yk))
        res.append(kr.true_residual(xs[-1]))
        if timing:
            marks["end_iter"].append(time())
        if res[-1] < tol and safety is True:                             # :296
            break
    timings = None
    if timing:                                                           # :300-312
        marks["end"] = time()
        it = np.asarray(marks["end_iter"]) - np.asarray(marks["start_iter"][: len(marks["end_iter"])])
        unc = it[:-constrained_steps]
        assembly = np.asarray(marks["end_con"]) - np.asarray(marks["start_con"])
        con = it[len(unc):] - assembly
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            timings = {"runtime": marks["end"] - marks["start"],
                       "iter_time_unconstrained": np.mean(unc),
                       "iter_time_constrained": np.mean(con),
                       "constraint_building": np.mean(assembly),
                       "constrained_steps": constrained_steps}
    if _record is not None:
        _record.update(H=kr.h.copy(), beta=kr.beta, ys=ys, q=kr.q, z=kr.z)
    return xs[-1], {"name": "cgmres", "x": xs, "res": res[1:], "steps": steps, "timings": timings}


def cgmres_prototype(A, b, x0, k, conlist=(), pre=None, _record=None):
    """solvers.py:328-445: constraints clist[:j] from iteration j on; always k iterations."""
    kr = _Krylov(A, b, x0, k, pre)
    xs = [kr.r0]
    res = []
    yk = None
    ys = []
    for j in range(k):
        kr.step(j)                                                       # no break (:376-377)
        rhs, Hj, func, jac = kr.small_problem(j)
        Z = kr.Z(j)
        clist = [ReducedInvariant(c, x0, Z).as_scipy() for c in conlist]             # :397-401
        y0 = np.zeros(j + 1)
        if j != 0:
            y0[:-1] = yk
        sol = _slsqp(func, jac, y0, clist[:j], 1e-20, tol=1e-15)                     # :411-415
        if np.isnan(max(sol.x)):
            warnings.warn("Constrained solve silently failed on iteration %d" % j)
            sol = _slsqp(func, jac, y0, [], 1e-20)                                   # :420-424
        _complain(j, sol)
        yk = sol.x
        ys.append(np.array(yk))
        xs.append(kr.iterate(j, yk))
        res.append(kr.true_residual(xs[-1]))
    if _record is not None:
        _record.update(H=kr.h.copy(), beta=kr.beta, ys=ys, q=kr.q, z=kr.z)
    return xs[-1], {"name": "geosolve", "x": xs, "res": res}
