"""Host side of the Krylov loop: the k-dimensional (constrained) least-squares problem.

north_star keeps exactly this on the host: "only the k-dimensional constrained minimisation
over Krylov coefficients and the Givens least-squares update stay on the host".

Two interchangeable engines minimise  f(y) = || beta e1 - H y ||^2  subject to the reduced
quadratic constraints  g_c(y) = term0 + term1.y + y^T term2 y = 0:

* ``slsqp``  -- scipy.optimize.minimize(method='SLSQP') with the reference's exact options
                (solvers.py:231-235, 251-255, 274-278, 411-415).  Parity mode.
* ``kkt``    -- QR factorisation of H (never normal equations) + Newton iteration on the
                Lagrange conditions in the variable u = R y - Q^T(beta e1).  Fast mode: it
                costs ~0.1 ms where SLSQP with ftol=1e-24 regularly runs to maxiter=1000.
"""
from __future__ import annotations

import ctypes as _C

import numpy as np
import scipy.optimize as spo

from . import _native as _nat

SUCCESS_MESSAGE = "Optimization terminated successfully"
_QUIET_MESSAGES = (SUCCESS_MESSAGE,
                   "`xtol` termination condition is satisfied.",
                   "`gtol` termination condition is satisfied.")


class SmallResult:
    """The fields of scipy's OptimizeResult that solvers.py reads (x, message)."""

    __slots__ = ("x", "message", "success", "nit", "fun")

    def __init__(self, x, message=SUCCESS_MESSAGE, success=True, nit=0, fun=None):
        self.x, self.message, self.success, self.nit, self.fun = x, message, success, nit, fun


def message_is_quiet(message: str) -> bool:
    """solvers.py:280-284 warns unless the message is one of three strings."""
    return message in _QUIET_MESSAGES


class ReducedConstraint:
    """Reduced quadratic invariant in Krylov coordinates (solvers.py:42-53).

    fun(y) = term0 + term1.y + y.term2.y ; jac(y) = term1 + 2 y.term2   (class form), or the
    caller's opaque callbacks evaluated with the host copy of Z (dict form).
    """

    def __init__(self, term0=None, term1=None, term2=None, callbacks=None, x0=None, Z=None):
        self.term0, self.term1, self.term2 = term0, term1, term2
        self.callbacks, self.x0, self.Z = callbacks, x0, Z
        self.quadratic = callbacks is None

    def fun(self, y):
        if self.quadratic:
            return self.term0 + self.term1 @ y + y @ self.term2 @ y
        return self.callbacks["func"](y, self.x0, self.Z)

    def jac(self, y):
        if self.quadratic:
            return self.term1 + 2 * y @ self.term2
        return self.callbacks["jac"](y, self.x0, self.Z)

    def as_scipy(self):
        return {"type": "eq", "fun": self.fun, "jac": self.jac}


def _objective(Hj, beta):
    rhs = np.zeros(Hj.shape[0])
    rhs[0] = beta

    def func(y):
        F = rhs - Hj @ y
        return np.inner(F, F)

    def jac(y):
        F = rhs - Hj @ y
        return -2.0 * (Hj.T @ F)

    return func, jac


def slsqp(Hj, beta, y0, constraints=(), ftol=1e-24, tol=None):
    """The reference's small solve, verbatim options (solvers.py:231-235 / 411-415)."""
    func, jac = _objective(Hj, beta)
    return spo.minimize(func, y0, tol=tol, jac=jac,
                        constraints=[c.as_scipy() for c in constraints],
                        method="SLSQP", options={"ftol": ftol, "maxiter": 1e3})


def lstsq(Hj, beta):
    """Unconstrained minimiser by QR of the (m+1) x m Hessenberg matrix (solvers.py:113)."""
    rhs = np.zeros(Hj.shape[0])
    rhs[0] = beta
    y = np.linalg.lstsq(Hj, rhs, rcond=None)[0]
    return SmallResult(y)


def kkt(Hj, beta, y0, constraints=(), max_newton=40):
    """Equality-constrained least squares by Newton on the KKT system.

    With H = Q R and c = Q^T(beta e1) the objective is |R y - c|^2 + const.  In the variable
    u = R y - c the objective Hessian is 2 I, so conditioning enters only through the
    constraint curvature S^T (T2 + T2^T) S with S = R^{-1}.  Opaque (dict-form) constraints
    are handled through their callbacks with a Gauss-Newton Hessian (their curvature is not
    available), which still converges because they are quadratics of tiny curvature * lambda.
    """
    m = Hj.shape[1]
    native = _kkt_native(Hj, beta, constraints)
    if native is not None:
        return native
    rhs = np.zeros(Hj.shape[0])
    rhs[0] = beta
    Q, R = np.linalg.qr(Hj)                       # reduced: Q (m+1, m), R (m, m)
    c = Q.T @ rhs
    diag = np.abs(np.diag(R))
    if m == 0 or diag.min() <= 1e-300:
        return lstsq(Hj, beta)
    cons = list(constraints)
    # (LAPACK getrf/getrs on the triangular R: no row is ever swapped, i.e. back substitution -- but without the
    #  threaded BLAS-3 trsm behind scipy's solve_triangular, whose worker threads have to be woken up for a
    #  20 x 20 system: 350 us per call measured inside this function, 30 us for this)
    y_ls = np.linalg.solve(R, c)
    if not cons:
        return SmallResult(y_ls)
    nc = len(cons)
    S = np.linalg.solve(R, np.eye(m))             # R^{-1}
    curv = []                                      # S^T (T2 + T2^T) S per quadratic constraint
    for con in cons:
        if con.quadratic:
            T = con.term2 + con.term2.T
            curv.append(S.T @ T @ S)
        else:
            curv.append(None)
    def evaluate(u_):
        y_ = S @ (c + u_)
        g_ = np.array([con.fun(y_) for con in cons], dtype=float)
        Jy = np.array([np.asarray(con.jac(y_), dtype=float).reshape(-1) for con in cons])
        return y_, g_, Jy @ S

    def newton(u):
        lam = np.zeros(nc)
        y, g, Ju = evaluate(u)
        scale = np.array([max(abs(con.term0), 1e-300) if con.quadratic else max(abs(g_i), 1.0)
                          for con, g_i in zip(cons, g)])
        converged = False
        nit = 0
        for nit in range(1, max_newton + 1):
            Hl = 2.0 * np.eye(m)
            for lam_c, Cc in zip(lam, curv):
                if Cc is not None:
                    Hl = Hl + lam_c * Cc
            grad = 2.0 * u + Ju.T @ lam
            KKT = np.block([[Hl, Ju.T], [Ju, np.zeros((nc, nc))]])
            rhs_k = -np.concatenate([grad, g])
            try:
                step = np.linalg.solve(KKT, rhs_k)
            except np.linalg.LinAlgError:
                step = np.linalg.lstsq(KKT, rhs_k, rcond=None)[0]
            if not np.all(np.isfinite(step)):
                break
            du, dl = step[:m], step[m:]
            # damped step on the constraint residual (the full step is nearly always accepted)
            t = 1.0
            gnorm = np.max(np.abs(g) / scale)
            for _ in range(12):
                y_n, g_n, Ju_n = evaluate(u + t * du)
                if np.max(np.abs(g_n) / scale) <= max(gnorm, 1e-15) * (1.0 + 1e-3) or gnorm < 1e-13:
                    break
                t *= 0.5
            u = u + t * du
            lam = lam + t * dl
            y, g, Ju = y_n, g_n, Ju_n
            small_step = np.linalg.norm(t * du) <= 1e-15 * max(np.linalg.norm(c), np.linalg.norm(u), 1e-300)
            feasible = np.max(np.abs(g) / scale) <= 4e-16
            stationary = np.linalg.norm(2.0 * u + Ju.T @ lam) <= 1e-13 * max(np.linalg.norm(Ju.T @ lam), 1e-300)
            if small_step or (feasible and stationary):
                converged = True
                break
        return y, float(u @ u), converged, nit

    # Start from the unconstrained minimiser (u = 0).  The quadratic constraints can have several
    # KKT points; when they are strongly active (prototype solver, tiny Krylov spaces) also start
    # from the caller's warm start -- the point SLSQP starts from -- and keep the better one.
    y, fval, converged, nit = newton(np.zeros(m))
    y0 = np.asarray(y0, dtype=float)
    if y0.shape == (m,) and (not converged or fval > 1e-4 * float(c @ c)):
        y_b, f_b, conv_b, nit_b = newton(R @ y0 - c)
        if conv_b and (not converged or f_b < fval):
            y, fval, converged, nit = y_b, f_b, conv_b, nit + nit_b
    if converged:
        y = _settle_signs(y, cons)
    msg = SUCCESS_MESSAGE if converged else "Iteration limit reached"
    return SmallResult(y, message=msg, success=converged, nit=nit, fun=fval)


NATIVE_KKT = True          # tests switch this off to compare the two implementations


def _kkt_native(Hj, beta, constraints):
    """The plain case -- all constraints class-form quadratics, Newton converging from the least-squares start --
    in C++ (spis_small_kkt: same algorithm, ~20 us instead of 0.4-1.2 ms of numpy calls; this solve sits on the
    critical path of every constrained iteration with the GPU idle).  None: not handled, take the route below."""
    if not NATIVE_KKT:
        return None
    cons = list(constraints)
    m = Hj.shape[1]
    if m < 1 or not cons or not all(c.quadratic for c in cons):
        return None
    try:
        lib = _nat.load_library()
    except Exception:
        return None
    H = np.ascontiguousarray(Hj, dtype=np.float64)
    nc = len(cons)
    t0 = np.array([float(c.term0) for c in cons], dtype=np.float64)
    t1 = np.ascontiguousarray(np.stack([np.asarray(c.term1, dtype=np.float64).reshape(m) for c in cons]))
    t2 = np.ascontiguousarray(np.stack([np.asarray(c.term2, dtype=np.float64).reshape(m, m) for c in cons]))
    y = np.empty(m)
    fval = _C.c_double(0.0)
    nit, handled = _C.c_int(0), _C.c_int(0)
    rc = lib.spis_small_kkt(m, m, _nat.dptr(H), float(beta), nc, _nat.dptr(t0), _nat.dptr(t1), _nat.dptr(t2),
                            _nat.dptr(y), _C.byref(fval), _C.byref(nit), _C.byref(handled))
    if rc != _nat.OK or not handled.value:
        return None
    # (spis_small_kkt has already moved y a few ulps to the accepted side of every constraint, spis_small_settle;)
    # the signs are settled HERE, with the very evaluation the acceptance test of solvers.py:266 uses -- normally
    # that is one look at g_c(y) and no further move
    return SmallResult(_settle_signs(y, cons), message=SUCCESS_MESSAGE, success=True, nit=nit.value, fun=fval.value)


def _settle_signs(y, cons, tries=8):
    """The reference accepts a constrained step only if max_c g_c(y) <= 1e-12 -- a SIGNED, ABSOLUTE
    test (solvers.py:14-18,266; quirks Q4/a6) -- and otherwise throws the constrained y away.  For
    invariants of size 1e4 one ulp is 3.6e-12, so a minimiser that is exact to rounding fails that
    test half of the time (SLSQP's final projection lands on g = 0.0 exactly more often than not).
    Newton's y is moved by a minimum-norm correction of a few ulps so that every g_c sits at or just
    below zero: the same feasible point to rounding, on the side of zero the reference accepts."""
    eps = np.finfo(float).eps
    scale = np.array([abs(c.term0) if c.quadratic else 1.0 for c in cons])
    for k in range(tries):
        g = np.array([c.fun(y) for c in cons], dtype=float)
        if not np.all(np.isfinite(g)) or g.max() <= 0.0:
            return y
        J = np.array([np.asarray(c.jac(y), dtype=float).reshape(-1) for c in cons])
        target = -(2.0 ** k) * 2.0 * eps * np.maximum(scale, np.abs(g))
        dy = np.linalg.lstsq(J, target - g, rcond=None)[0]
        if not np.all(np.isfinite(dy)):
            return y
        y = y + dy
    return y
