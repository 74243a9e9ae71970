"""B200-native drop-in for the reference's ``solvers.py`` (FGMRES / CGMRES / prototypical CGMRES).

Same public names, keyword names, constraint-list conventions and ``(x, dict)`` return
contract as /root/reference/solvers.py:

    gmres(A, b, x0, k, tol=1e-50, pre=None)                                   solvers.py:58
    cgmres(A, b, x0, k, tol=1e-8, contol=10, conlist=[], pre=None, timing=None)  solvers.py:131
    cgmres_p(A, b, x0, k, conlist=[], pre=None)                               solvers.py:328
    constraint_container(const, x0, Z), constraint_checker(x, const_list)     solvers.py:21,14

so each experiment's ``LinearSolver.py`` (lkdv, lkdvRK, swe, heat) can ``import`` this module
as ``solvers`` unchanged.  Every O(n) operation -- SpMV, preconditioner, Gram-Schmidt, the
iterate x0 + Z y, the true residual and the constraint Gram/projection terms -- runs as a
hand-written sm_100a kernel behind the C ABI of include/spis_b200.h; the host keeps the control
flow of the reference loop and the k-dimensional (constrained) least-squares solve.

Extra keyword-only arguments (all optional, defaults keep reference semantics):
    session       a DeviceSession with the system already resident on the GPU
    small_solver  'slsqp' (reference's scipy SLSQP calls, default) or 'kkt' (QR + Newton-KKT)
    lookahead     overlap Arnoldi step j+1 with the host solve of step j (default True)
    history       'lazy' (default: dict['x'] materialises iterates on demand) or 'eager'
    device        CUDA device ordinal

There is no CPU fallback: without libspis_b200.so or without an sm_100 GPU every solver raises.
"""
from __future__ import annotations

import collections.abc
import os
import sys
import threading
import warnings
import weakref
from time import time

import numpy as np
import scipy.sparse as sps

from . import _native as nat
from . import smallsolve
from .device import KrylovContext
from .preconditioners import BlockJacobiPreconditioner, JacobiPreconditioner

__all__ = ["gmres", "cgmres", "cgmres_p", "constraint_container", "constraint_checker",
           "DeviceSession", "configure"]

_CONFIG = {
    "small_solver": "slsqp",
    "lookahead": True,
    "history": "lazy",
    "device": 0,
    "orth": "cgs2",
    "spmv_format": "auto",
    "profile": False,
    "async_setup": True,
    "dual_spmv": True,
    "lazy_sessions": 2,        # solver-created sessions kept resident for lazy dict['x'] access (the newest ones)
    "pipeline": True,          # device-resident loop (Givens update and unconstrained iterates on the GPU) with small_solver='kkt'
    "early_download": True,    # cgmres: the result starts travelling to the host while the last iterate is formed and checked (int: row chunks)
    "ctx_options": {},         # raw spis_set_option pairs applied to every context a session creates (tuning / A-B runs)
}
_ORTH = {"cgs2": nat.ORTH_CGS2, "cgs1": nat.ORTH_CGS1, "mgs": nat.ORTH_MGS}
_FMT = {"auto": nat.FMT_AUTO, "sell": nat.FMT_SELL, "csr": nat.FMT_CSR, "sell2": nat.FMT_SELL2, "pattern": nat.FMT_PATTERN, "selld": nat.FMT_SELLD}
_BREAKDOWN = ("GMRES broke down, either initial guess is exact or , more likely, "
              "something has gone wrong.")


def configure(**kwargs):
    """Set module-wide defaults for the extension keywords (e.g. small_solver='kkt')."""
    for key, val in kwargs.items():
        if key not in _CONFIG:
            raise KeyError(f"unknown option {key!r}; known: {sorted(_CONFIG)}")
        _CONFIG[key] = val
    return dict(_CONFIG)


def _opt(name, value):
    return _CONFIG[name] if value is None else value


_TRACE = bool(os.environ.get("SPIS_TRACE"))


class _Buckets:
    """SPIS_TRACE=1: wall-clock spent in each phase of the Krylov loop, summed over iterations."""

    def __init__(self):
        self.acc = {}
        self.t = time()

    def mark(self, label):
        if _TRACE:
            now = time()
            self.acc[label] = self.acc.get(label, 0.0) + (now - self.t)
            self.t = now

    def report(self, name):
        if _TRACE and self.acc:
            sys.stderr.write("[spis trace] %s loop: %s\n" % (name, ", ".join("%s %.2f ms" % (k, v * 1e3) for k, v in self.acc.items())))


class _Trace:
    """SPIS_TRACE=1 prints the wall-clock of each host-side phase to stderr."""

    def __init__(self):
        self.t = time()

    def __call__(self, label):
        if _TRACE:
            now = time()
            sys.stderr.write("[spis trace] %-34s %8.2f ms\n" % (label, (now - self.t) * 1e3))
            self.t = now


# ==============================================================================================
# constraint helpers (solvers.py:14-53)
# ==============================================================================================
def constraint_checker(x, const_list):
    """Largest signed constraint value over a list of {'fun': ...} dicts (solvers.py:14-18)."""
    dev = 0
    for const in const_list:
        dev = max(dev, const["fun"](x))
    return dev


def _classify_constraint(const):
    """'class' (object with M, v, c), 'dict' ({'func','jac'} callbacks) or 'invalid' (solvers.py:24-30)."""
    if hasattr(const, "__dict__"):
        return "class"
    if type(const) is dict:
        return "dict"
    return "invalid"


class constraint_container:
    """API-compatible stand-alone container (solvers.py:21-53) for a HOST matrix Z (n x m).

    Class-form constraints get term0/term1/term2 from the same device kernels the solvers use
    (SpMV for M@Z, tall-skinny dots for Z^T MZ); dict-form constraints just store their callbacks.
    Inside gmres/cgmres/cgmres_p the equivalent object is built incrementally by DeviceSession.
    """

    def __init__(self, const, x0, Z, device=None):
        kind = _classify_constraint(const)
        if kind == "invalid":
            raise NotImplementedError("Constraints must be either dictionaries or classes")
        self.optimise = True if kind == "class" else None
        if not self.optimise:
            self.const, self.x0, self.Z = const, x0, Z
            return
        Zc = np.ascontiguousarray(np.asarray(Z, dtype=np.float64).T)        # rows = columns of Z
        m, n = Zc.shape
        x0 = nat.as_f64(x0, n)
        v = nat.as_f64(const.v, n)
        with KrylovContext(n, max(m, 1), device=_opt("device", device)) as ctx:
            ctx.upload_matrix(nat.SLOT_CON0, const.M)
            MZ = np.empty((m, n))
            for i in range(m):
                MZ[i] = ctx.op_spmv(nat.SLOT_CON0, Zc[i])
            Mx0 = ctx.op_spmv(nat.SLOT_CON0, x0)
            gram = np.empty((m, m))
            for i in range(m):
                gram[:, i] = ctx.op_mdot(Zc, MZ[i])[:m]                    # Z^T (M z_i)
            self.term0 = 0.5 * ctx.op_mdot(Mx0[None, :], x0)[0] + const.c + ctx.op_mdot(v[None, :], x0)[0]
            self.term1 = ctx.op_mdot(Zc, v)[:m] + ctx.op_mdot(MZ, x0)[:m]
            self.term2 = 0.5 * gram
            self.MZ = MZ.T

    def constraint_func(self, z):
        if self.optimise:
            return self.term0 + self.term1 @ z + z @ self.term2 @ z
        return self.const["func"](z, self.x0, self.Z)

    def constraint_jac(self, z):
        if self.optimise:
            return self.term1 + 2 * z @ self.term2
        return self.const["jac"](z, self.x0, self.Z)


# ==============================================================================================
# device session: one linear system + constraints resident on one GPU
# ==============================================================================================
class DeviceSession:
    """Uploads A, b, x0, the preconditioner and the class-form constraint data once.

    Replaces the host-resident scipy/numpy objects the reference keeps between iterations
    (solvers.py:149-181).  Re-usable across several solves of the same system (bench.py's
    device-resident timing; time stepping with a fixed operator).
    """

    def __init__(self, A, b, x0, k, conlist=(), pre=None, *, device=None, orth=None,
                 spmv_format=None, profile=None, ctx_factory=KrylovContext, async_setup=None):
        tr = _Trace()
        b = nat.as_f64(b)
        n = b.size
        self.n, self.k = n, int(k)
        self.x0_host = nat.as_f64(x0, n)
        self.ctx = ctx_factory(n, self.k, device=_opt("device", device))
        ctx = self.ctx
        tr("context create")
        if _TRACE and hasattr(ctx, "info"):
            sys.stderr.write("[spis trace] device block cache: %d hits, %d misses (%.1f MB) so far\n"
                             % (ctx.info("alloc_hits"), ctx.info("alloc_misses"), ctx.info("alloc_miss_bytes") / 1e6))
        ctx.set_option("orth", _ORTH[_opt("orth", orth)])
        ctx.set_option("spmv_format", _FMT[_opt("spmv_format", spmv_format)])
        ctx.set_option("profile", 1 if _opt("profile", profile) else 0)
        for key, val in dict(_CONFIG.get("ctx_options") or {}).items():
            ctx.set_option(key, val)
        if not (sps.issparse(A) or isinstance(A, np.ndarray)):
            raise TypeError("A must be a scipy.sparse matrix (or a dense array); got %r" % type(A))
        if A.shape[0] != n or A.shape[1] != n + getattr(ctx, "n_halo", 0):
            raise ValueError(f"A has shape {A.shape}, expected {(n, n + getattr(ctx, 'n_halo', 0))}")
        # the zero test of x0 (host threads) runs while the calling thread waits for the upload of A
        x0_scan = {}
        scanner = threading.Thread(target=lambda: x0_scan.setdefault("nz", nat.any_nonzero(self.x0_host)), daemon=True)
        scanner.start()
        ctx.upload_matrix(nat.SLOT_A, A)
        tr("upload A")
        ctx.upload_vec(nat.VEC_B, b)
        tr("upload b")
        scanner.join()
        x0_nonzero = self._any_rank(lambda: x0_scan["nz"] if "nz" in x0_scan else nat.any_nonzero(self.x0_host), "x0")
        tr("x0 zero scan")
        if x0_nonzero:                       # a new context's x0 buffer is already zero on the device
            ctx.upload_vec(nat.VEC_X0, self.x0_host)
        ctx.set_option("x0_is_zero", 0 if x0_nonzero else 1)
        tr("upload x0")
        self._host_pre = None
        self._setup_precond(pre)
        tr("preconditioner")
        self._Zhost = None
        self._Zrows = 0
        # The constraint data is first needed at the first constrained step (solvers.py:242-247), many
        # Krylov iterations from now: a helper thread scans and uploads it on the context's auxiliary
        # stream while the caller's thread already runs the loop.  (Row-sharded sessions take collective
        # decisions in _any_rank and stay on one thread.)
        self._bg = None
        self._bg_error = None
        self._native_jobs = False
        conlist = list(conlist)
        if len(conlist) > nat.MAX_SLOTS - nat.SLOT_CON0:
            raise ValueError("too many constraints")
        self._cons = [None] * len(conlist)
        sharded = type(self) is not DeviceSession          # row-sharded: the yes/no decisions were taken collectively (_preflags)
        if conlist and _opt("async_setup", async_setup) and hasattr(ctx, "use_aux_stream") and \
                (not sharded or (hasattr(ctx, "constraint_setup_async") and getattr(self, "_preflags", None) is not None
                                 and all(_classify_constraint(c) == "class" and sps.issparse(getattr(c, "M", None)) for c in conlist))):
            # one helper per constraint: the host scan of one (`0*A`) overlaps the PCIe upload of another.
            # Class-form constraints with a sparse M go to NATIVE helper threads (one C call each: they never
            # need the interpreter lock the loop's thread holds); anything else to a Python thread.
            self._bg = []
            for idx, const in enumerate(conlist):
                if (hasattr(ctx, "constraint_setup_async") and _classify_constraint(const) == "class"
                        and sps.issparse(getattr(const, "M", None))):
                    entry = {"kind": "class", "const": const, "error": None}
                    try:
                        if sharded:
                            ctx.constraint_setup_async(idx, const.M, const.v, float(const.c),
                                                       m_is_zero=not self._preflags[("M", idx)], v_is_zero=not self._preflags[("v", idx)])
                        else:
                            ctx.constraint_setup_async(idx, const.M, const.v, float(const.c))
                        self._native_jobs = True
                    except nat.NativeLibraryError:
                        raise
                    except Exception as exc:                # surfaces where the reference builds containers
                        entry["error"] = exc
                    self._cons[idx] = entry
                else:
                    self._bg.append(threading.Thread(target=self._setup_constraints_bg, args=(idx, const), daemon=True))
            for th in self._bg:
                th.start()
            tr("constraints (handed to helper threads)")
        else:
            for idx, const in enumerate(conlist):
                self._setup_constraint(idx, const)
            tr("constraints")

    def update(self, b=None, x0=None, constants=None, vectors=None):
        """Next system of a time loop over a FIXED operator (lkdv/Evolve.py:39-56 re-assembles every step, but
        only b, x0 and the invariant values change): new right-hand side / initial guess / constraint scalars
        `c` and linear terms `v` (one per class-form constraint, None = keep); A, the constraint matrices and the
        Krylov workspace stay where they are."""
        self._join_setup()
        ctx = self.ctx
        if b is not None:
            ctx.upload_vec(nat.VEC_B, nat.as_f64(b, self.n))
        if x0 is not None:
            self.x0_host = nat.as_f64(x0, self.n)
            x0_nonzero = self._any_rank(nat.any_nonzero(self.x0_host))
            ctx.upload_vec(nat.VEC_X0, self.x0_host)      # (zeros included: the buffer may hold an older guess)
            ctx.set_option("x0_is_zero", 0 if x0_nonzero else 1)
        if constants is not None:
            constants = list(constants)
            if len(constants) != len(self._cons):
                raise ValueError(f"{len(constants)} constants for {len(self._cons)} constraints")
            for idx, (entry, cc) in enumerate(zip(self._cons, constants)):
                if cc is None:
                    continue
                if entry["kind"] != "class" or entry["error"] is not None:
                    raise ValueError(f"constraint {idx} is not a class-form constraint held on the device")
                ctx.constraint_set_constant(idx, float(cc))
        if vectors is not None:
            vectors = list(vectors)
            if len(vectors) != len(self._cons):
                raise ValueError(f"{len(vectors)} vectors for {len(self._cons)} constraints")
            for idx, (entry, vv) in enumerate(zip(self._cons, vectors)):
                if vv is None:
                    continue
                if entry["kind"] != "class" or entry["error"] is not None:
                    raise ValueError(f"constraint {idx} is not a class-form constraint held on the device")
                vv = nat.as_f64(vv, self.n)
                ctx.constraint_set_vector(idx, vv if self._any_rank(nat.any_nonzero(vv)) else None)

    def _any_rank(self, flag, key=None):
        """Logical OR of a host-side decision over all ranks (identity on one GPU).  Every decision
        that changes the sequence of device reductions must be taken identically on all ranks.
        (`key` names the decision: a row-sharded session takes all of them in one collective up front and does not
        evaluate a callable `flag` again.)"""
        return bool(flag() if callable(flag) else flag)

    # -- preconditioner: solvers.py:149-161 ------------------------------------------------------
    def _setup_precond(self, pre):
        ctx = self.ctx
        if pre is None:
            ctx.set_precond(nat.PRE_NONE)
        elif isinstance(pre, JacobiPreconditioner):
            ctx.upload_vec(nat.VEC_PRE_DIAG, pre.dinv)
            ctx.set_precond(nat.PRE_JACOBI)
        elif isinstance(pre, BlockJacobiPreconditioner):
            ctx.upload_blocks(pre.inv_blocks, pre.stride_block, pre.stride_field)
            ctx.set_precond(nat.PRE_BLOCK)
        elif hasattr(pre, "solve"):
            self._host_pre = pre.solve                      # e.g. SuperLU from spilu (swe/TimedSolve.py:23)
            ctx.set_precond(nat.PRE_HOST)
        elif sps.issparse(pre) or (isinstance(pre, np.ndarray) and pre.ndim == 2):
            P = sps.csr_matrix(pre)
            if P.shape != (self.n, self.n):
                raise ValueError("Preconditioner not supported")
            coo = P.tocoo()
            if np.array_equal(coo.row, coo.col):
                ctx.upload_vec(nat.VEC_PRE_DIAG, P.diagonal())
                ctx.set_precond(nat.PRE_JACOBI)
            else:
                ctx.upload_matrix(nat.SLOT_PRE, P)
                ctx.set_precond(nat.PRE_CSR)
        else:
            def apply(vec, _pre=pre):                       # LinearOperator, pyamg, ... (heat/TimedSolve.py:30-31)
                try:
                    return _pre @ vec
                except Exception:
                    raise ValueError("Preconditioner not supported")
            self._host_pre = apply
            ctx.set_precond(nat.PRE_HOST)

    # -- constraints: solvers.py:22-40 -----------------------------------------------------------
    def _setup_constraints_bg(self, idx, const):
        try:
            self.ctx.use_aux_stream(True)
            try:
                self._setup_constraint(idx, const)
            finally:
                self.ctx.use_aux_stream(False)
        except BaseException as exc:                       # re-raised on the caller's thread by _join_setup
            if self._bg_error is None:
                self._bg_error = exc

    def _join_setup(self):
        if self._bg is not None:
            for th in self._bg:
                th.join()
            self._bg = None
        if self._native_jobs:
            self._native_jobs = False
            self.ctx.constraint_setup_wait()
        if self._bg_error is not None:
            exc, self._bg_error = self._bg_error, None
            raise exc

    def _setup_constraint(self, idx, const):
        tr = _Trace()
        if True:
            kind = _classify_constraint(const)
            entry = {"kind": kind, "const": const, "error": None}
            if kind == "class":
                try:
                    M, v, c = const.M, const.v, const.c
                    anynz = getattr(self.ctx, "any_nonzero", nat.any_nonzero)
                    if sps.issparse(M):
                        M_zero = not self._any_rank(lambda: M.nnz != 0 and anynz(M.data), ("M", idx))
                    else:
                        M = np.asarray(M, dtype=np.float64)
                        M_zero = not self._any_rank(lambda: M.any(), ("M", idx))
                    tr("  constraint %d: is M zero? %s" % (idx, M_zero))
                    slot = -1
                    if not M_zero:
                        slot = nat.SLOT_CON0 + idx
                        self.ctx.upload_matrix(slot, M)
                        tr("  constraint %d: upload M" % idx)
                    v = nat.as_f64(v, self.n)
                    if type(self) is DeviceSession:        # the library recognises an all-zero v itself
                        self.ctx.constraint_define(idx, slot, v, float(c))
                    else:                                  # row-sharded: every rank must take the same decision
                        self.ctx.constraint_define(idx, slot, v if self._any_rank(lambda: nat.any_nonzero(v), ("v", idx)) else None, float(c))
                    tr("  constraint %d: v" % idx)
                except nat.NativeLibraryError:
                    raise
                except Exception as exc:                    # surfaces where the reference builds containers
                    entry["error"] = exc
            self._cons[idx] = entry

    def containers(self, m):
        """Reduced constraints for Z = z[:m].T (the reference rebuilds these per step, solvers.py:242-247)."""
        self._join_setup()
        out = []
        for entry in self._cons:
            if entry["kind"] == "invalid":
                raise NotImplementedError("Constraints must be either dictionaries or classes")
            if entry["error"] is not None:
                raise entry["error"]
        classes = [idx for idx, entry in enumerate(self._cons) if entry["kind"] == "class"]
        batch = {}
        if len(classes) > 1 and hasattr(self.ctx, "constraint_terms_batch"):
            batch = dict(zip(classes, self.ctx.constraint_terms_batch(classes, m)))    # one device round trip for all of them
        for idx, entry in enumerate(self._cons):
            if entry["kind"] == "class":
                t0, t1, t2 = batch[idx] if idx in batch else self.ctx.constraint_terms(idx, m)
                out.append(smallsolve.ReducedConstraint(t0, t1, t2))
            else:
                out.append(smallsolve.ReducedConstraint(callbacks=entry["const"], x0=self.x0_host,
                                                        Z=self._host_Z(m)))
        return out

    def _host_Z(self, m):
        """Host copy of Z (n x m, Fortran-ordered view like np.transpose(z[:m]), solvers.py:207)."""
        if self._Zhost is None:
            self._Zhost = np.empty((self.k, self.n))
        if m > self._Zrows:
            self._Zhost[self._Zrows:m] = self.ctx.download_Z(self._Zrows, m)
            self._Zrows = m
        return self._Zhost[:m].T

    @property
    def n_constraints(self):
        self._join_setup()
        return len(self._cons)

    # -- Krylov primitives ------------------------------------------------------------------------
    def begin(self):
        self._Zrows = 0
        return self.ctx.solve_begin()

    def arnoldi_launch(self, j):
        if self._host_pre is not None:
            q = self.ctx.host_pre_get(j)
            self.ctx.host_pre_put(j, np.asarray(self._host_pre(q)))
        self.ctx.arnoldi_launch(j)

    def arnoldi_wait(self, j):
        return self.ctx.arnoldi_wait(j)

    @property
    def can_fuse_iterate(self):
        """True when the iterate of step j can be formed by the last projection sweep of step j+1 (CGS2,
        no preconditioner: Z is V).  Asked once per solve."""
        ctx = self.ctx
        return (self._host_pre is None and hasattr(ctx, "arnoldi_begin") and hasattr(ctx, "info")
                and ctx.info("can_fuse_iterate") == 1)

    def close(self):
        try:
            self._join_setup()
        finally:
            self.ctx.close()


class IterateHistory(collections.abc.Sequence):
    """dict['x'] of the reference (solvers.py:165-169,287,318): [r0, x_1, ..., x_steps].

    The iterates stay on the device as their Krylov coefficients; item access re-forms
    x_j = x0 + Z[:, :m_j] y_j with the same kernel that produced it and downloads it.
    """

    def __init__(self, session):
        self._session = session
        self._gen = session.ctx.generation
        self._ys = [None]                 # entry 0 is r0 (quirk: x[0] is the initial residual)
        self._cache = {}

    def _append(self, y):
        self._ys.append(np.array(y, dtype=np.float64, copy=True))

    def _set_cached(self, idx, arr):
        self._cache[idx % len(self._ys)] = arr

    def __len__(self):
        return len(self._ys)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return [self[i] for i in range(*idx.indices(len(self)))]
        idx = int(idx)
        if idx < 0:
            idx += len(self)
        if not 0 <= idx < len(self):
            raise IndexError("iterate index out of range")
        if idx in self._cache:
            return self._cache[idx]
        ctx = self._session.ctx
        if ctx.closed or ctx.generation != self._gen:
            raise RuntimeError("the device session that holds these iterates was closed or reused; "
                               "use history='eager' to copy all iterates to the host")
        if idx == 0:
            arr = ctx.download(nat.VEC_R0)
        else:
            ctx.form_iterate(self._ys[idx])
            arr = ctx.download(nat.VEC_X)
        self._cache[idx] = arr
        return arr

    def materialise(self):
        return [self[i] for i in range(len(self))]

    def detach(self):
        """Copy every iterate to the host; the history no longer needs (or pins) the device session."""
        for i in range(len(self)):
            self[i]
        return self


# ==============================================================================================
# shared Arnoldi driver
# ==============================================================================================
class _Arnoldi:
    """Feeds Hessenberg columns to the solvers; hides launch/wait lookahead (solvers.py:190-198)."""

    def __init__(self, sess, k, lookahead, beta=None):
        self.sess, self.k, self.lookahead = sess, k, bool(lookahead)
        self.H = np.zeros((k + 1, k))
        self._inflight = None                 # step whose column is on its way to the host
        self._begun = None                    # step with only its first half queued (fused-iterate mode)
        self._have = -1
        self._fuse = self.lookahead and sess.can_fuse_iterate
        # residual of iterate j measured by the SpMV of Arnoldi step j+2 (one pass over A for both products)
        self._dual = self._fuse and hasattr(sess.ctx, "arnoldi_begin_residual") and _opt("dual_spmv", None)
        self._ls_prev = None                  # unconstrained least-squares residual one column earlier
        # Givens recurrence of the unconstrained least-squares residual |beta e1 - H y|_min: used only
        # to decide whether launching the NEXT Arnoldi step ahead of time can be wasted work
        self._cs = np.zeros(k)
        self._sn = np.zeros(k)
        self._g = None if beta is None else float(beta)
        # ... and, kept along, the triangular factor and the rotated right-hand side: the unconstrained minimiser
        # of the 'kkt' engine is one back substitution (ls_solution) instead of an SVD-based lstsq per iteration
        self._R = np.zeros((k, k))
        self._gv = np.zeros(k + 1)
        if beta is not None:
            self._gv[0] = float(beta)

    def column(self, j):
        if self._have == j:                   # fetched while the previous iterate/residual pair ran
            return self.H[: j + 2, j].copy()
        if self._begun == j:                  # first half queued, nothing to fuse into the second
            self.sess.ctx.arnoldi_finish(j)
            self._begun, self._inflight = None, j
        if self._inflight != j:
            self.sess.arnoldi_launch(j)
        col = self.sess.arnoldi_wait(j)
        self._inflight = None
        self._store(j, col)
        return col

    def _store(self, j, col):
        self._have = j
        self.H[: j + 2, j] = col
        self._ls_prev = self._g if self._g is None else abs(self._g)
        if self._g is not None:
            r = np.array(col, dtype=np.float64)
            for i in range(j):
                a, b = r[i], r[i + 1]
                r[i] = self._cs[i] * a + self._sn[i] * b
                r[i + 1] = -self._sn[i] * a + self._cs[i] * b
            den = np.hypot(r[j], r[j + 1])
            if den > 0:
                self._cs[j], self._sn[j] = r[j] / den, r[j + 1] / den
                self._g = -self._sn[j] * self._g
                self._R[: j, j] = r[:j]
                self._R[j, j] = den
                gj = self._gv[j]
                self._gv[j], self._gv[j + 1] = self._cs[j] * gj, -self._sn[j] * gj
            else:
                self._g = None

    def residual(self, yk, j, may_end, tol=None):
        """x_j = x0 + Z yk and ||A x_j - b|| (solvers.py:287,290).  While the device works on them the host
        collects Hessenberg column j+1 (its Arnoldi step was queued BEFORE this pair) and, unless the loop
        is about to end, queues step j+2 BEHIND the pair: the device never waits for the host."""
        ctx = self.sess.ctx
        if not (self.lookahead and hasattr(ctx, "iterate_residual_launch")):
            return ctx.iterate_residual(yk)
        rides = False
        if self._begun == j + 1:
            # the last projection of step j+1 sweeps the same basis rows the iterate needs: one pass for both
            ctx.arnoldi_finish(j + 1, yk)
            self._begun, self._inflight = None, j + 1
            # x_j and q_{j+2} now exist: if step j+2 is going to run anyway, its SpMV measures ||A x_j - b|| on
            # the way (one pass over A for both).  Column j+1 is not known yet, so "going to run" is a guess
            # from the convergence rate; a wrong guess costs what a wasted lookahead always cost, half a step.
            rides = bool(self._dual) and not may_end and j + 2 < self.k and self._next_step_expected(tol)
            if rides:
                ctx.arnoldi_begin_residual(j + 2)
                self._begun = j + 2
            else:
                ctx.residual_launch()
        else:
            ctx.iterate_residual_launch(yk)
        if self._inflight == j + 1:
            col = self.sess.arnoldi_wait(j + 1)
            self._inflight = None
            self._store(j + 1, col)
            # step j+2 is only useful if neither this iteration nor the next one ends the loop; the next
            # one can only end if even its unconstrained minimiser is below tol (known now, from column j+1)
            next_may_end = tol is not None and self.ls_residual() < tol
            if not rides and not may_end and not next_may_end and col[j + 2] != 0 and j + 2 < self.k:
                self._queue(j + 2)
        return ctx.iterate_residual_wait()

    def _next_step_expected(self, tol):
        """Will the iteration after this one need another Arnoldi step?  Extrapolates the unconstrained
        least-squares residual by its last reduction factor (solvers that never stop early pass tol=None)."""
        if tol is None or self._g is None:
            return True
        ls = abs(self._g)
        rate = 1.0 if not self._ls_prev else min(1.0, ls / self._ls_prev)
        return ls * rate * rate >= tol

    def ls_solution(self, m):
        """argmin_y |beta e1 - H[:m+1, :m] y| from the Givens QR kept by _store (None when it is not tracked or R
        is singular to working precision: the caller then takes the general least-squares route)."""
        if self._g is None or self._have != m - 1:
            return None
        R = self._R[:m, :m]
        d = np.abs(np.diag(R))
        if m == 0 or not d.min() > 1e-14 * d.max():
            return None
        return np.linalg.solve(R, self._gv[:m])        # triangular: LU never swaps a row, i.e. back substitution

    def ls_residual(self):
        """min_y |beta e1 - H_j y| after the last column() (inf when not tracked)."""
        return np.inf if self._g is None else abs(self._g)

    def predicted_residual(self, y, beta):
        """|beta e1 - H_j y| for a given coefficient vector: the residual the device will report for
        x0 + Z y up to rounding (Arnoldi relation A Z = V H)."""
        m = len(y)
        r = -(self.H[: m + 1, :m] @ y)
        r[0] += beta
        return float(np.linalg.norm(r))

    def _queue(self, j):
        if self._fuse:
            self.sess.ctx.arnoldi_begin(j)        # second half follows once the iterate coefficients exist
            self._begun = j
        else:
            self.sess.arnoldi_launch(j)
            self._inflight = j

    def prefetch(self, j):
        if self.lookahead and j < self.k and self._inflight is None and self._begun is None and self._have != j:
            self._queue(j)

    def drain(self):
        if self._begun is not None:               # queued speculatively, not needed: finish it so the context is idle
            self.sess.ctx.arnoldi_finish(self._begun)
            self._inflight, self._begun = self._begun, None
        if self._inflight is not None:
            self.sess.arnoldi_wait(self._inflight)
            self._inflight = None


class _Pipeline:
    """Drives the device-resident Krylov loop (spis_pipe_begin / spis_step_enqueue, include/spis_b200.h).

    Whole Arnoldi steps are queued AHEAD of the host: the Givens update and the least-squares coefficients y_j are
    computed by a one-warp kernel, the unconstrained iterate x_j = x0 + Z y_j (solvers.py:287) is formed by a sweep that
    was queued before y_j existed, and its true residual (solvers.py:290) is measured by the SpMV of a later step.  The
    host reads records (Hessenberg column, y_j, min |beta e1 - H y|, residual norms) from mapped memory, follows the
    reference's control flow with them and decides how far ahead to queue.  Index conventions:
        lag = 1 without a preconditioner: step s forms x_{s-1} in its last sweep (Z is V) and measures x_{s-2};
        lag = 0 with a device preconditioner: step s forms x_s with a sweep over Z and measures x_{s-1}.
    The device stops forming iterates by itself (phase word) once a measured residual is <= thr, so steps queued
    ahead of the host's decision never overwrite an iterate the host is responsible for.
    """

    DEPTH = 3                      # steps queued beyond the last record the host has seen (records live in rings of 8)

    def __init__(self, sess, k, beta, thr, tol, last_unconstrained):
        self.sess, self.ctx, self.k = sess, sess.ctx, k
        self.beta, self.thr, self.tol = float(beta), float(thr), tol
        self.thr2 = self.thr * self.thr
        self.last_it = last_unconstrained          # largest iterate index the device may form (cgmres: k-2, gmres: k-1)
        self.lag = int(sess.ctx.info("pipe_lag"))
        self.H = np.zeros((k + 1, k))
        self.enq = 0
        self.rec = {}                              # step -> (col, y, info)
        self.it_step = {}                          # iterate index -> step that was asked to form it
        self.res_ticket = {}                       # iterate index -> ticket of its residual measurement
        self.measured = set()
        self.device_iter = True                    # False once the host forms the iterates itself
        self.x_holds = None                        # iterate index known to be in the X buffer
        self.early = None                          # (iterate index, page-locked array) of a download started early
        self._ls_prev = None
        self.ctx.set_option("spmv_dual", 1 if _opt("dual_spmv", None) else 0)
        self.ctx.pipe_begin(self.thr, not (self.beta > self.thr))

    # -- queueing ---------------------------------------------------------------------------------------------
    def _enqueue(self):
        s = self.enq
        i = s - self.lag
        want_it = self.device_iter and 0 <= i <= self.last_it
        ip = s - 1 - self.lag                      # the iterate step s-1 was asked to form sits in X when step s multiplies
        want_res = ip >= 0 and self.it_step.get(ip) == s - 1 and ip not in self.res_ticket and ip not in self.measured
        ticket = self.ctx.step_enqueue(s, want_res, want_it)
        if want_res:
            self.res_ticket[ip] = ticket
        if want_it:
            self.it_step[i] = s
        self.enq = s + 1

    def _record(self, s):
        if s not in self.rec:
            while self.enq <= s:
                self._enqueue()
            self.rec[s] = self.ctx.step_wait(s)
            self.rec.pop(s - 16, None)
        return self.rec[s]

    def column(self, j):
        col, y, info = self._record(j)
        self.H[: j + 2, j] = col
        self._speculate(j, info)
        return col

    def _speculate(self, j, info):
        """Queue the steps that are (almost) certainly needed so that the device never waits for the host: step s is
        needed iff iteration s-1 does not end the loop, which it cannot while even the unconstrained minimiser
        |beta e1 - H y|_min (known for step j, extrapolated by the last reduction factor beyond) stays above tol."""
        if not info["valid"]:
            return
        ls = info["ls"]
        rate = 1.0 if not self._ls_prev else min(1.0, ls / self._ls_prev)
        self._ls_prev = ls
        est = ls
        for s in range(j + 1, min(self.k, j + 1 + self.DEPTH)):
            if self.tol is not None:
                need = est >= self.tol if s == j + 1 else est * rate >= self.tol
                if not need:
                    break
            if self.enq <= s:
                self._enqueue()
            est *= rate

    def prefetch(self, s):
        if s < self.k:
            while self.enq <= s:
                self._enqueue()

    # -- small problem ------------------------------------------------------------------------------------------
    def ls_solution(self, m):
        col, y, info = self._record(m - 1)
        return y.copy() if info["valid"] else None

    def ls_residual(self, j):
        info = self._record(j)[2]
        return info["ls"] if info["valid"] else np.inf

    def predicted_residual(self, y):
        m = len(y)
        r = -(self.H[: m + 1, :m] @ y)
        r[0] += self.beta
        return float(np.linalg.norm(r))

    # -- iterates -------------------------------------------------------------------------------------------------
    def device_iterate_residual(self, j, yk, cannot_end):
        """||A x_j - b|| of the UNCONSTRAINED iterate x_j = x0 + Z yk.  Normally the device has formed (or is about to
        form) it from its own copy of yk; the norm rides on the SpMV of the next step when that step is needed anyway,
        else it takes a pass of its own.  `cannot_end`: iteration j cannot be the last one (cgmres only stops on a
        constrained step).  Returns (norm, go): go = the next iteration is still unconstrained as far as the residual
        is concerned (solvers.py:230)."""
        ctx = self.ctx
        s_form = j + self.lag
        formed = False
        if self.device_iter and j <= self.last_it and s_form < self.k:
            while self.enq <= s_form:
                self._enqueue()
            if self.it_step.get(j) == s_form:
                formed = self._record(s_form)[2]["phase"] == 0
        if not formed:
            # the device did not form this one (phase word set by a vanishing pivot, last Krylov vector, ...)
            self.device_iter = False
            res = ctx.iterate_residual(yk)
            self.x_holds = j
            self.measured.add(j)
            return res, res > self.thr
        self.x_holds = j
        if j not in self.res_ticket and s_form + 1 < self.k and self.enq == s_form + 1:
            # step s_form+1 is needed iff iteration s_form does not end the loop
            need = self.tol is None or (cannot_end and self.lag == 0) or self.ls_residual(s_form) >= self.tol
            if need:
                self._enqueue()                   # its SpMV measures x_j on the way
        if j in self.res_ticket:
            res, res2, go = ctx.resid_wait(self.res_ticket[j])
            self.measured.add(j)
            if not go and res2 > self.thr2:       # the phase word was set for another reason: the host forms the iterates from now on
                self.device_iter = False
            return res, res2 > self.thr2
        if self.enq > s_form + 1:
            # a later step is queued without the measurement (cannot happen with the rules above); re-form to be safe
            res = ctx.iterate_residual(yk)
        else:
            ctx.residual_launch()
            res = ctx.iterate_residual_wait()
        self.measured.add(j)
        return res, res > self.thr

    def host_iterate_residual(self, j, yk, likely_last=False):
        """x_j = x0 + Z yk with coefficients from the host (constrained steps, fallbacks) and its residual norm.
        likely_last: the loop is expected to end on this iterate (solvers.py:296-297), so its download -- 8n bytes
        over PCIe, the largest non-kernel item of a solve -- starts while it is being formed and checked."""
        self.device_iter = False
        self.drop_download()
        if likely_last and _opt("early_download", None) and hasattr(self.ctx, "iterate_residual_launch_dl"):
            ed = _opt("early_download", None)
            # row chunks: ~2.5M rows each, at most 4 (a row-sharded strip of a million rows goes in one piece: the
            # chunk kernels and their copy hand-overs would cost more than the overlap gains)
            # A row-sharded run queues the copy behind the residual check instead (chunks = 0): a device-to-host copy
            # in flight holds up the peers' NVLink stores, and the check ends in an exchange (see the C side).
            # (Measured: at 5 M rows per rank the overlap still wins, 9.05 against 9.49 ms per solve on 2 GPUs; at
            #  2.5 M and below the two forms are equal, and the stall-free one is used.)
            if ed is True:
                chunks = max(1, min(4, self.ctx.n // 2_500_000))
                if chunks < 2 and self.ctx.info("sharded"):
                    chunks = 0
            else:
                chunks = int(ed)
            buf = self.ctx.iterate_residual_launch_dl(yk, chunks)
            if buf is not None:
                self.early = (j, buf)
        else:
            self.ctx.iterate_residual_launch(yk)
        res = self.ctx.iterate_residual_wait()
        self.x_holds = j
        self.measured.add(j)
        return res

    def drop_download(self):
        """An early download whose iterate was not the last one: wait for the copy and forget the buffer."""
        if self.early is not None:
            self.ctx.download_join()
            self.early = None

    def take_download(self, j_last):
        """The host copy of x_{j_last} if its early download was started (complete on return), else None."""
        if self.early is None:
            return None
        j, buf = self.early
        self.ctx.download_join()
        self.early = None
        return buf if (j == j_last and self.x_holds == j_last) else None

    def host_takes_over(self):
        """From now on the host forms every iterate (constrained phase): steps queued later do not touch X."""
        self.device_iter = False

    def finish(self, j_last, y_last):
        """Leave x_{j_last} in the X buffer (a step queued ahead may have formed a later iterate)."""
        if j_last is not None and self.x_holds != j_last:
            self.drop_download()
            self.ctx.form_iterate(y_last)
            self.x_holds = j_last


def _use_pipeline(sess, lookahead, engine):
    ctx = sess.ctx
    return (bool(lookahead) and engine == "kkt" and _opt("pipeline", None) and sess._host_pre is None
            and hasattr(ctx, "step_enqueue") and ctx.info("device_pipeline") == 1)


def _acquire(session, A, b, x0, k, conlist, pre, device, orth):
    if session is not None:
        if session.k < k:
            raise ValueError(f"session was created for k={session.k} < {k}")
        return session
    sess = DeviceSession(A, b, x0, k, conlist=conlist, pre=pre, device=device, orth=orth)
    sess._owned_by_solver = True
    return sess


# Sessions the solvers created themselves and that are kept alive only because a lazy dict['x'] may still be read:
# the newest `lazy_sessions` of them.  A caller that keeps many info dicts (convergence studies, SingleSolve-style
# scripts) would otherwise pin one multi-GB device workspace per call until the dicts are dropped.
_LAZY_SESSIONS = collections.OrderedDict()
_EAGER_BYTES = 32 << 20          # histories up to this size are simply copied to the host (and the session closed)
_EVICT_COPY_BYTES = 1 << 30      # an evicted history up to this size is copied to the host first, larger ones lapse


def _retire_lazy(keep):
    for key in list(_LAZY_SESSIONS):
        sess, href = _LAZY_SESSIONS[key]
        hist = href()
        if hist is None or sess.ctx.closed:
            _LAZY_SESSIONS.pop(key)
            if not sess.ctx.closed:
                sess.close()
    while len(_LAZY_SESSIONS) > keep:
        key, (sess, href) = _LAZY_SESSIONS.popitem(last=False)
        hist = href()
        if hist is not None and len(hist) * sess.n * 8 <= _EVICT_COPY_BYTES:
            hist.detach()
        sess.close()


def _finish_history(hist, history_mode, x_last):
    """dict['x'] for the caller.  'lazy' (default) keeps the iterates on the device as Krylov coefficients; for a
    session the solver created itself that pins the device workspace ((k+1) n doubles of basis and more), so small
    histories are copied eagerly and only the newest `lazy_sessions` lazy ones stay resident (older ones are copied to
    the host if they fit 1 GiB, else they lapse and say so when read)."""
    if x_last is not None:
        hist._set_cached(len(hist) - 1, x_last)
    if history_mode not in ("lazy", "eager"):
        raise ValueError("history must be 'lazy' or 'eager'")
    sess = hist._session
    own = getattr(sess, "_owned_by_solver", False)
    if history_mode == "eager" or (own and len(hist) * sess.n * 8 <= _EAGER_BYTES):
        out = hist.materialise()
        if own:
            sess.close()
        return out if history_mode == "eager" else hist.detach()
    if own:
        _retire_lazy(max(int(_opt("lazy_sessions", None)), 1) - 1)
        _LAZY_SESSIONS[id(sess)] = (sess, weakref.ref(hist))
    return hist


def _abandon(sess):
    """An exception is leaving a solver: do not leave a session it created to the garbage collector."""
    if getattr(sess, "_owned_by_solver", False):
        try:
            sess.close()
        except Exception:
            pass


def _unconstrained(engine, Hj, beta, y0, ftol, arn=None):
    if engine == "slsqp":
        return smallsolve.slsqp(Hj, beta, y0, (), ftol=ftol, tol=None)
    y = arn.ls_solution(Hj.shape[1]) if arn is not None else None
    if y is not None:
        return smallsolve.SmallResult(y)
    return smallsolve.lstsq(Hj, beta)


def _constrained(engine, Hj, beta, y0, cons, ftol, tol):
    if engine == "slsqp":
        return smallsolve.slsqp(Hj, beta, y0, cons, ftol=ftol, tol=tol)
    return smallsolve.kkt(Hj, beta, y0, cons)


def _warn_message(j, res):
    if not smallsolve.message_is_quiet(res.message):
        warnings.warn("Iteration %d failed with message '%s'" % (j, res.message), RuntimeWarning)


# ==============================================================================================
# FGMRES (solvers.py:58-127)
# ==============================================================================================
def gmres(A, b, x0, k, tol=1e-50, pre=None, *, session=None, lookahead=None, history=None,
          device=None, orth=None, small_solver=None):
    """Right-preconditioned flexible GMRES; returns (x_last, {'name','x','res','steps'})."""
    sess = _acquire(session, A, b, x0, k, (), pre, device, orth)
    try:
        return _gmres_run(sess, k, tol, lookahead, history)
    except BaseException:
        _abandon(sess)
        raise


def _gmres_run(sess, k, tol, lookahead, history):
    beta = sess.begin()                                   # r0, ||r0||, q0       (solvers.py:78-88)
    if _use_pipeline(sess, _opt("lookahead", lookahead), "kkt"):
        return _gmres_pipelined(sess, k, tol, beta, _opt("history", history))
    arn = _Arnoldi(sess, k, _opt("lookahead", lookahead), beta)
    hist = IterateHistory(sess)
    residual = [beta]
    steps = 0
    x_last = None
    for j in range(k):
        steps = j + 1
        col = arn.column(j)                               # (solvers.py:94-100)
        if not col[j + 1] != 0:
            warnings.warn(_BREAKDOWN)                     # (solvers.py:104-106)
            break
        if not arn.ls_residual() < tol:                   # this step cannot be the last: run ahead
            arn.prefetch(j + 1)
        yk = smallsolve.lstsq(arn.H[: j + 2, : j + 1], beta).x        # (solvers.py:113)
        residual.append(arn.residual(yk, j, arn.ls_residual() < tol, tol))   # x_j and ||A x_j - b|| (solvers.py:115-116)
        hist._append(yk)
        if residual[-1] < tol:
            break
    arn.drain()
    if len(hist) > 1:
        x_last = sess.ctx.download(nat.VEC_X, pinned=True)
    else:
        x_last = hist[0]
    info = {"name": "gmres",
            "x": _finish_history(hist, _opt("history", history), x_last),
            "res": residual[1:],
            "steps": steps}
    return x_last, info


def _gmres_pipelined(sess, k, tol, beta, history_mode):
    """gmres on the device-resident loop: y_j is the Givens least-squares solution (what np.linalg.lstsq returns for
    a full-rank Hessenberg matrix, solvers.py:113) computed on the GPU; a vanishing pivot hands it back to lstsq."""
    pipe = _Pipeline(sess, k, beta, tol, tol, last_unconstrained=k - 1)
    hist = IterateHistory(sess)
    residual = [beta]
    steps = 0
    yk = None
    for j in range(k):
        steps = j + 1
        col = pipe.column(j)                              # (solvers.py:94-100)
        if not col[j + 1] != 0:
            warnings.warn(_BREAKDOWN)                     # (solvers.py:104-106)
            break
        yk = pipe.ls_solution(j + 1)
        if yk is None:
            yk = smallsolve.lstsq(pipe.H[: j + 2, : j + 1], beta).x        # (solvers.py:113)
            res = pipe.host_iterate_residual(j, yk)
        else:
            res, _go = pipe.device_iterate_residual(j, yk, cannot_end=False)
        residual.append(res)                              # (solvers.py:115-116)
        hist._append(yk)
        if residual[-1] < tol:
            break
    pipe.finish(len(hist) - 2 if len(hist) > 1 else None, hist._ys[-1] if len(hist) > 1 else None)
    x_last = sess.ctx.download(nat.VEC_X, pinned=True) if len(hist) > 1 else hist[0]
    info = {"name": "gmres",
            "x": _finish_history(hist, history_mode, x_last),
            "res": residual[1:],
            "steps": steps}
    return x_last, info


# ==============================================================================================
# CGMRES (solvers.py:131-323)
# ==============================================================================================
def cgmres(A, b, x0, k, tol=1e-8, contol=10, conlist=[], pre=None, timing=None, *,
           session=None, small_solver=None, lookahead=None, history=None, device=None, orth=None):
    """Conservative FGMRES: unconstrained until residual <= contol*tol, then the reduced
    quadratic invariants are imposed as equality constraints on the Krylov coefficients."""
    engine = _opt("small_solver", small_solver)
    if engine not in ("slsqp", "kkt"):
        raise ValueError("small_solver must be 'slsqp' or 'kkt'")
    jit = None
    if timing:
        jit = {"start": time(), "start_iter": [], "end_iter": [],
               "start_constraints": [], "end_constraints": []}
    tr = _Trace()
    sess = _acquire(session, A, b, x0, k, conlist, pre, device, orth)
    try:
        return _cgmres_run(sess, k, tol, contol, timing, jit, engine, lookahead, history, tr)
    except BaseException:
        _abandon(sess)
        raise


def _cgmres_run(sess, k, tol, contol, timing, jit, engine, lookahead, history, tr):
    ctol = 1e-12                                          # (solvers.py:138)
    safety = None                                         # (solvers.py:163)
    beta = sess.begin()
    if _use_pipeline(sess, _opt("lookahead", lookahead), engine):
        return _cgmres_pipelined(sess, k, tol, contol, beta, timing, jit, _opt("history", history), tr)
    arn = _Arnoldi(sess, k, _opt("lookahead", lookahead), beta)
    tr("cgmres: session + r0")
    hist = IterateHistory(sess)
    residual = [beta]
    constrained_steps = 0
    steps = 0
    yk = None
    bk = _Buckets()
    for j in range(k):
        if timing:
            jit["start_iter"].append(time())
        steps = j + 1
        bk.mark("host")
        col = arn.column(j)
        bk.mark("arnoldi wait")
        if not col[j + 1] != 0:
            warnings.warn(_BREAKDOWN)                     # (solvers.py:199-202)
            break
        Hj = arn.H[: j + 2, : j + 1]
        y0 = np.zeros(j + 1)
        if j != 0:
            y0[:-1] = yk                                  # warm start (solvers.py:225-227)
        if residual[-1] > contol * tol and j < k - 1 and safety is None:      # (solvers.py:230)
            arn.prefetch(j + 1)
            res = _unconstrained(engine, Hj, beta, y0, ctol ** 2, arn)
        else:
            try:
                if timing:
                    constrained_steps += 1
                    jit["start_constraints"].append(time())
                bk.mark("host")
                cons = sess.containers(j + 1)             # (solvers.py:242-247)
                bk.mark("constraint terms")
                if timing:
                    jit["end_constraints"].append(time())
                # The loop can only end in this branch (solvers.py:296).  If even the unconstrained
                # minimiser is above tol it will not end now, so the next Arnoldi step is launched
                # before the host solve; otherwise wait for y and launch only if the residual it
                # implies says the loop goes on (a wrong guess costs overlap, never correctness).
                may_end = arn.ls_residual() < tol
                if not may_end:
                    arn.prefetch(j + 1)
                res = _constrained(engine, Hj, beta, y0, cons, ctol ** 2, None)   # (solvers.py:251-255)
                if may_end and not arn.predicted_residual(res.x, beta) < tol:
                    arn.prefetch(j + 1)
                if not timing and np.isnan(max(res.x)):
                    raise ValueError("constrained solve returned NaN")           # (solvers.py:258-260)
                safety = True
                if not timing:
                    dev = constraint_checker(res.x, [c.as_scipy() for c in cons])
                    if dev > ctol:
                        # The reference sets safety=False and then dies on a missing attribute
                        # inside its try block, i.e. lands in the unconstrained fallback
                        # (solvers.py:266-278).  Same outcome here.
                        safety = False
                        raise RuntimeError("Iteration %d failed to preserve constraints with "
                                           "deviation of %e" % (j, dev))
            except (nat.SpisError, nat.NativeLibraryError):
                raise                                     # device failures are never swallowed
            except Exception:
                warnings.warn("Constrained solve failed, defaulted to standard solve for iteration %d."
                              " Problem likely overconstrained, a smaller solver tolerance may be "
                              "required." % j, RuntimeWarning)
                if timing and len(jit["end_constraints"]) < len(jit["start_constraints"]):
                    jit["end_constraints"].append(time())
                arn.prefetch(j + 1)
                res = _unconstrained(engine, Hj, beta, y0, ctol ** 2, arn)         # (solvers.py:274-278)
        _warn_message(j, res)
        yk = res.x
        bk.mark("small solve + host")
        residual.append(arn.residual(yk, j, arn.predicted_residual(yk, beta) < tol, tol))   # (solvers.py:287,290)
        bk.mark("iterate+residual")
        hist._append(yk)
        if timing:
            jit["end_iter"].append(time())
        if residual[-1] < tol and safety is True:         # (solvers.py:296-297)
            break
    arn.drain()
    bk.report("cgmres")
    tr("cgmres: Krylov loop (%d steps)" % steps)
    if timing:                                            # (solvers.py:300-312)
        jit["end"] = time()
        iter_time = np.asarray(jit["end_iter"]) - np.asarray(jit["start_iter"][: len(jit["end_iter"])])
        iter_unconstrained = iter_time[:-constrained_steps]
        assembly = np.asarray(jit["end_constraints"]) - np.asarray(jit["start_constraints"])
        iter_constrained = iter_time[len(iter_unconstrained):] - assembly
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            timings = {"runtime": jit["end"] - jit["start"],
                       "iter_time_unconstrained": np.mean(iter_unconstrained),
                       "iter_time_constrained": np.mean(iter_constrained),
                       "constraint_building": np.mean(assembly),
                       "constrained_steps": constrained_steps}
    else:
        timings = None
    if len(hist) > 1:
        x_last = sess.ctx.download(nat.VEC_X, pinned=True)
    else:
        x_last = hist[0]
    tr("cgmres: download x")
    info = {"name": "cgmres",
            "x": _finish_history(hist, _opt("history", history), x_last),
            "res": residual[1:],
            "steps": steps,
            "timings": timings}
    return x_last, info


def _cgmres_timings(jit, constrained_steps):
    """The reference's timing dictionary (solvers.py:300-312)."""
    jit["end"] = time()
    iter_time = np.asarray(jit["end_iter"]) - np.asarray(jit["start_iter"][: len(jit["end_iter"])])
    iter_unconstrained = iter_time[:-constrained_steps]
    assembly = np.asarray(jit["end_constraints"]) - np.asarray(jit["start_constraints"])
    iter_constrained = iter_time[len(iter_unconstrained):] - assembly
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return {"runtime": jit["end"] - jit["start"],
                "iter_time_unconstrained": np.mean(iter_unconstrained),
                "iter_time_constrained": np.mean(iter_constrained),
                "constraint_building": np.mean(assembly),
                "constrained_steps": constrained_steps}


def _cgmres_pipelined(sess, k, tol, contol, beta, timing, jit, history_mode, tr):
    """cgmres (solvers.py:186-297) on the device-resident loop, small_solver='kkt'.  The control flow is the
    reference's; what changes is who computes what in an UNCONSTRAINED iteration: the device (Givens update, y_j, x_j,
    the residual norm, and the `residual[-1] > contol*tol` test for the steps it has already been given), while this
    loop follows through the records and owns every constrained step."""
    ctol = 1e-12                                          # (solvers.py:138)
    thr = contol * tol
    pipe = _Pipeline(sess, k, beta, thr, tol, last_unconstrained=k - 2)
    tr("cgmres: session + r0")
    hist = IterateHistory(sess)
    residual = [beta]
    go = beta > thr                                       # `residual[-1] > contol*tol` for the next iteration
    safety = None                                         # (solvers.py:163)
    constrained_steps = 0
    steps = 0
    yk = None
    bk = _Buckets()
    for j in range(k):
        if timing:
            jit["start_iter"].append(time())
        steps = j + 1
        bk.mark("host")
        col = pipe.column(j)
        bk.mark("arnoldi wait")
        if not col[j + 1] != 0:
            warnings.warn(_BREAKDOWN)                     # (solvers.py:199-202)
            break
        Hj = pipe.H[: j + 2, : j + 1]

        def unconstrained():
            y = pipe.ls_solution(j + 1)
            return smallsolve.SmallResult(y) if y is not None else smallsolve.lstsq(Hj, beta)

        if go and j < k - 1 and safety is None:           # (solvers.py:230)
            y_dev = pipe.ls_solution(j + 1)               # the device's least-squares coefficients (None: pivot trouble)
            bk.mark("small solve + host")
            if y_dev is not None:
                yk = y_dev
                r, go = pipe.device_iterate_residual(j, yk, cannot_end=True)
            else:
                yk = smallsolve.lstsq(Hj, beta).x
                r = pipe.host_iterate_residual(j, yk)
                go = r > thr
            residual.append(r)                            # (solvers.py:287,290)
        else:
            pipe.host_takes_over()
            y0 = np.zeros(j + 1)
            if j != 0:
                y0[:-1] = yk                              # warm start (solvers.py:225-227)
            likely_last = False
            try:
                if timing:
                    constrained_steps += 1
                    jit["start_constraints"].append(time())
                bk.mark("host")
                cons = sess.containers(j + 1)             # (solvers.py:242-247)
                bk.mark("constraint terms")
                if timing:
                    jit["end_constraints"].append(time())
                # the loop can only end in this branch (solvers.py:296): see cgmres above
                may_end = pipe.ls_residual(j) < tol
                if not may_end:
                    pipe.prefetch(j + 1)
                res = _constrained("kkt", Hj, beta, y0, cons, ctol ** 2, None)   # (solvers.py:251-255)
                likely_last = bool(may_end and pipe.predicted_residual(res.x) < tol)
                if may_end and not likely_last:
                    pipe.prefetch(j + 1)
                if not timing and np.isnan(max(res.x)):
                    raise ValueError("constrained solve returned NaN")           # (solvers.py:258-260)
                safety = True
                if not timing:
                    dev = constraint_checker(res.x, [c.as_scipy() for c in cons])
                    if dev > ctol:
                        safety = False                    # (solvers.py:266-278, quirk Q4: unconstrained fallback)
                        raise RuntimeError("Iteration %d failed to preserve constraints with "
                                           "deviation of %e" % (j, dev))
            except (nat.SpisError, nat.NativeLibraryError):
                raise                                     # device failures are never swallowed
            except Exception:
                warnings.warn("Constrained solve failed, defaulted to standard solve for iteration %d."
                              " Problem likely overconstrained, a smaller solver tolerance may be "
                              "required." % j, RuntimeWarning)
                if timing and len(jit["end_constraints"]) < len(jit["start_constraints"]):
                    jit["end_constraints"].append(time())
                pipe.prefetch(j + 1)
                likely_last = False
                res = unconstrained()                     # (solvers.py:274-278)
            _warn_message(j, res)
            yk = res.x
            bk.mark("small solve + host")
            residual.append(pipe.host_iterate_residual(j, yk, likely_last and safety is True))   # (solvers.py:287,290)
            go = residual[-1] > thr
        bk.mark("iterate+residual")
        hist._append(yk)
        if timing:
            jit["end_iter"].append(time())
        if residual[-1] < tol and safety is True:         # (solvers.py:296-297)
            break
    j_last = len(hist) - 2 if len(hist) > 1 else None
    pipe.finish(j_last, hist._ys[-1] if len(hist) > 1 else None)
    bk.report("cgmres")
    tr("cgmres: Krylov loop (%d steps)" % steps)
    timings = _cgmres_timings(jit, constrained_steps) if timing else None
    x_last = pipe.take_download(j_last)
    if x_last is None:
        x_last = sess.ctx.download(nat.VEC_X, pinned=True) if len(hist) > 1 else hist[0]
    tr("cgmres: download x")
    info = {"name": "cgmres",
            "x": _finish_history(hist, history_mode, x_last),
            "res": residual[1:],
            "steps": steps,
            "timings": timings}
    return x_last, info


# ==============================================================================================
# prototypical CGMRES (solvers.py:328-445)
# ==============================================================================================
def cgmres_p(A, b, x0, k, conlist=[], pre=None, *, session=None, small_solver=None,
             lookahead=None, history=None, device=None, orth=None):
    """Constraints are switched on one per iteration (clist[:j]); always runs k iterations."""
    engine = _opt("small_solver", small_solver)
    if engine not in ("slsqp", "kkt"):
        raise ValueError("small_solver must be 'slsqp' or 'kkt'")
    sess = _acquire(session, A, b, x0, k, conlist, pre, device, orth)
    try:
        return _cgmres_p_run(sess, k, engine, lookahead, history)
    except BaseException:
        _abandon(sess)
        raise


def _cgmres_p_run(sess, k, engine, lookahead, history):
    arn = _Arnoldi(sess, k, _opt("lookahead", lookahead))
    beta = sess.begin()
    hist = IterateHistory(sess)
    residual = []                                         # no initial residual (solvers.py:352)
    yk = None
    for j in range(k):
        col = arn.column(j)                               # no break on breakdown (solvers.py:376-377)
        Hj = arn.H[: j + 2, : j + 1]
        cons = sess.containers(j + 1)                     # all constraints, every step (solvers.py:397-401)
        if col[j + 1] != 0:
            arn.prefetch(j + 1)
        y0 = np.zeros(j + 1)
        if j != 0:
            y0[:-1] = yk
        res = _constrained(engine, Hj, beta, y0, cons[:j], 1e-20, 1e-15)      # (solvers.py:411-415)
        if np.isnan(max(res.x)):                          # (solvers.py:418-424)
            warnings.warn("Constrained solve silently failed on iteration %d" % j)
            res = _unconstrained(engine, Hj, beta, y0, 1e-20)
        _warn_message(j, res)
        yk = res.x
        residual.append(arn.residual(yk, j, False))       # (solvers.py:434-437)
        hist._append(yk)
    arn.drain()
    x_last = sess.ctx.download(nat.VEC_X, pinned=True) if len(hist) > 1 else hist[0]
    info = {"name": "geosolve",
            "x": _finish_history(hist, _opt("history", history), x_last),
            "res": residual}
    return x_last, info
