"""KrylovContext: Python owner of one `spis_ctx` (one linear system resident on one B200).

Every method is a direct call into libspis_b200.so; arrays crossing this boundary are host
numpy arrays.  See include/spis_b200.h for the contract of each entry point and the
reference lines (solvers.py:<line>) it replaces.
"""
from __future__ import annotations

import ctypes as C
import numpy as np
import scipy.sparse as sps

from . import _native as nat


def _csr_arrays(A):
    """Return (csr, indptr int32, indices int32, data f64) views suitable for the C ABI."""
    if not sps.issparse(A):
        A = sps.csr_matrix(np.asarray(A, dtype=np.float64))
    A = A.tocsr()
    if A.nnz >= 2**31 - 1:
        raise ValueError("matrices with nnz >= 2^31 must be row-sharded over several GPUs")
    indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    return A, indptr, indices, data


class KrylovContext:
    """One GPU-resident Krylov workspace: A, b, x0, basis q/z, constraint matrices."""

    def __init__(self, n: int, k_max: int, device: int = 0, n_halo: int = 0, stream=None):
        self._lib = nat.load_library()
        self._h = C.c_void_p()
        self.n, self.k_max, self.device, self.n_halo = int(n), int(k_max), int(device), int(n_halo)
        rc = self._lib.spis_ctx_create(self.device, self.n, self.n_halo, self.k_max,
                                       C.c_void_p(stream) if stream else None, C.byref(self._h))
        if rc != nat.OK:
            self._h = C.c_void_p()
            raise nat.SpisError(rc, self._lib.spis_last_global_error().decode())
        self._callbacks = None   # keep CFUNCTYPE objects alive
        self.generation = 0

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.spis_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):  # pragma: no cover - interpreter shutdown ordering
        try:
            self.close()
        except Exception:
            pass

    @property
    def closed(self) -> bool:
        return not self._h

    def _check(self, rc: int):
        if rc != nat.OK:
            raise nat.SpisError(rc, self._lib.spis_last_error(self._h).decode())

    def _live(self):
        if not self._h:
            raise RuntimeError("KrylovContext is closed")

    # -- options / info -----------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        self._live()
        self._check(self._lib.spis_set_option(self._h, key.encode(), int(value)))

    def info(self, key: str) -> int:
        self._live()
        out = C.c_int64(0)
        self._check(self._lib.spis_get_info(self._h, key.encode(), C.byref(out)))
        return out.value

    # -- uploads --------------------------------------------------------------------------------
    def upload_matrix(self, slot: int, A):
        self._live()
        A, indptr, indices, data = _csr_arrays(A)
        if A.shape[0] != self.n:
            raise ValueError(f"matrix has {A.shape[0]} rows, context owns {self.n}")
        self._check(self._lib.spis_upload_csr(self._h, slot, A.shape[0], A.shape[1], A.nnz,
                                              nat.iptr(indptr), nat.iptr(indices), nat.dptr(data)))

    def upload_vec(self, which: int, v):
        self._live()
        v = nat.as_f64(v, self.n)
        self._check(self._lib.spis_upload_vec(self._h, which, nat.dptr(v), self.n))

    def upload_blocks(self, blocks: np.ndarray, stride_block: int, stride_field: int):
        self._live()
        blocks = np.ascontiguousarray(blocks, dtype=np.float64)
        nblk, bs, bs2 = blocks.shape
        if bs != bs2:
            raise ValueError("blocks must be (nblk, bs, bs)")
        self._check(self._lib.spis_upload_blocks(self._h, bs, nblk, stride_block, stride_field, nat.dptr(blocks)))

    def any_nonzero(self, a) -> bool:
        """`np.any(a)` for a float64 buffer: page-locked arrays are scanned by the GPU over PCIe."""
        self._live()
        a = np.asarray(a)
        if a.dtype != np.float64 or not a.flags.c_contiguous or a.size < (1 << 16):
            return bool(np.any(a))
        out = C.c_int(0)
        self._check(self._lib.spis_any_nonzero(self._h, nat.dptr(a), a.size, C.byref(out)))
        return bool(out.value)

    def use_aux_stream(self, on: bool):
        """Uploads issued by the CALLING THREAD go to the auxiliary stream (see spis_b200.h)."""
        self._live()
        self._check(self._lib.spis_thread_use_aux_stream(self._h, 1 if on else 0))

    def set_precond(self, kind: int):
        self._live()
        self._check(self._lib.spis_set_precond(self._h, kind))

    # -- Krylov loop ------------------------------------------------------------------------------
    def solve_begin(self) -> float:
        self._live()
        beta = C.c_double(0.0)
        self._check(self._lib.spis_solve_begin(self._h, C.byref(beta)))
        self.generation += 1
        return beta.value

    def arnoldi_launch(self, j: int):
        self._check(self._lib.spis_arnoldi_launch(self._h, j))

    def arnoldi_begin(self, j: int):
        self._check(self._lib.spis_arnoldi_begin(self._h, j))

    def arnoldi_begin_residual(self, j: int):
        """arnoldi_begin(j) whose SpMV also measures ||A x - b|| of the iterate in X (iterate_residual_wait collects it)."""
        self._check(self._lib.spis_arnoldi_begin_residual(self._h, j))

    def arnoldi_finish(self, j: int, y_iterate=None):
        if y_iterate is None:
            self._check(self._lib.spis_arnoldi_finish(self._h, j, 0, None))
        else:
            y = nat.as_f64(y_iterate)
            self._check(self._lib.spis_arnoldi_finish(self._h, j, y.size, nat.dptr(y)))

    def residual_launch(self):
        self._check(self._lib.spis_residual_launch(self._h))

    def arnoldi_wait(self, j: int) -> np.ndarray:
        out = np.empty(j + 2, dtype=np.float64)
        self._check(self._lib.spis_arnoldi_wait(self._h, j, nat.dptr(out)))
        return out

    def arnoldi_step(self, j: int) -> np.ndarray:
        out = np.empty(j + 2, dtype=np.float64)
        self._check(self._lib.spis_arnoldi_step(self._h, j, nat.dptr(out)))
        return out

    def iterate_residual(self, y) -> float:
        y = nat.as_f64(y)
        res = C.c_double(0.0)
        self._check(self._lib.spis_iterate_residual(self._h, y.size, nat.dptr(y), C.byref(res)))
        return res.value

    def iterate_residual_launch(self, y):
        y = nat.as_f64(y)
        self._check(self._lib.spis_iterate_residual_launch(self._h, y.size, nat.dptr(y)))

    def iterate_residual_launch_dl(self, y, chunks: int = 4):
        """iterate_residual_launch for a final candidate: returns the page-locked array x_j is being streamed into
        (valid after download_join), or None when no page-locked buffer was available (plain launch)."""
        y = nat.as_f64(y)
        buf = nat.pinned_empty(self.n)
        started = C.c_int(0)
        self._check(self._lib.spis_iterate_residual_launch_dl(self._h, y.size, nat.dptr(y), nat.dptr(buf), int(chunks), C.byref(started)))
        return buf if started.value else None

    def download_join(self):
        self._check(self._lib.spis_download_join(self._h))

    def iterate_residual_wait(self) -> float:
        res = C.c_double(0.0)
        self._check(self._lib.spis_iterate_residual_wait(self._h, C.byref(res)))
        return res.value

    def form_iterate(self, y):
        y = nat.as_f64(y)
        self._check(self._lib.spis_form_iterate(self._h, y.size, nat.dptr(y)))

    # -- pipelined loop (include/spis_b200.h: spis_pipe_begin ...) ----------------------------------
    def pipe_begin(self, thr: float, phase0: bool):
        """Device-side Givens state and phase word for this solve; thr = residual norm at or below which the
        device stops forming unconstrained iterates (contol*tol for cgmres, tol for gmres)."""
        self._check(self._lib.spis_pipe_begin(self._h, float(thr), 1 if phase0 else 0))

    def step_enqueue(self, j: int, want_residual: bool, want_iterate: bool) -> int:
        """Queue Arnoldi step j; returns the ticket of its residual measurement (-1 without one)."""
        t = C.c_int64(-1)
        self._check(self._lib.spis_step_enqueue(self._h, j, (1 if want_residual else 0) | (2 if want_iterate else 0), C.byref(t)))
        return t.value

    def step_wait(self, j: int):
        """(h[:j+2, j], y_ls (j+1), info) once step j's record is there; info: valid, ls, h2next, nw2, s2, phase."""
        col = np.empty(j + 2, dtype=np.float64)
        y = np.empty(j + 1, dtype=np.float64)
        info = np.empty(8, dtype=np.float64)
        self._check(self._lib.spis_step_wait(self._h, j, nat.dptr(col), nat.dptr(y), nat.dptr(info)))
        return col, y, {"valid": bool(info[1]), "ls": float(info[2]), "norm2": float(info[3]), "nw2": float(info[4]),
                        "s2": float(info[5]), "phase": int(info[6])}

    def resid_wait(self, ticket: int):
        """(||A x - b||, its square as reduced on the device, go) of a residual measurement queued by step_enqueue;
        go = the device is still forming iterates."""
        res2 = C.c_double(0.0)
        go = C.c_int(0)
        self._check(self._lib.spis_resid_wait(self._h, int(ticket), C.byref(res2), C.byref(go)))
        return float(np.sqrt(res2.value)), res2.value, bool(go.value)

    # -- constraints ------------------------------------------------------------------------------
    def constraint_define(self, c: int, mat_slot: int, v, cc: float):
        self._live()
        vp = None
        if v is not None:
            v = nat.as_f64(v, self.n)
            vp = nat.dptr(v)
        self._check(self._lib.spis_constraint_define(self._h, c, mat_slot, vp, float(cc)))

    def constraint_set_constant(self, c: int, cc: float):
        self._live()
        self._check(self._lib.spis_constraint_set_constant(self._h, c, float(cc)))

    def constraint_set_vector(self, c: int, v):
        self._live()
        if v is None:
            self._check(self._lib.spis_constraint_set_vector(self._h, c, None))
        else:
            v = nat.as_f64(v, self.n)
            self._check(self._lib.spis_constraint_set_vector(self._h, c, nat.dptr(v)))

    def constraint_setup_async(self, c: int, M, v, cc: float, m_is_zero=None, v_is_zero=None):
        """Stage class-form constraint c (sparse M, vector v, scalar cc) on a native helper thread; the arrays
        handed to the library are kept alive here until constraint_setup_wait().  m_is_zero / v_is_zero: answers a
        row-sharded session has already agreed on with the other ranks (None: the library scans)."""
        self._live()
        M, indptr, indices, data = _csr_arrays(M)
        if M.shape[0] != self.n:
            raise ValueError(f"constraint matrix has {M.shape[0]} rows, context owns {self.n}")
        v = None if v_is_zero else nat.as_f64(v, self.n)
        self._async_refs = getattr(self, "_async_refs", [])
        self._async_refs.append((indptr, indices, data, v))
        self._check(self._lib.spis_constraint_setup_async2(self._h, c, M.shape[0], M.shape[1], M.nnz, nat.iptr(indptr),
                                                           nat.iptr(indices), nat.dptr(data), nat.dptr(v) if v is not None else None,
                                                           float(cc), -1 if m_is_zero is None else int(bool(m_is_zero))))

    def constraint_setup_wait(self):
        self._live()
        try:
            self._check(self._lib.spis_constraint_setup_wait(self._h))
        finally:
            self._async_refs = []

    def constraint_terms(self, c: int, m: int):
        t0 = C.c_double(0.0)
        t1 = np.empty(m, dtype=np.float64)
        t2 = np.empty((m, m), dtype=np.float64)
        self._check(self._lib.spis_constraint_terms(self._h, c, m, C.byref(t0), nat.dptr(t1), nat.dptr(t2)))
        return t0.value, t1, t2

    def constraint_terms_batch(self, cs, m: int):
        """[(term0, term1, term2)] for the constraints cs, one device round trip when each is one column behind m."""
        cs = np.ascontiguousarray(cs, dtype=np.int32)
        nc = cs.size
        t0 = np.empty(nc, dtype=np.float64)
        t1 = np.empty((nc, m), dtype=np.float64)
        t2 = np.empty((nc, m, m), dtype=np.float64)
        self._check(self._lib.spis_constraint_terms_batch(self._h, nc, nat.iptr(cs), m, nat.dptr(t0), nat.dptr(t1), nat.dptr(t2)))
        return [(float(t0[i]), t1[i], t2[i]) for i in range(nc)]

    # -- downloads / bridges ------------------------------------------------------------------------
    def download(self, which: int, j: int = 0, pinned: bool = False) -> np.ndarray:
        self._live()
        out = nat.pinned_empty(self.n) if pinned else np.empty(self.n, dtype=np.float64)
        self._check(self._lib.spis_download_vec(self._h, which, j, nat.dptr(out), self.n))
        return out

    def download_Z(self, j0: int, j1: int) -> np.ndarray:
        self._live()
        out = np.empty((j1 - j0, self.n), dtype=np.float64)
        self._check(self._lib.spis_download_Z(self._h, j0, j1, nat.dptr(out)))
        return out

    def host_pre_get(self, j: int) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        self._check(self._lib.spis_host_pre_get(self._h, j, nat.dptr(out)))
        return out

    def host_pre_put(self, j: int, z):
        z = nat.as_f64(z, self.n)
        self._check(self._lib.spis_host_pre_put(self._h, j, nat.dptr(z)))

    def set_collectives(self, allreduce, halo):
        """allreduce(device_ptr:int, count:int) sums a device buffer over all ranks in place;
        halo(send_ptr:int, recv_ptr:int) exchanges the packed ghost entries (see spis_b200.h)."""
        self._live()

        def _ar(_user, ptr, count):
            try:
                allreduce(ptr, count)
                return 0
            except Exception:  # pragma: no cover - surfaced as SpisError by the caller
                import traceback
                traceback.print_exc()
                return 1

        def _halo(_user, send_ptr, recv_ptr):
            try:
                halo(send_ptr, recv_ptr)
                return 0
            except Exception:  # pragma: no cover
                import traceback
                traceback.print_exc()
                return 1

        cb = (nat.ALLREDUCE_FN(_ar) if allreduce else nat.ALLREDUCE_FN(),
              nat.HALO_FN(_halo) if halo else nat.HALO_FN())
        self._callbacks = cb
        self._check(self._lib.spis_set_collectives(self._h, cb[0], cb[1], None))

    def halo_set_plan(self, send_idx):
        self._live()
        send_idx = np.ascontiguousarray(send_idx, dtype=np.int32)
        self._check(self._lib.spis_halo_set_plan(self._h, nat.iptr(send_idx), send_idx.size))

    def xcomm_create(self, rank: int, world: int, halo_cap: int) -> bytes:
        """Allocate this rank's NVLink comm buffer; returns its CUDA IPC handle (64 bytes)."""
        self._live()
        buf = C.create_string_buffer(64)
        self._check(self._lib.spis_xcomm_create(self._h, rank, world, int(halo_cap), buf, 64))
        return buf.raw

    def xcomm_connect(self, handles: bytes):
        self._live()
        self._check(self._lib.spis_xcomm_connect(self._h, C.c_char_p(handles)))

    def attach_comm(self, comm_handle):
        """Use a persistent NVLink communicator (spis_comm_create, owned by distributed.TorchComm)."""
        self._live()
        self._check(self._lib.spis_ctx_attach_comm(self._h, comm_handle))

    def xcomm_set_halo(self, dest_rank, dest_off, send_to, recv_from):
        self._live()
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (dest_rank, dest_off, send_to, recv_from)]
        self._check(self._lib.spis_xcomm_set_halo(self._h, nat.iptr(a[0]) if a[0].size else None,
                                                  nat.iptr(a[1]) if a[1].size else None, nat.iptr(a[2]), nat.iptr(a[3])))

    def xcomm_stats(self) -> dict:
        """Cycles spent waiting for peers inside fused reductions / halo exchanges since the last call (and their counts)."""
        self._live()
        out = (C.c_uint64 * 4)()
        self._check(self._lib.spis_xcomm_stats(self._h, out))
        return {"reduce_wait_cycles": int(out[0]), "reductions": int(out[1]), "halo_wait_cycles": int(out[2]), "halo_exchanges": int(out[3])}

    def sync(self):
        self._live()
        self._check(self._lib.spis_sync(self._h))

    # -- measurement --------------------------------------------------------------------------------
    def profile(self) -> dict:
        """{class: {'ms', 'bytes', 'launches', 'gbs'}} accumulated since reset_profile()."""
        self._live()
        ms = np.zeros(nat.PROF_CLASSES)
        by = np.zeros(nat.PROF_CLASSES)
        ln = np.zeros(nat.PROF_CLASSES, dtype=np.int64)
        self._check(self._lib.spis_get_profile(self._h, nat.dptr(ms), nat.dptr(by), ln.ctypes.data_as(C.POINTER(C.c_int64))))
        mv = np.zeros(nat.PROF_CLASSES)
        self._check(self._lib.spis_get_profile_moved(self._h, nat.dptr(mv)))
        gp = np.zeros(nat.PROF_CLASSES)
        self._check(self._lib.spis_get_profile_gaps(self._h, nat.dptr(gp)))
        out = {}
        for i, name in enumerate(nat.PROF_NAMES):
            out[name] = {"ms": float(ms[i]), "bytes": float(by[i]), "launches": int(ln[i]),
                         "gbs": float(by[i] / ms[i] * 1e-6) if ms[i] > 0 else None,
                         "moved_bytes": float(mv[i]), "gbs_moved": float(mv[i] / ms[i] * 1e-6) if ms[i] > 0 else None,
                         "idle_before_ms": float(gp[i])}
        return out

    def profile_trace(self):
        """Profile mode: [(class name, start ms, duration ms)] of the launches the last profile() call resolved."""
        self._live()
        n = C.c_int64(0)
        self._check(self._lib.spis_get_profile_trace(self._h, None, None, None, 0, C.byref(n)))
        cls = np.zeros(max(n.value, 1), dtype=np.int32)
        t0 = np.zeros(max(n.value, 1)); dt = np.zeros(max(n.value, 1))
        self._check(self._lib.spis_get_profile_trace(self._h, cls.ctypes.data_as(C.POINTER(C.c_int32)), nat.dptr(t0), nat.dptr(dt),
                                                     n.value, C.byref(n)))
        return [(nat.PROF_NAMES[int(c)], float(a), float(b)) for c, a, b in zip(cls[:n.value], t0[:n.value], dt[:n.value])]

    def timer_start(self):
        self._check(self._lib.spis_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double(0.0)
        self._check(self._lib.spis_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def reset_profile(self):
        self._live()
        self._check(self._lib.spis_reset_profile(self._h))

    # -- single-kernel entry points (tests / tuning) ----------------------------------------------
    def op_spmv(self, slot: int, x) -> np.ndarray:
        x = nat.as_f64(x)
        y = np.empty(self.n, dtype=np.float64)
        self._check(self._lib.spis_op_spmv(self._h, slot, nat.dptr(x), nat.dptr(y)))
        return y

    def op_mdot(self, V, w) -> np.ndarray:
        V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1, self.n) if np.size(V) else np.zeros((0, self.n))
        w = nat.as_f64(w, self.n)
        out = np.empty(V.shape[0] + 1, dtype=np.float64)
        self._check(self._lib.spis_op_mdot(self._h, V.shape[0], nat.dptr(V) if V.shape[0] else None, nat.dptr(w), nat.dptr(out)))
        return out

    def op_lincomb(self, V, base, coef, sign: float = 1.0, want_sumsq: bool = True):
        V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1, self.n) if np.size(V) else np.zeros((0, self.n))
        m = V.shape[0]
        coef = nat.as_f64(coef, m)
        out = np.empty(self.n, dtype=np.float64)
        ss = C.c_double(0.0)
        bp = None
        if base is not None:
            base = nat.as_f64(base, self.n)
            bp = nat.dptr(base)
        self._check(self._lib.spis_op_lincomb(self._h, m, nat.dptr(V) if m else None, bp, nat.dptr(coef) if m else None,
                                              float(sign), nat.dptr(out), C.byref(ss) if want_sumsq else None))
        return (out, ss.value) if want_sumsq else out

    def op_orth_mid(self, V, w, coef):
        """Fused middle of CGS2: returns (w - coef @ V, V @ (w - coef @ V)) from one pass over V."""
        V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1, self.n)
        m = V.shape[0]
        w = nat.as_f64(w, self.n)
        coef = nat.as_f64(coef, m)
        w_out = np.empty(self.n, dtype=np.float64)
        dots = np.empty(m, dtype=np.float64)
        self._check(self._lib.spis_op_orth_mid(self._h, m, nat.dptr(V), nat.dptr(w), nat.dptr(coef),
                                               nat.dptr(w_out), nat.dptr(dots)))
        return w_out, dots

    def op_precond(self, q) -> np.ndarray:
        q = nat.as_f64(q, self.n)
        z = np.empty(self.n, dtype=np.float64)
        self._check(self._lib.spis_op_precond(self._h, nat.dptr(q), nat.dptr(z)))
        return z

    def bench_kernel(self, prof_class: int, m: int, reps: int = 20):
        """(ms per launch, algorithmic bytes per launch) for one kernel class on resident data."""
        ms = C.c_double(0.0)
        by = C.c_double(0.0)
        self._check(self._lib.spis_bench_kernel(self._h, prof_class, m, reps, C.byref(ms), C.byref(by)))
        return ms.value, by.value
