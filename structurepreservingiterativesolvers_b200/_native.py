"""ctypes binding of the C ABI in include/spis_b200.h (libspis_b200.so).

The binding is deliberately thin: one Python function per exported symbol, numpy arrays in,
numpy arrays out.  There is no CPU fallback -- if the shared library is missing, or no sm_100
device is visible, every entry point raises (`NativeLibraryError` / `SpisError`).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "lib", "libspis_b200.so")

# constants mirrored from include/spis_b200.h
ABI_VERSION = 5
OK, E_INVALID, E_CUDA, E_NOMEM, E_UNSUPPORTED = 0, -1, -2, -3, -4
SLOT_A, SLOT_PRE, SLOT_CON0, MAX_SLOTS = 0, 1, 2, 18
VEC_B, VEC_X0, VEC_R0, VEC_Q, VEC_Z, VEC_X, VEC_PRE_DIAG, VEC_W = range(8)
PRE_NONE, PRE_JACOBI, PRE_CSR, PRE_BLOCK, PRE_HOST = range(5)
ORTH_CGS2, ORTH_CGS1, ORTH_MGS = range(3)
FMT_AUTO, FMT_SELL, FMT_CSR, FMT_SELL2, FMT_PATTERN, FMT_SELLD = range(6)
PROF_SPMV, PROF_MDOT, PROF_LINCOMB, PROF_SCALE, PROF_PRECOND, PROF_OTHER, PROF_ORTHMID, PROF_SPMV_AUX, PROF_CLASSES = range(9)
PROF_NAMES = ("spmv", "mdot", "lincomb", "scale", "precond", "other", "orthmid", "spmv_aux")

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)
HALO_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_ctx = C.c_void_p

# name -> (restype, argtypes): every symbol declared in include/spis_b200.h
SIGNATURES = {
    "spis_abi_version": (C.c_int, []),
    "spis_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "spis_ctx_create": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.POINTER(_ctx)]),
    "spis_ctx_destroy": (C.c_int, [_ctx]),
    "spis_last_error": (C.c_char_p, [_ctx]),
    "spis_last_global_error": (C.c_char_p, []),
    "spis_pinned_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "spis_pinned_free": (C.c_int, [C.c_void_p]),
    "spis_pinned_trim": (C.c_int, []),
    "spis_device_trim": (C.c_int, []),
    "spis_host_any_nonzero": (C.c_int, [_dp, C.c_size_t, C.POINTER(C.c_int)]),
    "spis_any_nonzero": (C.c_int, [_ctx, _dp, C.c_size_t, C.POINTER(C.c_int)]),
    "spis_set_option": (C.c_int, [_ctx, C.c_char_p, C.c_int64]),
    "spis_get_info": (C.c_int, [_ctx, C.c_char_p, _lp]),
    "spis_upload_csr": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int64, C.c_int64, _ip, _ip, _dp]),
    "spis_upload_vec": (C.c_int, [_ctx, C.c_int, _dp, C.c_int64]),
    "spis_upload_blocks": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int64, C.c_int64, _dp]),
    "spis_set_precond": (C.c_int, [_ctx, C.c_int]),
    "spis_thread_use_aux_stream": (C.c_int, [_ctx, C.c_int]),
    "spis_solve_begin": (C.c_int, [_ctx, _dp]),
    "spis_arnoldi_launch": (C.c_int, [_ctx, C.c_int]),
    "spis_arnoldi_wait": (C.c_int, [_ctx, C.c_int, _dp]),
    "spis_arnoldi_step": (C.c_int, [_ctx, C.c_int, _dp]),
    "spis_arnoldi_begin": (C.c_int, [_ctx, C.c_int]),
    "spis_arnoldi_begin_residual": (C.c_int, [_ctx, C.c_int]),
    "spis_arnoldi_finish": (C.c_int, [_ctx, C.c_int, C.c_int, _dp]),
    "spis_residual_launch": (C.c_int, [_ctx]),
    "spis_iterate_residual": (C.c_int, [_ctx, C.c_int, _dp, _dp]),
    "spis_iterate_residual_launch": (C.c_int, [_ctx, C.c_int, _dp]),
    "spis_iterate_residual_launch_dl": (C.c_int, [_ctx, C.c_int, _dp, _dp, C.c_int, C.POINTER(C.c_int)]),
    "spis_download_join": (C.c_int, [_ctx]),
    "spis_iterate_residual_wait": (C.c_int, [_ctx, _dp]),
    "spis_form_iterate": (C.c_int, [_ctx, C.c_int, _dp]),
    "spis_pipe_begin": (C.c_int, [_ctx, C.c_double, C.c_int]),
    "spis_step_enqueue": (C.c_int, [_ctx, C.c_int, C.c_int, _lp]),
    "spis_step_wait": (C.c_int, [_ctx, C.c_int, _dp, _dp, _dp]),
    "spis_resid_wait": (C.c_int, [_ctx, C.c_int64, _dp, C.POINTER(C.c_int)]),
    "spis_constraint_define": (C.c_int, [_ctx, C.c_int, C.c_int, _dp, C.c_double]),
    "spis_constraint_terms": (C.c_int, [_ctx, C.c_int, C.c_int, _dp, _dp, _dp]),
    "spis_constraint_setup_async": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int64, C.c_int64, _ip, _ip, _dp, _dp, C.c_double]),
    "spis_constraint_setup_async2": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int64, C.c_int64, _ip, _ip, _dp, _dp, C.c_double, C.c_int]),
    "spis_constraint_setup_wait": (C.c_int, [_ctx]),
    "spis_constraint_set_constant": (C.c_int, [_ctx, C.c_int, C.c_double]),
    "spis_constraint_set_vector": (C.c_int, [_ctx, C.c_int, _dp]),
    "spis_h2d_bytes": (C.c_longlong, [C.c_int]),
    "spis_host_find_patterns": (C.c_int, [_ip, _ip, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_uint16), _ip,
                                          C.POINTER(C.c_int), C.POINTER(C.c_int), _lp]),
    "spis_constraint_terms_batch": (C.c_int, [_ctx, C.c_int, _ip, C.c_int, _dp, _dp, _dp]),
    "spis_small_settle": (C.c_int, [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.POINTER(C.c_int)]),
    "spis_small_kkt": (C.c_int, [C.c_int, C.c_int, _dp, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "spis_download_vec": (C.c_int, [_ctx, C.c_int, C.c_int, _dp, C.c_int64]),
    "spis_download_Z": (C.c_int, [_ctx, C.c_int, C.c_int, _dp]),
    "spis_host_pre_get": (C.c_int, [_ctx, C.c_int, _dp]),
    "spis_host_pre_put": (C.c_int, [_ctx, C.c_int, _dp]),
    "spis_set_collectives": (C.c_int, [_ctx, ALLREDUCE_FN, HALO_FN, C.c_void_p]),
    "spis_halo_set_plan": (C.c_int, [_ctx, _ip, C.c_int64]),
    "spis_xcomm_create": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64]),
    "spis_xcomm_connect": (C.c_int, [_ctx, C.c_void_p]),
    "spis_xcomm_set_halo": (C.c_int, [_ctx, _ip, _ip, _ip, _ip]),
    "spis_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "spis_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "spis_comm_destroy": (C.c_int, [C.c_void_p]),
    "spis_comm_capacity": (C.c_int, [C.c_void_p, _lp, _lp]),
    "spis_comm_allreduce": (C.c_int, [C.c_void_p, _dp, C.c_int]),
    "spis_ctx_attach_comm": (C.c_int, [_ctx, C.c_void_p]),
    "spis_xcomm_stats": (C.c_int, [_ctx, C.POINTER(C.c_uint64)]),
    "spis_sync": (C.c_int, [_ctx]),
    "spis_get_profile": (C.c_int, [_ctx, _dp, _dp, _lp]),
    "spis_reset_profile": (C.c_int, [_ctx]),
    "spis_get_profile_moved": (C.c_int, [_ctx, _dp]),
    "spis_get_profile_gaps": (C.c_int, [_ctx, _dp]),
    "spis_get_profile_trace": (C.c_int, [_ctx, _ip, _dp, _dp, C.c_int64, _lp]),
    "spis_timer_start": (C.c_int, [_ctx]),
    "spis_timer_stop": (C.c_int, [_ctx, _dp]),
    "spis_op_spmv": (C.c_int, [_ctx, C.c_int, _dp, _dp]),
    "spis_op_mdot": (C.c_int, [_ctx, C.c_int, _dp, _dp, _dp]),
    "spis_op_lincomb": (C.c_int, [_ctx, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp]),
    "spis_op_precond": (C.c_int, [_ctx, _dp, _dp]),
    "spis_op_orth_mid": (C.c_int, [_ctx, C.c_int, _dp, _dp, _dp, _dp, _dp]),
    "spis_bench_kernel": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, _dp, _dp]),
}


class NativeLibraryError(RuntimeError):
    """libspis_b200.so is missing or does not match this package (there is no CPU fallback)."""


class SpisError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[spis {code}] {message}")
        self.code = code


_lib = None


def load_library(path: str | None = None):
    """dlopen the shared library and attach the prototypes (idempotent)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SPIS_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise NativeLibraryError(
            f"{p} not found. Build it with `python -m structurepreservingiterativesolvers_b200.build` "
            "(nvcc, sm_100a). This package has no CPU fallback.")
    try:
        lib = C.CDLL(p)
    except OSError as exc:  # pragma: no cover - depends on the host
        raise NativeLibraryError(f"cannot load {p}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise NativeLibraryError(f"{p} does not export {name}; rebuild the library") from exc
        fn.restype = res
        fn.argtypes = args
    if lib.spis_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"{p} has ABI {lib.spis_abi_version()}, binding expects {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


def as_f64(a, n=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float64)
    if out.ndim != 1:
        out = out.reshape(-1)
    if n is not None and out.size != n:
        raise ValueError(f"expected a vector of length {n}, got {out.size}")
    return out


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def iptr(a: np.ndarray):
    return a.ctypes.data_as(_ip)


def device_count() -> int:
    lib = load_library()
    cnt = C.c_int(0)
    rc = lib.spis_device_count(C.byref(cnt))
    if rc != OK:
        raise SpisError(rc, lib.spis_last_global_error().decode())
    return cnt.value


class _PinnedOwner:
    """Owns one pooled page-locked buffer; exposes it to numpy through __array_interface__."""

    def __init__(self, lib, ptr, n):
        self._lib, self._ptr, self._bytes = lib, ptr, n * 8
        self.__array_interface__ = {"data": (ptr, False), "shape": (n,), "typestr": "<f8", "version": 3}

    def __del__(self):  # pragma: no cover - finaliser
        global _pinned_out
        try:
            self._lib.spis_pinned_free(C.c_void_p(self._ptr))
            with _pinned_lock:
                _pinned_out -= self._bytes
        except Exception:
            pass


# Page-locked result buffers are owned by the CALLER once returned (x_last of every solve): a caller that keeps
# the solution of every time step (the reference's Evolve `sol` list) must not pin unbounded host memory, so the
# bytes outstanding are counted and requests over the budget get ordinary pageable arrays.
_PINNED_LIMIT = 2 << 30          # bytes of result buffers handed out from the pool and still alive
_pinned_out = 0
_pinned_lock = threading.Lock()


def pinned_outstanding() -> int:
    """Bytes of page-locked result buffers currently owned by callers."""
    return _pinned_out


def pinned_empty(n: int) -> np.ndarray:
    """float64 vector in page-locked memory (falls back to a normal array over the budget)."""
    global _pinned_out
    lib = load_library()
    nbytes = n * 8
    with _pinned_lock:
        if _pinned_out + nbytes > _PINNED_LIMIT:
            return np.empty(n, dtype=np.float64)
        _pinned_out += nbytes
    ptr = C.c_void_p()
    if lib.spis_pinned_alloc(nbytes, C.byref(ptr)) != OK or not ptr.value:
        with _pinned_lock:
            _pinned_out -= nbytes
        return np.empty(n, dtype=np.float64)
    return np.asarray(_PinnedOwner(lib, ptr.value, n))


def any_nonzero(a) -> bool:
    """Fast `np.any(a)` for large float64 buffers (threads in the C library)."""
    a = np.asarray(a)
    if a.dtype != np.float64 or not a.flags.c_contiguous or a.size < (1 << 16):
        return bool(np.any(a))
    out = C.c_int(0)
    if load_library().spis_host_any_nonzero(dptr(a), a.size, C.byref(out)) != OK:
        return bool(np.any(a))
    return bool(out.value)
