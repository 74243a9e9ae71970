"""Kernel-level parity through the C ABI: each sm_100a kernel against numpy/scipy on seeded inputs.

fp64 tolerances: a dot product of n terms summed in a different order differs by <= ~n*eps*|a||b|;
the bounds below are 1e-13 relative to the natural scale of each quantity.
"""
import numpy as np
import pytest
import scipy.sparse as sps

from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.device import KrylovContext
from structurepreservingiterativesolvers_b200.preconditioners import BlockJacobiPreconditioner
from structurepreservingiterativesolvers_b200.problems import heat, lkdv

pytestmark = pytest.mark.gpu


def ragged_matrix(n, seed, max_len=40, empty_every=7):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, max_len, size=n)
    lens[::empty_every] = 0                       # empty rows
    lens[n // 2] = 3 * max_len                    # one long row
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    cols = rng.integers(0, n, size=indptr[-1])
    vals = rng.standard_normal(indptr[-1])
    return sps.csr_matrix((vals, cols, indptr), shape=(n, n))


@pytest.mark.parametrize("fmt", [nat.FMT_SELL, nat.FMT_SELL2, nat.FMT_CSR, nat.FMT_AUTO])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4097])
def test_spmv_ragged(n, fmt):
    A = ragged_matrix(n, seed=n) if n > 40 else sps.random(n, n, density=0.6, random_state=n, format="csr")
    rng = np.random.default_rng(1)
    x = rng.standard_normal(n)
    with KrylovContext(n, 2) as ctx:
        ctx.set_option("spmv_format", fmt)
        ctx.upload_matrix(nat.SLOT_A, A)
        y = ctx.op_spmv(nat.SLOT_A, x)
    ref = A @ x
    scale = np.abs(A) @ np.abs(x) + 1e-300
    assert np.max(np.abs(y - ref) / scale) <= 1e-14


def test_spmv_pattern_storage_detection_and_fallback():
    """spmv_format=auto stores a matrix as 16-bit stencil ids + a stencil table only when it has at most
    4096 distinct rows (verified entry by entry on the device); everything else keeps SELL / CSR."""
    rng = np.random.default_rng(11)
    # (a) uniform periodic stencil: one pattern for interior rows, wrap-around variants at the ends
    n = 20_000
    T = sps.diags([np.full(n - 1, -1.0), np.full(n, 2.5), np.full(n - 1, -1.25)], [-1, 0, 1], format="lil")
    T[0, n - 1] = -1.0; T[n - 1, 0] = -1.25
    T = T.tocsr()
    x = rng.standard_normal(n)
    with KrylovContext(n, 2) as ctx:
        ctx.upload_matrix(nat.SLOT_A, T)
        assert ctx.info(f"fmt:{nat.SLOT_A}") == nat.FMT_PATTERN and ctx.info(f"npat:{nat.SLOT_A}") == 3
        np.testing.assert_allclose(ctx.op_spmv(nat.SLOT_A, x), T @ x, rtol=1e-14, atol=1e-14)
    # (b) variable coefficients: every row its own pattern -> general storage, same answer
    V = sps.diags([rng.standard_normal(n - 1), rng.standard_normal(n), rng.standard_normal(n - 1)], [-1, 0, 1], format="csr")
    with KrylovContext(n, 2) as ctx:
        ctx.upload_matrix(nat.SLOT_A, V)
        assert ctx.info(f"fmt:{nat.SLOT_A}") in (nat.FMT_SELL, nat.FMT_CSR)
        np.testing.assert_allclose(ctx.op_spmv(nat.SLOT_A, x), V @ x, rtol=1e-13, atol=1e-13)
        ctx.set_option("spmv_format", nat.FMT_PATTERN)
        with pytest.raises(nat.SpisError):
            ctx.upload_matrix(nat.SLOT_A, V)
    # (c) values that differ in the last bit are different patterns (lossless or nothing)
    data = np.full(3 * n - 2, 1.0)
    W = sps.diags([data[: n - 1], data[: n], data[: n - 1]], [-1, 0, 1], format="csr")
    W.data[7] = np.nextafter(1.0, 2.0)
    with KrylovContext(n, 2) as ctx:
        ctx.upload_matrix(nat.SLOT_A, W)
        assert ctx.info(f"fmt:{nat.SLOT_A}") == nat.FMT_PATTERN and ctx.info(f"npat:{nat.SLOT_A}") == 4
        y = ctx.op_spmv(nat.SLOT_A, x)
    np.testing.assert_allclose(y, W @ x, rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize("case", ["tridiag", "lkdv", "perturbed", "swe"])
def test_row_patterns_found_on_the_host_equal_the_device_detection(case):
    """With option host_pattern = 1 spis_upload_csr looks for the row stencils in the caller's arrays with host threads
    first (matrices above host_pattern_min_nnz entries): the CSR arrays then never cross PCIe.  Same storage decision,
    same number of stencils and the same SpMV bits as the detection on the device (the default)."""
    from structurepreservingiterativesolvers_b200.problems import swe
    rng = np.random.default_rng(17)
    if case == "tridiag":
        n = 50_001
        T = sps.diags([np.full(n - 1, -1.0), np.full(n, 2.5), np.full(n - 1, -1.25)], [-1, 0, 1], format="lil")
        T[0, n - 1] = -1.0; T[n - 1, 0] = -1.25
        A = T.tocsr()
    elif case in ("lkdv", "perturbed"):
        A = lkdv.linforms(space="CG", M=33_350, mlength=0.8 * 33_350)[0]["A"].tocsr()
        if case == "perturbed":
            A = A.copy(); A.data = A.data * (1.0 + 1e-13 * rng.standard_normal(A.nnz))
    else:
        A = swe.linforms(M=40, mlength=32.0)[0]["A"].tocsr()       # column - row drifts from row to row: no patterns
    n = A.shape[0]
    x = rng.standard_normal(n)
    out = []
    for host in (1, 0):
        with KrylovContext(n, 2) as ctx:
            ctx.set_option("host_pattern", host)
            ctx.set_option("host_pattern_min_nnz", 0)
            ctx.set_option("host_threads", 3)
            ctx.upload_matrix(nat.SLOT_A, A)
            fmt = ctx.info(f"fmt:{nat.SLOT_A}")
            out.append((fmt, ctx.info(f"npat:{nat.SLOT_A}") if fmt == nat.FMT_PATTERN else -1, ctx.info(f"fw_fields:{nat.SLOT_A}"),
                        ctx.op_spmv(nat.SLOT_A, x)))
    assert out[0][:3] == out[1][:3]
    assert (out[0][0] == nat.FMT_PATTERN) == (case in ("tridiag", "lkdv"))
    np.testing.assert_array_equal(out[0][3], out[1][3])
    scale = np.abs(A) @ np.abs(x)
    assert np.max(np.abs(out[0][3] - A @ x) / scale) <= 1e-14


def test_spmv_value_dictionary_storage():
    """SELLD: 8-bit codes for the values when the matrix holds at most 256 distinct doubles (bit patterns)."""
    from structurepreservingiterativesolvers_b200.problems import swe
    d, _ = swe.linforms(M=40, mlength=32.0)                          # rows do not repeat as stencils, values do
    A = d["A"]
    n = A.shape[0]
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n)
    with KrylovContext(n, 2) as ctx:
        ctx.upload_matrix(nat.SLOT_A, A)                            # auto
        assert ctx.info(f"fmt:{nat.SLOT_A}") == nat.FMT_SELLD and 1 <= ctx.info(f"ndict:{nat.SLOT_A}") <= 256
        y = ctx.op_spmv(nat.SLOT_A, x)
        ctx.set_option("spmv_format", nat.FMT_SELL)
        ctx.upload_matrix(nat.SLOT_A, A)
        np.testing.assert_array_equal(y, ctx.op_spmv(nat.SLOT_A, x))   # same values, same order: same bits
    scale = np.abs(A) @ np.abs(x)
    assert np.max(np.abs(y - A @ x) / scale) <= 1e-14
    # 300 distinct values (incl. -0.0 / +0.0, which are different codes): refused when asked for explicitly
    B = sps.diags([np.arange(1.0, 301.0)], [0], format="csr")
    with KrylovContext(300, 2) as ctx:
        ctx.set_option("spmv_format", nat.FMT_SELLD)
        with pytest.raises(nat.SpisError):
            ctx.upload_matrix(nat.SLOT_A, B)
    Z = sps.csr_matrix((np.array([0.0, -0.0, 2.0, 0.0]), np.array([0, 1, 2, 3]), np.arange(5)), shape=(4, 4))
    with KrylovContext(4, 2) as ctx:
        ctx.set_option("spmv_format", nat.FMT_SELLD)
        ctx.upload_matrix(nat.SLOT_A, Z)
        assert ctx.info(f"ndict:{nat.SLOT_A}") == 3
        np.testing.assert_array_equal(ctx.op_spmv(nat.SLOT_A, np.array([1.0, 1.0, 1.0, np.pi])), Z @ np.array([1.0, 1.0, 1.0, np.pi]))


def test_spmv_duplicates_and_unsorted_indices():
    n = 257
    rng = np.random.default_rng(5)
    rows = rng.integers(0, n, 4000); cols = rng.integers(0, n, 4000); vals = rng.standard_normal(4000)
    A = sps.coo_matrix((vals, (rows, cols)), shape=(n, n))
    csr = sps.csr_matrix((vals, cols, np.searchsorted(np.sort(rows), np.arange(n + 1))), shape=(n, n))   # duplicates kept, unsorted
    csr = sps.csr_matrix((vals[np.argsort(rows, kind="stable")], cols[np.argsort(rows, kind="stable")], csr.indptr), shape=(n, n))
    x = rng.standard_normal(n)
    with KrylovContext(n, 2) as ctx:
        ctx.upload_matrix(nat.SLOT_A, csr)
        y = ctx.op_spmv(nat.SLOT_A, x)
    np.testing.assert_allclose(y, A @ x, rtol=0, atol=1e-12)


@pytest.mark.parametrize("fmt", [nat.FMT_SELL, nat.FMT_SELL2, nat.FMT_CSR, nat.FMT_PATTERN, nat.FMT_SELLD])
def test_spmv_fem_operators(fmt):
    d, _ = lkdv.linforms(space="CG", M=33_350, mlength=0.8 * 33_350)        # n = 100 050
    h, _ = heat.linforms(M=150)
    rng = np.random.default_rng(2)
    for A in (d["A"], d["L"] - d["M"], h["A"]):
        n = A.shape[0]
        x = rng.standard_normal(n)
        with KrylovContext(n, 2) as ctx:
            ctx.set_option("spmv_format", fmt)
            ctx.upload_matrix(nat.SLOT_A, A)
            y1 = ctx.op_spmv(nat.SLOT_A, x)
            y2 = ctx.op_spmv(nat.SLOT_A, x)
            assert ctx.info(f"fmt:{nat.SLOT_A}") == fmt
        np.testing.assert_array_equal(y1, y2)                      # deterministic
        scale = np.abs(A) @ np.abs(x) + 1e-300
        assert np.max(np.abs(y1 - A @ x) / scale) <= 1e-14


@pytest.mark.parametrize("fmt", [nat.FMT_SELL, nat.FMT_SELLD, nat.FMT_PATTERN, nat.FMT_CSR])
def test_spmv_dual_matches_the_two_separate_products(fmt):
    """spis_arnoldi_begin_residual: w = A q_1 and ||A x - b|| from ONE pass over A.  Same Hessenberg column bits as
    the separate SpMV (same per-row summation order), the norm to rounding (per-CTA partial sums grouped differently);
    formats without a dual kernel (CSR) take two launches behind the same entry point."""
    from structurepreservingiterativesolvers_b200.problems import swe
    rng = np.random.default_rng(23)
    A = (swe.linforms(M=40, mlength=32.0)[0]["A"] if fmt in (nat.FMT_SELLD, nat.FMT_CSR)
         else lkdv.linforms(space="CG", M=33_350, mlength=0.8 * 33_350)[0]["A"])
    n = A.shape[0]
    b, x0, y = rng.standard_normal(n), rng.standard_normal(n), np.array([0.37])
    out = []
    for dual in (1, 0):
        with KrylovContext(n, 4) as ctx:
            ctx.set_option("spmv_format", fmt)
            ctx.set_option("spmv_dual", dual)
            ctx.upload_matrix(nat.SLOT_A, A)
            ctx.upload_vec(nat.VEC_B, b)
            ctx.upload_vec(nat.VEC_X0, x0)
            beta = ctx.solve_begin()
            col0 = ctx.arnoldi_step(0)
            ctx.form_iterate(y)                                   # X = x0 + y[0] q_0
            x = ctx.download(nat.VEC_X)
            ctx.arnoldi_begin_residual(1)
            res = ctx.iterate_residual_wait()
            ctx.arnoldi_finish(1)
            col1 = ctx.arnoldi_wait(1)
            out.append((beta, col0, x, res, col1))
    ref = np.linalg.norm(A @ out[0][2] - b)
    for k in (0, 1):
        assert abs(out[k][3] - ref) <= 1e-13 * ref
    np.testing.assert_array_equal(out[0][2], out[1][2])
    np.testing.assert_array_equal(out[0][4], out[1][4])


def _modes_through_abi(ctx, A, b, x0):
    """(A x0 via mode 0, ||b - A x0|| via mode 1, ||A x0 - b|| via mode 2) of whatever kernel the context picks."""
    ctx.upload_vec(nat.VEC_B, b)
    ctx.upload_vec(nat.VEC_X0, x0)
    y = ctx.op_spmv(nat.SLOT_A, x0)
    beta = ctx.solve_begin()
    res = ctx.iterate_residual(np.zeros(0))
    return y, beta, res


@pytest.mark.parametrize("fmt", [nat.FMT_SELL, nat.FMT_SELLD, nat.FMT_PATTERN])
def test_spmv_kernel_generations_agree(fmt):
    """spmv_variant 0 (first-generation kernels) and 1 (id prefetch for row patterns, software-pipelined
    slices for SELL) walk every row in the same order: same bits in mode 0; the fused norms of modes 1 and 2
    differ only in how the per-CTA partial sums are grouped."""
    from structurepreservingiterativesolvers_b200.problems import swe
    rng = np.random.default_rng(17)
    mats = []
    if fmt != nat.FMT_PATTERN:
        mats.append(swe.linforms(M=40, mlength=32.0)[0]["A"])                     # widths 16 / 9, <= 256 distinct values
        R = ragged_matrix(4097, seed=3)
        R.data[:] = rng.integers(-5, 6, size=R.nnz).astype(np.float64)           # few distinct values: SELLD applies
        mats.append(R)
        E = sps.csr_matrix((3000, 3000)); E = E.tolil(); E[5, 7] = 2.0; E[2999, 0] = -1.0   # slices that are entirely empty
        mats.append(E.tocsr())
    mats.append(lkdv.linforms(space="CG", M=33_350, mlength=0.8 * 33_350)[0]["A"])
    mats.append(lkdv.linforms(space="DG", M=500)[0]["A"])
    for A in mats:
        n = A.shape[0]
        b, x0 = rng.standard_normal(n), rng.standard_normal(n)
        out = []
        for var in (0, 1):
            with KrylovContext(n, 2) as ctx:
                ctx.set_option("spmv_format", fmt)
                ctx.set_option("spmv_variant", var)
                ctx.upload_matrix(nat.SLOT_A, A)
                assert ctx.info(f"fmt:{nat.SLOT_A}") == fmt
                out.append(_modes_through_abi(ctx, A, b, x0))
        ref = np.linalg.norm(b - A @ x0)
        for y, beta, res in out:
            np.testing.assert_array_equal(y, out[0][0])
            assert abs(beta - ref) <= 1e-13 * ref and abs(res - ref) <= 1e-13 * ref



@pytest.mark.parametrize("n", [2, 1023, 1024, 1025, 2049, 300_007])
@pytest.mark.parametrize("m", [0, 1, 3, 4, 5, 21, 39])
@pytest.mark.parametrize("variant", [0, 1])            # 0: chosen by size; 1: per-thread register sums (mdot_reg_kernel)
def test_mdot(n, m, variant):
    rng = np.random.default_rng(n + m)
    V = rng.standard_normal((m, n))
    w = rng.standard_normal(n)
    with KrylovContext(n, 40) as ctx:
        ctx.set_option("mdot_variant", variant)
        out1 = ctx.op_mdot(V, w)
        out2 = ctx.op_mdot(V, w)
    np.testing.assert_array_equal(out1, out2)                      # fixed summation order
    ref = np.concatenate([V @ w, [w @ w]])
    scale = np.concatenate([np.abs(V) @ np.abs(w), [w @ w]]) + 1e-300
    assert np.max(np.abs(out1 - ref) / scale) <= 1e-13


@pytest.mark.parametrize("variant", [1, 2, 4, 8])
def test_mdot_variants_and_grid_sizes(variant):
    n, m = 70_001, 13
    rng = np.random.default_rng(7)
    V = rng.standard_normal((m, n)); w = rng.standard_normal(n)
    ref = np.concatenate([V @ w, [w @ w]])
    for ctas in (1, 4, 16):
        with KrylovContext(n, 16) as ctx:
            ctx.set_option("mdot_variant", variant)
            ctx.set_option("ctas_per_sm", ctas)
            out = ctx.op_mdot(V, w)
        np.testing.assert_allclose(out, ref, rtol=1e-12)


@pytest.mark.parametrize("n", [2, 1023, 1025, 4096, 300_007])
@pytest.mark.parametrize("m", [0, 1, 4, 7, 21])
@pytest.mark.parametrize("with_base", [True, False])
def test_lincomb(n, m, with_base):
    rng = np.random.default_rng(n * 31 + m)
    V = rng.standard_normal((m, n))
    base = rng.standard_normal(n) if with_base else None
    coef = rng.standard_normal(m)
    with KrylovContext(n, 24) as ctx:
        out, ss = ctx.op_lincomb(V, base, coef, sign=-1.0)
    ref = (base if with_base else 0.0) - coef @ V if m else (base if with_base else np.zeros(n))
    ref = np.asarray(ref, dtype=float) * np.ones(n)
    scale = (np.abs(base) if with_base else 0.0) + np.abs(coef) @ np.abs(V) + 1e-300 if m else np.ones(n)
    assert np.max(np.abs(out - ref) / scale) <= 1e-14
    assert abs(ss - ref @ ref) <= 1e-12 * max(ref @ ref, 1e-300)


@pytest.mark.parametrize("variant", [2, 4, 8])
def test_lincomb_variants(variant):
    n, m = 50_003, 11
    rng = np.random.default_rng(9)
    V = rng.standard_normal((m, n)); base = rng.standard_normal(n); coef = rng.standard_normal(m)
    with KrylovContext(n, 16) as ctx:
        ctx.set_option("lincomb_variant", variant)
        out, ss = ctx.op_lincomb(V, base, coef, sign=1.0)
    np.testing.assert_allclose(out, base + coef @ V, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("n", [2, 255, 256, 257, 511, 513, 4096, 300_007])
@pytest.mark.parametrize("m", [1, 2, 7, 8, 9, 21, 33, 50, 53])
def test_orth_mid_fused_kernel(n, m):
    """TMA-staged fused (w -= V c ; dots = V w): same w bits as the lincomb kernel, dots to rounding."""
    rng = np.random.default_rng(17 * n + m)
    V = rng.standard_normal((m, n))
    w = rng.standard_normal(n)
    coef = rng.standard_normal(m)
    with KrylovContext(n, 64) as ctx:
        w1, d1 = ctx.op_orth_mid(V, w, coef)
        w2, d2 = ctx.op_orth_mid(V, w, coef)
        wl, _ = ctx.op_lincomb(V, w, coef, sign=-1.0)
    np.testing.assert_array_equal(w1, w2)
    np.testing.assert_array_equal(d1, d2)                           # deterministic
    np.testing.assert_array_equal(w1, wl)                           # same fma chain as the unfused kernel
    ref_w = w - coef @ V
    scale = np.abs(w) + np.abs(coef) @ np.abs(V) + 1e-300
    assert np.max(np.abs(w1 - ref_w) / scale) <= 1e-14
    ref_d = V @ w1
    dscale = np.abs(V) @ np.abs(w1) + 1e-300
    assert np.max(np.abs(d1 - ref_d) / dscale) <= 1e-13


def test_orth_mid_refuses_what_does_not_fit():
    """m = 64 needs 2 x 65 x 2 KB of staging: more than one SM has; the entry point says so and the
    Arnoldi step takes the two-kernel path on its own (test_gpu_solvers covers k > 55)."""
    n, m = 1000, 64
    rng = np.random.default_rng(0)
    with KrylovContext(n, 64) as ctx:
        with pytest.raises(nat.SpisError) as err:
            ctx.op_orth_mid(rng.standard_normal((m, n)), rng.standard_normal(n), rng.standard_normal(m))
        assert err.value.code == nat.E_UNSUPPORTED


@pytest.mark.parametrize("stages", [2, 3, 8])
def test_orth_mid_ring_depths(stages):
    n, m = 1_000_003, 5
    rng = np.random.default_rng(3)
    V = rng.standard_normal((m, n)); w = rng.standard_normal(n); coef = rng.standard_normal(m)
    with KrylovContext(n, 8) as ctx:
        ctx.set_option("orth_mid_max_stages", stages)
        w1, d1 = ctx.op_orth_mid(V, w, coef)
    np.testing.assert_allclose(w1, w - coef @ V, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(d1, V @ w1, rtol=1e-11, atol=1e-9)


def test_preconditioner_kernels():
    d, _ = lkdv.linforms(space="CG", M=1000, mlength=800.0)
    A = d["A"]
    n = A.shape[0]
    rng = np.random.default_rng(4)
    q = rng.standard_normal(n)
    with KrylovContext(n, 2) as ctx:
        ctx.upload_vec(nat.VEC_PRE_DIAG, 1.0 / A.diagonal())
        ctx.set_precond(nat.PRE_JACOBI)
        np.testing.assert_allclose(ctx.op_precond(q), q / A.diagonal(), rtol=1e-15)
        for layout, bs in (("field", 3), ("contiguous", 3), ("contiguous", 8), ("contiguous", 1)):
            P = BlockJacobiPreconditioner(A, bs, layout)
            ctx.upload_blocks(P.inv_blocks, P.stride_block, P.stride_field)
            ctx.set_precond(nat.PRE_BLOCK)
            np.testing.assert_allclose(ctx.op_precond(q), P @ q, rtol=1e-12, atol=1e-12)
        Pm = sps.tril(A).tocsr()
        ctx.upload_matrix(nat.SLOT_PRE, Pm)
        ctx.set_precond(nat.PRE_CSR)
        np.testing.assert_allclose(ctx.op_precond(q), Pm @ q, rtol=1e-12, atol=1e-12)


def test_any_nonzero_host_and_pinned_paths():
    """Zero detection of constraint data (`0*A`, lkdv/LinearSolver.py:30): host threads by default (also
    for page-locked buffers: PCIe is the scarce resource of an end-to-end solve), optionally the copy engine plus a
    kernel for page-locked ones (pinned_scan_dma); -0.0 is zero, NaN is not."""
    import torch
    n = 300_001
    with KrylovContext(1000, 2) as ctx:
        for pinned in (False, True, "dma"):
            ctx.set_option("pinned_scan_dma", 1 if pinned == "dma" else 0)   # page-locked: host threads (default) or copy engine + kernel
            def buf(a):
                return torch.from_numpy(a).pin_memory().numpy() if pinned else a
            z = buf(np.zeros(n))
            assert ctx.any_nonzero(z) is False
            z[:] = -0.0
            assert ctx.any_nonzero(z) is False
            for pos in (0, 1, n // 2, n - 2, n - 1):
                z[:] = 0.0
                z[pos] = 1e-300
                assert ctx.any_nonzero(z) is True
            z[:] = 0.0
            z[n - 1] = np.nan
            assert ctx.any_nonzero(z) is True
            z[:] = 0.0
            assert ctx.any_nonzero(z[1:]) is False                   # 8-byte aligned only
            z[5] = 2.0
            assert ctx.any_nonzero(z[1:]) is True and ctx.any_nonzero(z[6:]) is False


def test_errors_are_reported_not_swallowed():
    with KrylovContext(100, 4) as ctx:
        with pytest.raises(nat.SpisError):
            ctx.arnoldi_step(0)                                     # before solve_begin
        with pytest.raises(ValueError):
            ctx.upload_vec(nat.VEC_B, np.zeros(99))
        with pytest.raises(ValueError):
            ctx.upload_matrix(nat.SLOT_A, sps.identity(99, format="csr"))
        with pytest.raises(nat.SpisError):
            ctx.op_spmv(3, np.zeros(100))                           # slot never uploaded
        with pytest.raises(nat.SpisError):
            ctx.set_option("no_such_option", 1)
    with pytest.raises(nat.SpisError):
        KrylovContext(100, 4, device=99)


def test_bench_kernel_entry_point():
    d, _ = lkdv.linforms(space="CG", M=33_350, mlength=0.8 * 33_350)
    n = d["A"].shape[0]
    with KrylovContext(n, 8) as ctx:
        ctx.upload_matrix(nat.SLOT_A, d["A"])
        for cls in (nat.PROF_SPMV, nat.PROF_MDOT, nat.PROF_LINCOMB, nat.PROF_SCALE, nat.PROF_ORTHMID):
            ms, by = ctx.bench_kernel(cls, 6, reps=3)
            assert ms > 0 and by > 0


@pytest.mark.parametrize("fmt", [nat.FMT_SELL, nat.FMT_SELLD, nat.FMT_PATTERN, nat.FMT_CSR])
def test_constraint_stage_multi_vector_spmv_is_bit_identical(fmt):
    """The constraint stage forms M z_j for groups of four / two Krylov columns from ONE pass over M
    (spmv_pattern_multi_kernel / spmv_sell_multi_kernel): same per-row summation order as the single-vector
    kernels, so term1 and term2 (solvers.py:35-36) are the same bits with the grouping switched off; against
    numpy to rounding.  m = 7 exercises groups of 4, 2 and 1."""
    from structurepreservingiterativesolvers_b200.problems import swe
    rng = np.random.default_rng(29)
    if fmt in (nat.FMT_SELLD, nat.FMT_CSR):
        d = swe.linforms(M=40, mlength=32.0)[0]
    else:
        d = lkdv.linforms(space="CG", M=33_350, mlength=0.8 * 33_350)[0]
    A, L = d["A"], d["L"].tocsr()
    n = A.shape[0]
    b, x0, v = rng.standard_normal(n), 0.01 * rng.standard_normal(n), rng.standard_normal(n)
    m = 7
    out = []
    for multi in (1, 0):
        with KrylovContext(n, 8) as ctx:
            ctx.set_option("spmv_format", fmt)
            ctx.set_option("spmv_multi", multi)
            ctx.upload_matrix(nat.SLOT_A, A)
            ctx.upload_matrix(nat.SLOT_CON0, L)
            ctx.upload_vec(nat.VEC_B, b)
            ctx.upload_vec(nat.VEC_X0, x0)
            ctx.constraint_define(0, nat.SLOT_CON0, v, 0.25)
            ctx.solve_begin()
            for j in range(m):
                ctx.arnoldi_step(j)
            t0, t1, t2 = ctx.constraint_terms(0, m)
            Z = ctx.download_Z(0, m)
            out.append((t0, t1, t2, Z))
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][2], out[1][2])
    t0, t1, t2, Z = out[0]
    LZ = (L @ Z.T)
    ref2 = 0.5 * Z @ LZ
    ref1 = Z @ v + x0 @ LZ
    assert np.max(np.abs(t2 - ref2)) <= 1e-13 * np.max(np.abs(ref2))
    assert np.max(np.abs(t1 - ref1)) <= 1e-12 * np.max(np.abs(ref1))
    assert abs(t0 - (0.5 * x0 @ (L @ x0) + v @ x0 + 0.25)) <= 1e-13 * max(1.0, abs(t0))


@pytest.mark.parametrize("m,first,x0_zero", [(21, 0, True), (24, 0, False), (9, 0, True), (40, 3, False), (49, 0, True), (17, 9, True)])
def test_one_pass_constraint_reduction(m, first, x0_zero):
    """gram_kernel (FP64 mma.sync m8n8k4 tiles): term2 = Z^T M Z and x0.MZ for >= 8 new columns of a symmetric M in
    one pass over Z and M Z (solvers.py:33-36), against the four-columns-per-pass path (option gram = 0) and numpy.
    `first` columns are reduced beforehand (the incremental case c0 > 0: only rows i <= column are formed, the rest
    is mirrored); n is not a multiple of the 16-row chunk, m of the 8-row tile."""
    rng = np.random.default_rng(31 + m)
    d = lkdv.linforms(space="CG", M=3_337, mlength=0.8 * 3_337)[0]
    A, L = d["A"], d["L"].tocsr()
    L = (L + L.T).tocsr() * 0.5
    n = A.shape[0]
    assert n % 16 != 0
    b, v = rng.standard_normal(n), rng.standard_normal(n)
    x0 = np.zeros(n) if x0_zero else 0.01 * rng.standard_normal(n)
    out = []
    for gram in (1, 0):
        with KrylovContext(n, 50) as ctx:
            ctx.set_option("gram", gram)
            ctx.upload_matrix(nat.SLOT_A, A)
            ctx.upload_matrix(nat.SLOT_CON0, L)
            ctx.upload_vec(nat.VEC_B, b)
            ctx.upload_vec(nat.VEC_X0, x0)
            ctx.set_option("x0_is_zero", 1 if x0_zero else 0)
            ctx.constraint_define(0, nat.SLOT_CON0, v, 0.25)
            ctx.solve_begin()
            for j in range(m):
                ctx.arnoldi_step(j)
            if first:
                ctx.constraint_terms(0, first)
            ctx.reset_profile()
            t0, t1, t2 = ctx.constraint_terms(0, m)
            launches = ctx.profile()["mdot"]["launches"]
            Z = ctx.download_Z(0, m)
            out.append((t0, t1, t2, Z, launches))
    t0, t1, t2, Z, launches = out[0]
    assert out[1][4] > launches                            # one gram launch per 24 columns instead of one mdotm per 4
    LZ = (L @ Z.T)
    ref2 = 0.5 * Z @ LZ
    ref1 = Z @ v + x0 @ LZ
    assert np.max(np.abs(t2 - ref2)) <= 1e-13 * np.max(np.abs(ref2))
    assert np.max(np.abs(t1 - ref1)) <= 1e-12 * np.max(np.abs(ref1))
    np.testing.assert_array_equal(t2, t2.T)
    np.testing.assert_allclose(t2, out[1][2], rtol=0, atol=1e-13 * np.max(np.abs(ref2)))
    np.testing.assert_allclose(t1, out[1][1], rtol=0, atol=1e-12 * np.max(np.abs(ref1)))
    assert t0 == out[1][0]


@pytest.mark.parametrize("m,x0_zero,own_v", [(21, True, False), (24, False, True), (12, True, True)])
def test_linear_invariants_ride_in_the_one_pass_reduction(m, x0_zero, own_v):
    """A linear invariant (M == 0, term1 = v.Z: lkdv's mass, lkdv/LinearSolver.py:28-32) that is asked for before the
    quadratic one with the same columns pending costs no pass over Z of its own: v is one more column of the
    quadratic constraint's Gram pass, whose terms are kept for the following call.  Same numbers as the separate
    passes (option gram = 0) to rounding, and as numpy."""
    rng = np.random.default_rng(131 + m)
    d = lkdv.linforms(space="CG", M=3_337, mlength=0.8 * 3_337)[0]
    A, L = d["A"], d["L"].tocsr()
    L = (L + L.T).tocsr() * 0.5
    n = A.shape[0]
    b, v_lin, v_q = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    x0 = np.zeros(n) if x0_zero else 0.01 * rng.standard_normal(n)
    out = []
    for gram in (1, 0):
        with KrylovContext(n, 30) as ctx:
            ctx.set_option("gram", gram)
            ctx.upload_matrix(nat.SLOT_A, A)
            ctx.upload_matrix(nat.SLOT_CON0 + 1, L)
            ctx.upload_vec(nat.VEC_B, b)
            ctx.upload_vec(nat.VEC_X0, x0)
            ctx.set_option("x0_is_zero", 1 if x0_zero else 0)
            ctx.constraint_define(0, -1, v_lin, 0.5)
            ctx.constraint_define(1, nat.SLOT_CON0 + 1, v_q if own_v else None, 0.25)
            ctx.solve_begin()
            for j in range(m):
                ctx.arnoldi_step(j)
            ctx.constraint_terms(1, 2)                     # settles the symmetry test and term0 of the quadratic one ...
            ctx.constraint_terms(0, 2)                     # ... and both are two columns in: c0 = 2 for the batch below
            ctx.reset_profile()
            lin = ctx.constraint_terms(0, m)
            after_lin = ctx.profile()["mdot"]["launches"]
            quad = ctx.constraint_terms(1, m)
            after_quad = ctx.profile()["mdot"]["launches"]
            Z = ctx.download_Z(0, m)
            out.append((lin, quad, Z, after_lin, after_quad))
    (lin, quad, Z, after_lin, after_quad), ref = out
    assert after_lin == 1 and after_quad == 1              # one Gram launch served both calls
    assert ref[4] > 1
    LZ = L @ Z.T
    np.testing.assert_allclose(lin[1], Z @ v_lin, rtol=0, atol=1e-12 * np.abs(Z @ v_lin).max())
    assert lin[0] == ref[0][0] and np.all(lin[2] == 0.0)
    ref1 = (Z @ v_q if own_v else 0.0) + x0 @ LZ
    np.testing.assert_allclose(quad[1], ref1, rtol=0, atol=1e-12 * max(np.abs(ref1).max(), 1e-300))
    np.testing.assert_allclose(quad[2], 0.5 * Z @ LZ, rtol=0, atol=1e-13 * np.abs(Z @ LZ).max())
    np.testing.assert_allclose(lin[1], ref[0][1], rtol=0, atol=1e-12 * np.abs(ref[0][1]).max())
    np.testing.assert_allclose(quad[1], ref[1][1], rtol=0, atol=1e-12 * max(np.abs(ref[1][1]).max(), 1e-300))
    np.testing.assert_allclose(quad[2], ref[1][2], rtol=0, atol=1e-13 * np.abs(ref[1][2]).max())


@pytest.mark.parametrize("x0_zero,nquad", [(True, 2), (False, 2), (True, 3), (False, 1)])
def test_constraint_terms_of_all_constraints_in_one_pass(x0_zero, nquad):
    """spis_constraint_terms_batch: when every constraint is one column behind (a constrained iteration,
    solvers.py:242-247; every iteration of cgmres_p) the quadratic ones share one pass over Z -- their M z_col as the
    right-hand sides of one mdotm launch -- one reduction and one synchronisation.  Same terms as the per-constraint
    calls (option batch_terms = 0) to rounding, and as numpy; fewer launches."""
    rng = np.random.default_rng(41 + nquad)
    d = lkdv.linforms(space="CG", M=3_337, mlength=0.8 * 3_337)[0]
    A = d["A"]
    n = A.shape[0]
    Ms = []
    for q in range(nquad):
        L = (d["L"] if q % 2 == 0 else d["A"]).tocsr() * (1.0 + 0.25 * q)
        Ms.append(((L + L.T) * 0.5).tocsr())
    b = rng.standard_normal(n)
    x0 = np.zeros(n) if x0_zero else 0.01 * rng.standard_normal(n)
    vs = [rng.standard_normal(n) if q != 1 else None for q in range(nquad)] + [rng.standard_normal(n)]
    steps = 11
    out = []
    for batch in (1, 0):
        with KrylovContext(n, 16) as ctx:
            ctx.set_option("batch_terms", batch)
            ctx.upload_matrix(nat.SLOT_A, A)
            for q, Mq in enumerate(Ms):
                ctx.upload_matrix(nat.SLOT_CON0 + q, Mq)
            ctx.upload_vec(nat.VEC_B, b)
            ctx.upload_vec(nat.VEC_X0, x0)
            ctx.set_option("x0_is_zero", 1 if x0_zero else 0)
            for q in range(nquad):
                ctx.constraint_define(q, nat.SLOT_CON0 + q, vs[q], 0.25 + q)
            ctx.constraint_define(nquad, -1, vs[nquad], -0.5)                      # a linear one
            ctx.solve_begin()
            cs = list(range(nquad + 1))
            terms, launches = [], 0
            for j in range(steps):
                ctx.arnoldi_step(j)
                ctx.reset_profile()
                terms.append(ctx.constraint_terms_batch(cs, j + 1))
                launches += ctx.profile()["mdot"]["launches"]
            out.append((terms, ctx.download_Z(0, steps), launches))
    (tb, Z, lb), (ts, _, ls) = out
    if nquad > 1:
        assert lb < ls
    for j in range(steps):
        Zj = Z[: j + 1]
        for c in range(nquad + 1):
            t0, t1, t2 = tb[j][c]
            s0, s1, s2 = ts[j][c]
            assert t0 == s0
            if c < nquad:
                MZ = Ms[c] @ Zj.T
                ref2 = 0.5 * Zj @ MZ
                ref1 = (Zj @ vs[c] if vs[c] is not None else 0.0) + x0 @ MZ
            else:
                ref2 = np.zeros((j + 1, j + 1)); ref1 = Zj @ vs[c]
            sc2 = max(np.abs(ref2).max(), 1e-300); sc1 = max(np.abs(ref1).max(), 1e-300)
            np.testing.assert_allclose(t2, ref2, rtol=0, atol=1e-13 * sc2)
            np.testing.assert_allclose(t1, ref1, rtol=0, atol=1e-12 * sc1)
            np.testing.assert_allclose(t2, s2, rtol=0, atol=1e-13 * sc2)
            np.testing.assert_allclose(t1, s1, rtol=0, atol=1e-12 * sc1)


@pytest.mark.parametrize("fmt", [nat.FMT_SELL, nat.FMT_SELLD])
def test_sell_sigma_row_sorting(fmt):
    """SELL-C-sigma (option sell_sigma, sigma = 256): rows sorted by length inside windows of 256 so that a slice is
    not padded to its longest row.  Applied only when it removes >= 5 % of the stored entries (swe: 16- and 9-entry
    rows interleaved; ragged rows).  A row keeps its entries in order; which of the two accumulators a trailing
    entry lands in depends on the slice width, so results agree with the unsorted storage to rounding."""
    from structurepreservingiterativesolvers_b200.problems import swe
    rng = np.random.default_rng(31)
    mats = [(swe.linforms(M=40, mlength=32.0)[0]["A"], True)]
    R = ragged_matrix(4097, seed=5)
    R.data[:] = rng.integers(-5, 6, size=R.nnz).astype(np.float64)
    mats.append((R, True))
    mats.append((lkdv.linforms(space="CG", M=33_350, mlength=0.8 * 33_350)[0]["A"], False))   # uniform rows: left alone
    for A, expect_gain in mats:
        n = A.shape[0]
        b, x0 = rng.standard_normal(n), rng.standard_normal(n)
        out = []
        for sigma in (1, 0):
            with KrylovContext(n, 2) as ctx:
                ctx.set_option("spmv_format", fmt)
                ctx.set_option("sell_sigma", sigma)
                ctx.upload_matrix(nat.SLOT_A, A)
                assert ctx.info(f"fmt:{nat.SLOT_A}") == fmt
                out.append(_modes_through_abi(ctx, A, b, x0) + (ctx.info(f"nnz_padded:{nat.SLOT_A}"),))
        scale = np.abs(A) @ np.abs(x0) + 1e-300
        assert np.max(np.abs(out[0][0] - out[1][0]) / scale) <= 1e-15
        assert np.max(np.abs(out[0][0] - A @ x0) / scale) <= 1e-14
        ref = np.linalg.norm(b - A @ x0)
        for y, beta, res, padded in out:
            assert abs(beta - ref) <= 1e-13 * ref and abs(res - ref) <= 1e-13 * ref
        if expect_gain:
            assert out[0][3] <= 0.95 * out[1][3] and out[0][3] >= A.nnz
        else:
            assert out[0][3] == out[1][3]


# ---- round 2: row patterns with x staged through shared memory (spmv_fw_kernel) -----------------------------------------
def _fw_matrices():
    from structurepreservingiterativesolvers_b200.problems import heat, lkdvRK
    out = [("lkdv P1, 3 fields", lkdv.linforms(space="CG", M=33_350, mlength=0.8 * 33_350)[0], 3),
           ("lkdv P1, N odd", lkdv.linforms(space="CG", M=4_097, mlength=0.8 * 4_097)[0], 3),
           ("lkdv DG1", lkdv.linforms(space="DG", M=1_500)[0], 3),
           ("lkdvRK P1, 2 stages x 3 fields", lkdvRK.linforms(M=5_000, space="CG", mlength=4000.0)[0], 6),
           ("heat P1 on a 30 x 30 grid", heat.linforms(M=29)[0], None)]          # boundary rows: too many stencils for the table, gather kernels
    return out


@pytest.mark.parametrize("case", range(5))
def test_field_window_spmv_is_bit_identical_to_the_gather_kernels(case):
    """spmv_fw_kernel (x windows of every field block staged in shared memory by TMA bulk copies, all field blocks of a
    node range by one CTA) against the L1-gather row-pattern kernels: same per-row summation order, so mode 0, the dual
    product and the grouped constraint products give the same bits; the fused norms agree to rounding."""
    name, dic, fields = _fw_matrices()[case]
    A = sps.csr_matrix(dic["A"])
    n = A.shape[0]
    rng = np.random.default_rng(5 + case)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for fw in (1, 0):
        with KrylovContext(n, 6) as ctx:
            ctx.set_option("spmv_format", nat.FMT_PATTERN)
            ctx.set_option("spmv_fw", 7 if fw else 0)             # every product kind through the staged kernel
            ctx.upload_matrix(nat.SLOT_A, A)
            if fields is not None:
                assert ctx.info(f"fw_fields:{nat.SLOT_A}") == (fields if fw else 0), name
            y, beta, res = _modes_through_abi(ctx, A, b, x0)
            # the dual product of an Arnoldi step and a few more steps (grouped products come with the constraint stage below)
            col0 = ctx.arnoldi_step(0)
            ctx.form_iterate(np.array([0.37]))
            ctx.arnoldi_begin_residual(1)
            res1 = ctx.iterate_residual_wait()
            ctx.arnoldi_finish(1)
            col1 = ctx.arnoldi_wait(1)
            w = ctx.download(nat.VEC_W)
            out.append((y, beta, res, col0, col1, res1, w))
    ref = np.linalg.norm(b - A @ x0)
    np.testing.assert_allclose(out[0][0], A @ x0, rtol=0, atol=1e-13 * np.abs(A).dot(np.abs(x0)).max())
    np.testing.assert_array_equal(out[0][0], out[1][0])
    # (beta comes out of a fused norm whose per-CTA grouping differs: everything after q0 = r0 / beta agrees to rounding)
    np.testing.assert_allclose(out[0][3], out[1][3], rtol=1e-13)
    np.testing.assert_allclose(out[0][4], out[1][4], rtol=1e-12, atol=1e-13 * np.abs(out[1][4]).max())
    for o in out:
        assert abs(o[1] - ref) <= 1e-13 * ref and abs(o[2] - ref) <= 1e-13 * ref
    assert abs(out[0][5] - out[1][5]) <= 1e-13 * max(out[1][5], 1e-300)


def test_field_window_spmv_in_the_constraint_stage_and_with_ghost_columns():
    """Grouped products M z_j (4 and 2 Krylov columns per pass) through the field-window kernel, and a matrix with ghost
    columns as a row-sharded strip has them (entries that fit no window are gathered one by one): terms and products
    equal the gather kernels' bit for bit."""
    from structurepreservingiterativesolvers_b200 import solvers, wrappers
    d, _ = lkdv.linforms(space="CG", M=20_000, mlength=16_000.0)
    A, b = d["A"], d["b"]
    n = b.size
    x0 = 0.01 * np.cos(np.arange(n))
    cl = wrappers.lkdv.conlist(d, x0)
    terms = []
    for fw in (1, 0):
        sess = solvers.DeviceSession(A, b, x0, 12, conlist=cl, spmv_format="pattern", async_setup=False)
        sess.ctx.set_option("spmv_fw", 7 if fw else 0)
        sess.begin()
        for j in range(8):
            sess.arnoldi_launch(j); sess.arnoldi_wait(j)
        terms.append([sess.ctx.constraint_terms(c, 7) for c in range(3)] + [sess.ctx.constraint_terms(c, 8) for c in range(3)])
        sess.close()
    for a, c in zip(*terms):
        assert abs(a[0] - c[0]) <= 1e-13 * max(abs(c[0]), 1.0)
        np.testing.assert_allclose(a[1], c[1], rtol=1e-11, atol=1e-13 * np.abs(c[1]).max())
        np.testing.assert_allclose(a[2], c[2], rtol=1e-11, atol=1e-13 * np.abs(c[2]).max())
    # ghost columns: the first rank's strip of a 2-way sharded system (owned columns first, ghosts behind them)
    from structurepreservingiterativesolvers_b200.partition import FieldBlockPartition, localize, take_rows
    part = FieldBlockPartition(3, 20_000, 2)
    (A_loc,), plan = localize([take_rows(A, part, 0)], part, 0)
    n_loc = A_loc.shape[0]
    xx = np.random.default_rng(0).standard_normal(A_loc.shape[1])
    ys = []
    for fw in (1, 0):
        with KrylovContext(n_loc, 2, n_halo=plan.n_halo) as ctx:
            ctx.set_option("spmv_format", nat.FMT_PATTERN)
            ctx.set_option("spmv_fw", 7 if fw else 0)
            ctx.upload_matrix(nat.SLOT_A, A_loc)
            assert ctx.info(f"fw_fields:{nat.SLOT_A}") == (3 if fw else 0)
            ys.append(ctx.op_spmv(nat.SLOT_A, xx))
    np.testing.assert_array_equal(ys[0], ys[1])
    np.testing.assert_allclose(ys[0], A_loc @ xx, rtol=0, atol=1e-12 * np.abs(A_loc @ xx).max())


# ---- round 2: SELL / SELLD with per-tile x windows in shared memory and 16-bit columns (spmv_sellw_kernel) ---------------
@pytest.mark.parametrize("fmt", [nat.FMT_SELL, nat.FMT_SELLD])
@pytest.mark.parametrize("case", ["swe", "lkdv", "lkdvRK", "ragged"])
def test_sellw_spmv_is_bit_identical_to_the_plain_sell_kernels(case, fmt):
    """Window analysis at upload + staged kernels against the L1-gather SELL / SELLD kernels: mode 0, the dual product
    and the grouped products give the same bits (same per-row order); the fused norms agree to rounding.  A matrix
    whose tiles do not decompose into a few windows (random columns) silently keeps the plain kernels."""
    from structurepreservingiterativesolvers_b200.problems import lkdvRK, swe
    rng = np.random.default_rng(11)
    if case == "swe":
        A = swe.linforms(M=40, mlength=32.0)[0]["A"]
    elif case == "lkdv":
        A = lkdv.linforms(space="CG", M=33_350, mlength=0.8 * 33_350)[0]["A"]
    elif case == "lkdvRK":
        A = lkdvRK.linforms(M=5_000, space="CG", mlength=4000.0)[0]["A"]
    else:
        A = ragged_matrix(20_011, seed=5)
        A.data[:] = rng.integers(-5, 6, size=A.nnz).astype(np.float64)
    A = sps.csr_matrix(A)
    n = A.shape[0]
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for sw in (7, 0):
        with KrylovContext(n, 6) as ctx:
            ctx.set_option("spmv_format", fmt)
            ctx.set_option("spmv_sellw", sw)
            ctx.upload_matrix(nat.SLOT_A, A)
            assert ctx.info(f"fmt:{nat.SLOT_A}") == fmt
            cap = ctx.info(f"sellw_cap:{nat.SLOT_A}")
            assert (cap > 0) == (sw != 0 and case != "ragged"), (case, cap)
            y, beta, res = _modes_through_abi(ctx, A, b, x0)
            col0 = ctx.arnoldi_step(0)
            ctx.form_iterate(np.array([0.37]))
            ctx.arnoldi_begin_residual(1)
            res1 = ctx.iterate_residual_wait()
            ctx.arnoldi_finish(1)
            col1 = ctx.arnoldi_wait(1)
            out.append((y, beta, res, col0, col1, res1))
    ref = np.linalg.norm(b - A @ x0)
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_allclose(out[0][0], A @ x0, rtol=0, atol=1e-13 * abs(A).dot(np.abs(x0)).max())
    np.testing.assert_allclose(out[0][3], out[1][3], rtol=1e-13)
    np.testing.assert_allclose(out[0][4], out[1][4], rtol=1e-12, atol=1e-13 * np.abs(out[1][4]).max())
    for o in out:
        assert abs(o[1] - ref) <= 1e-13 * ref and abs(o[2] - ref) <= 1e-13 * ref
    assert abs(out[0][5] - out[1][5]) <= 1e-13 * max(out[1][5], 1e-300)


def test_sellw_in_the_constraint_stage():
    """Grouped products M z_j of the constraint stage (4 / 2 columns per pass over M) through the staged SELL kernel."""
    from structurepreservingiterativesolvers_b200 import solvers, wrappers
    from structurepreservingiterativesolvers_b200.problems import swe
    d, _ = swe.linforms(M=30, mlength=24.0)
    A, b = d["A"], d["b"]
    x0 = 0.01 * np.cos(np.arange(b.size))
    cl = wrappers.swe.conlist(d, x0)
    terms = []
    for sw in (7, 0):
        sess = solvers.DeviceSession(A, b, x0, 12, conlist=cl, async_setup=False)
        sess.ctx.set_option("spmv_sellw", sw)
        sess.begin()
        for j in range(8):
            sess.arnoldi_launch(j); sess.arnoldi_wait(j)
        terms.append([sess.ctx.constraint_terms(c, 7) for c in range(2)] + [sess.ctx.constraint_terms(c, 8) for c in range(2)])
        sess.close()
    for a, c in zip(*terms):
        assert abs(a[0] - c[0]) <= 1e-13 * max(abs(c[0]), 1.0)
        np.testing.assert_allclose(a[1], c[1], rtol=1e-11, atol=1e-13 * max(np.abs(c[1]).max(), 1e-300))
        np.testing.assert_allclose(a[2], c[2], rtol=1e-11, atol=1e-13 * max(np.abs(c[2]).max(), 1e-300))
