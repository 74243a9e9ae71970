"""CPU tests of the host side of solvers.py (control flow, quirks, history, small solvers).

The device is replaced by tests/fake_ctx.FakeKrylovContext (numpy; CGS2 arithmetic like the CUDA
library).  These tests therefore pin the HOST logic against the reference's golden outputs; the
`-m gpu` tests repeat the same comparisons with the real kernels.
"""
import warnings

import numpy as np
import pytest
import scipy.sparse as sps

import cases
import helpers
from fake_ctx import FakeKrylovContext
from tolerances import tolerance
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200 import smallsolve, solvers, wrappers
from structurepreservingiterativesolvers_b200.preconditioners import (BlockJacobiPreconditioner,
                                                                      JacobiPreconditioner)
from structurepreservingiterativesolvers_b200.problems import lkdv


@pytest.mark.parametrize("engine", ["slsqp", "kkt"])
@pytest.mark.parametrize("name", list(cases.CASES))
def test_host_flow_matches_reference(name, engine, golden):
    x, info, dic, prob = helpers.run_product(name, ctx_factory=FakeKrylovContext, small_solver=engine)
    assert info.get("steps", -1) == int(golden[f"{name}/steps"])
    assert len(info["res"]) == len(golden[f"{name}/res"])
    assert helpers.rel_diff(x, golden[f"{name}/x_last"]) <= tolerance(name)
    X = golden[f"{name}/X"]
    assert len(info["x"]) == len(X)
    helpers.check_r0(info, dic, cases.instantiate(name)[3], golden, name)   # Q1: x[0] is r0
    # the returned vector is the last entry of the history (solvers.py:323)
    np.testing.assert_array_equal(info["x"][-1], x)
    # 'kkt' may legitimately pick another KKT point while the prototype solver's early steps are
    # infeasible / strongly nonlinear (3 constraints, 4 unknowns); only 'slsqp' is held to the
    # reference on every intermediate iterate there
    proto = info["name"] == "geosolve"
    if engine == "slsqp" or not proto:
        helpers.check_histories(name, info, dic, golden)


def test_return_dict_keys():
    x, info, *_ = helpers.run_product("lkdv_cg_tol6", ctx_factory=FakeKrylovContext)
    assert set(info) == {"name", "x", "res", "steps", "timings"} and info["name"] == "cgmres"
    assert info["timings"] is None
    x, info, *_ = helpers.run_product("lkdv_dg1_gmres", ctx_factory=FakeKrylovContext)
    assert set(info) == {"name", "x", "res", "steps"} and info["name"] == "gmres"
    x, info, *_ = helpers.run_product("lkdv_dg1_proto", ctx_factory=FakeKrylovContext)
    assert set(info) == {"name", "x", "res"} and info["name"] == "geosolve"
    assert len(info["res"]) == 20 and len(info["x"]) == 21        # k residuals, r0 + k iterates


def test_timing_dict():
    x, info, *_ = helpers.run_product("lkdv_dg1_tol6_timing", ctx_factory=FakeKrylovContext)
    t = info["timings"]
    assert sorted(t) == ["constrained_steps", "constraint_building", "iter_time_constrained",
                         "iter_time_unconstrained", "runtime"]
    assert t["constrained_steps"] == 1 and t["runtime"] > 0


def test_lazy_history_semantics():
    x, info, dic, _ = helpers.run_product("lkdv_cg_tol6", ctx_factory=FakeKrylovContext)
    hist = info["x"]
    assert len(hist) == info["steps"] + 1
    assert isinstance(hist[1:], list) and len(hist[1:]) == info["steps"]
    for j in range(1, len(hist)):                                  # visualise.py reads x[j], res[j-1]
        r = np.linalg.norm(dic["A"] @ hist[j] - dic["b"])
        assert abs(r - info["res"][j - 1]) <= 1e-10 * np.linalg.norm(dic["b"])
    with pytest.raises(IndexError):
        hist[len(hist)]
    info["z0"] = 1                                                  # callers mutate the dict (heat/SingleSolve.py:51)


def test_eager_history_is_a_list():
    x, info, *_ = helpers.run_product("lkdv_cg_tol6", ctx_factory=FakeKrylovContext, history="eager")
    assert isinstance(info["x"], list) and all(isinstance(a, np.ndarray) for a in info["x"])


def test_stale_history_raises():
    spec, dic, prob, x0, pre = cases.instantiate("lkdv_cg_tol6")
    cl = wrappers.lkdv.conlist(dic, x0)
    sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 50, conlist=cl, ctx_factory=FakeKrylovContext)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, info1 = solvers.cgmres(dic["A"], dic["b"], x0, 50, tol=1e-6, conlist=cl, session=sess)
        _, info2 = solvers.cgmres(dic["A"], dic["b"], x0, 50, tol=1e-6, conlist=cl, session=sess)
    info1["x"][-1]                       # cached last iterate stays available
    with pytest.raises(RuntimeError):
        info1["x"][1]
    info2["x"][1]


def test_lookahead_overlaps_but_keeps_results():
    spec, dic, prob, x0, pre = cases.instantiate("lkdv_cg_tol6")
    cl = wrappers.lkdv.conlist(dic, x0)
    outs = []
    for la in (True, False):
        sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 50, conlist=cl, ctx_factory=FakeKrylovContext)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, info = solvers.cgmres(dic["A"], dic["b"], x0, 50, tol=1e-6, conlist=cl, session=sess, lookahead=la)
        outs.append((x, sess.ctx.log))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    log = outs[0][1]
    # with lookahead the first half of Arnoldi step j+1 is queued before iterate j exists, and the iterate is
    # then formed by the last projection sweep of that step (no separate iterate pass)
    assert log.index(("begin_step", 1)) < log.index(("fused_iterate", 1))
    # (the last iterate has no following Arnoldi step to ride on: it is the only separate iterate pass)
    assert [e for e in log if e[0] == "iterate"] == [("iterate", info["steps"])]
    rides = [e for e in log if e[0] == "residual_rides"]
    assert len([e for e in log if e[0] == "fused_iterate"]) == len([e for e in log if e[0] == "residual_launch"]) + len(rides) == info["steps"] - 1
    # while the loop is far from the tolerance the residual of iterate j is measured by the SpMV of Arnoldi step
    # j+2 (one pass over A for both products): that step is begun right behind the sweep that formed the iterate
    assert len(rides) >= info["steps"] // 2
    for e in rides:
        at = log.index(e)
        assert log[at - 1][0] == "fused_iterate" and log[at + 1] == ("begin_step", e[1])
    # ... and never for the step that ends the loop: nothing is begun that is not needed
    assert max(e[1] for e in log if e[0] == "begin_step") <= info["steps"]
    log_nl = outs[1][1]
    assert log_nl.index(("launch", 1)) > log_nl.index(("iterate", 1))
    assert not [e for e in log_nl if e[0] == "fused_iterate"]
    # every step that was begun is waited for exactly once (also a speculative one after convergence)
    for lg in (log, log_nl):
        assert sorted(e[1] for e in lg if e[0] == "begin_step") == sorted(e[1] for e in lg if e[0] == "wait")


def test_constraints_only_built_in_constrained_phase():
    spec, dic, prob, x0, pre = cases.instantiate("lkdv_cg_tol6")
    cl = wrappers.lkdv.conlist(dic, x0)
    sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 50, conlist=cl, ctx_factory=FakeKrylovContext)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        solvers.cgmres(dic["A"], dic["b"], x0, 50, tol=1e-6, conlist=cl, session=sess)
    terms = [e for e in sess.ctx.log if e[0] == "terms"]
    first_m = min(e[2] for e in terms)
    assert first_m > 1                                              # unconstrained early iterations
    # the mass constraint has M = 0*A (explicit zeros): no matrix slot was uploaded for it
    assert sess.ctx.cons[0]["slot"] < 0 and sess.ctx.cons[1]["slot"] >= 0


def test_constraints_are_staged_by_a_helper_thread():
    """DeviceSession hands the constraint scan + upload to a helper thread (auxiliary stream on the GPU)
    and joins it before the reduced constraints are first needed; async_setup=False keeps one thread."""
    import threading
    spec, dic, prob, x0, pre = cases.instantiate("lkdv_cg_tol6")
    cl = wrappers.lkdv.conlist(dic, x0)
    results = []
    for async_setup in (True, False):
        sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 50, conlist=cl, ctx_factory=FakeKrylovContext,
                                     async_setup=async_setup)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, info = solvers.cgmres(dic["A"], dic["b"], x0, 50, tol=1e-6, conlist=cl, session=sess)
        # one helper thread per constraint (the host scan of one overlaps the upload of another)
        assert getattr(sess.ctx, "aux_switches", 0) == (len(cl) if async_setup else 0)
        assert not getattr(sess.ctx, "aux_on_caller_thread", False)
        assert sess.n_constraints == 3
        results.append(x)
        sess.close()
    np.testing.assert_array_equal(results[0], results[1])


def test_helper_thread_errors_surface_on_the_callers_thread():
    spec, dic, prob, x0, pre = cases.instantiate("lkdv_cg_tol6")

    class Broken:
        M, c = dic["M"], 0.0
        v = np.zeros(7)                       # wrong length: the reference fails when it builds the container
    sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 10, conlist=[Broken()], ctx_factory=FakeKrylovContext)
    with pytest.raises(ValueError):
        sess.containers(1)
    sess.close()


def test_invalid_constraint_type():
    spec, dic, prob, x0, pre = cases.instantiate("lkdv_cg_tol6")
    bad = [("not", "a", "constraint")]
    sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 5, conlist=bad, ctx_factory=FakeKrylovContext)
    with pytest.raises(NotImplementedError):                        # solvers.py:30 via cgmres_p (no try)
        solvers.cgmres_p(dic["A"], dic["b"], x0, 5, conlist=bad, session=sess)
    sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 5, conlist=bad, ctx_factory=FakeKrylovContext)
    with pytest.warns(RuntimeWarning, match="Constrained solve failed"):   # swallowed by cgmres' bare except
        x, info = solvers.cgmres(dic["A"], dic["b"], x0, 5, tol=1e-12, conlist=bad, session=sess)
    assert info["steps"] == 5


def test_breakdown_returns_r0():
    # exact initial guess up to round-off is not needed: A = I, b = e1, x0 = 0 -> Krylov space is 1-D
    n = 16
    A = sps.identity(n, format="csr")
    b = np.zeros(n); b[0] = 2.0
    sess = solvers.DeviceSession(A, b, np.zeros(n), 4, ctx_factory=FakeKrylovContext)
    with pytest.warns(UserWarning, match="broke down"):
        x, info = solvers.gmres(A, b, np.zeros(n), 4, session=sess)
    assert info["steps"] == 1 and len(info["x"]) == 1 and info["res"] == []
    np.testing.assert_array_equal(x, b)                             # Q3: breakdown at j=0 returns r0


def test_preconditioner_dispatch():
    spec, dic, prob, x0, pre = cases.instantiate("heat_tol7")
    A = dic["A"]
    mk = lambda p: solvers.DeviceSession(A, dic["b"], x0, 3, pre=p, ctx_factory=FakeKrylovContext).ctx.pre_kind
    assert mk(None) == nat.PRE_NONE
    assert mk(sps.diags(1.0 / A.diagonal())) == nat.PRE_JACOBI
    assert mk(JacobiPreconditioner(A)) == nat.PRE_JACOBI
    d3, _ = lkdv.linforms(space="CG", M=12)
    s3 = solvers.DeviceSession(d3["A"], d3["b"], np.zeros(36), 3, pre=BlockJacobiPreconditioner(d3["A"], 3, "field"),
                               ctx_factory=FakeKrylovContext)
    assert s3.ctx.pre_kind == nat.PRE_BLOCK
    assert mk(sps.linalg.spilu(A.tocsc())) == nat.PRE_HOST
    assert mk(sps.tril(A).tocsr()) == nat.PRE_CSR
    assert mk(sps.linalg.aslinearoperator(A)) == nat.PRE_HOST

    class Nope:
        pass
    sess = solvers.DeviceSession(A, dic["b"], x0, 3, pre=Nope(), ctx_factory=FakeKrylovContext)
    with pytest.raises(ValueError, match="Preconditioner not supported"):
        solvers.gmres(A, dic["b"], x0, 3, session=sess)


def test_block_jacobi_matches_explicit_matrix():
    d, _ = lkdv.linforms(space="CG", M=40)
    A = d["A"]
    rng = np.random.default_rng(0)
    v = rng.standard_normal(A.shape[0])
    for layout, bs in (("field", 3), ("contiguous", 3), ("contiguous", 4)):
        P = BlockJacobiPreconditioner(A, bs, layout)
        np.testing.assert_allclose(P @ v, P.tocsr() @ v, rtol=1e-12, atol=1e-12)
        # inverse of the block diagonal: P * blockdiag(A) = I on the block pattern
        idx = (np.arange(P.nblk)[:, None] * P.stride_block + np.arange(bs)[None, :] * P.stride_field)
        blk0 = A[idx[0]][:, idx[0]].toarray()
        np.testing.assert_allclose(P.inv_blocks[0] @ blk0, np.eye(bs), atol=1e-10)
    J = JacobiPreconditioner(A)
    np.testing.assert_allclose(J @ v, v / A.diagonal())


def test_kkt_agrees_with_slsqp_on_feasible_problems():
    rng = np.random.default_rng(3)
    for m in (3, 6, 12):
        H = np.triu(rng.standard_normal((m + 1, m)), -1) + 3 * np.eye(m + 1, m)
        beta = 2.5
        S = rng.standard_normal((m, m)); T2 = 0.01 * (S + S.T)
        t1 = rng.standard_normal(m)
        y_ls = np.linalg.lstsq(H, np.r_[beta, np.zeros(m)], rcond=None)[0]
        t0 = -(t1 @ y_ls + y_ls @ T2 @ y_ls) + 1e-3            # slightly violated at the LS solution
        con = smallsolve.ReducedConstraint(t0, t1, T2)
        r_k = smallsolve.kkt(H, beta, np.zeros(m), [con])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r_s = smallsolve.slsqp(H, beta, y_ls, [con], ftol=1e-24)
        assert r_k.success and abs(con.fun(r_k.x)) < 1e-13
        assert np.linalg.norm(r_k.x - r_s.x) <= 1e-7 * np.linalg.norm(r_s.x)
        f = lambda y: np.sum((np.r_[beta, np.zeros(m)] - H @ y) ** 2)
        assert f(r_k.x) <= f(r_s.x) * (1 + 1e-9) + 1e-15


def test_configure_rejects_unknown_keys():
    with pytest.raises(KeyError):
        solvers.configure(nonsense=1)
    old = solvers.configure()["small_solver"]
    solvers.configure(small_solver="kkt")
    try:
        x, info, *_ = helpers.run_product("lkdv_cg_tol6", ctx_factory=FakeKrylovContext)
        assert info["steps"] == 10
    finally:
        solvers.configure(small_solver=old)


def test_system_export_roundtrip(tmp_path):
    from structurepreservingiterativesolvers_b200 import io
    d, _ = lkdv.linforms(space="DG", M=10)
    io.save_system(tmp_path / "sys.npz", d)
    e = io.load_system(tmp_path / "sys.npz")
    assert set(e) == set(d)
    for key in d:
        if sps.issparse(d[key]):
            assert (abs(d[key] - e[key])).nnz == 0 and e[key].nnz == d[key].nnz
        else:
            np.testing.assert_array_equal(np.asarray(d[key]), np.asarray(e[key]))


def test_evolve_resident_session_equals_fresh_uploads():
    """wrappers.evolve (lkdv/Evolve.py:18-65): with the system resident only b and the three invariant values are
    sent per time step (DeviceSession.update); the results are those of a fresh session per step."""
    outs = []
    for resident in (True, False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            outs.append(wrappers.evolve(N=100, M=20, k=30, tol=1e-8, steps=4, resident=resident,
                                        ctx_factory=FakeKrylovContext, small_solver="kkt"))
    a, b = outs
    assert len(a["sol"]) == 5 and a["steps"] == b["steps"] and a["time"] == b["time"]
    for za, zb in zip(a["sol"], b["sol"]):
        np.testing.assert_array_equal(za, zb)
    # CGMRES keeps the three invariants of every step's initial state: drift over four steps at round-off level
    assert max(a["dm"].max(), a["dmo"].max(), a["de"].max()) < 1e-11
    assert a["dm"][0] == 0 and np.all(np.diff(a["time"]) > 0)


def test_givens_least_squares_matches_lstsq():
    """The 'kkt' engine takes the unconstrained minimiser from the Givens QR the Arnoldi driver keeps anyway (one
    back substitution per iteration instead of an SVD-based lstsq): same y to rounding at every step."""
    from structurepreservingiterativesolvers_b200 import smallsolve
    spec, dic, prob, x0, pre = cases.instantiate("lkdv_cg_tol6")
    sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 30, ctx_factory=FakeKrylovContext)
    beta = sess.begin()
    arn = solvers._Arnoldi(sess, 30, False, beta)
    for j in range(12):
        arn.column(j)
        y = arn.ls_solution(j + 1)
        ref = smallsolve.lstsq(arn.H[: j + 2, : j + 1], beta).x
        assert y is not None and np.max(np.abs(y - ref)) <= 1e-12 * np.max(np.abs(ref))
        assert abs(arn.ls_residual() - arn.predicted_residual(y, beta)) <= 1e-10 * beta
    assert arn.ls_solution(5) is None                      # only for the column count just stored


def test_session_update_and_evolve_gmres():
    """DeviceSession.update: new right-hand side / guess / constraint scalars on a resident system; wrappers.evolve
    in its GMRES flavour (the comparison lkdv/Evolve.py:72-86 makes)."""
    spec, dic, prob, x0, pre = cases.instantiate("lkdv_cg_tol6")
    cl = wrappers.lkdv.conlist(dic, x0)
    sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 40, conlist=cl, ctx_factory=FakeKrylovContext)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        xa, _ = solvers.cgmres(dic["A"], dic["b"], x0, 40, tol=1e-8, conlist=cl, session=sess, small_solver="kkt")
        xa = np.array(xa)
        b2 = 2.0 * dic["b"]
        sess.update(b=b2, x0=0.1 * xa, constants=[2.0 * c.c if i == 0 else 4.0 * c.c for i, c in enumerate(cl)])
        xb, _ = solvers.cgmres(dic["A"], b2, 0.1 * xa, 40, tol=1e-8, conlist=cl, session=sess, small_solver="kkt")
    # the doubled system with consistently scaled invariants (mass linear, the others quadratic) has the doubled solution
    assert np.linalg.norm(xb - 2.0 * xa) <= 1e-7 * np.linalg.norm(xa)
    with pytest.raises(ValueError):
        sess.update(constants=[1.0])
    sess.close()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = wrappers.evolve(N=100, M=20, k=30, tol=1e-8, steps=3, solver="gmres", ctx_factory=FakeKrylovContext)
    assert len(out["sol"]) == 4 and len(out["steps"]) == 3 and out["dm"][0] == 0
    with pytest.raises(ValueError):
        wrappers.evolve(solver="direct")


def test_native_kkt_matches_the_python_solver():
    """spis_small_kkt (C++, host code of the library) against smallsolve.kkt's numpy implementation: same Newton
    iteration, same answer to rounding; whatever is not the plain converged case is declined (handled = 0) and stays
    with the Python solver."""
    rng = np.random.default_rng(11)
    for m, nc in ((3, 1), (7, 2), (12, 3), (25, 2), (50, 2)):
        H = np.triu(rng.standard_normal((m + 1, m)), -1) + 3 * np.eye(m + 1, m)
        beta = 1.7
        y_ls = np.linalg.lstsq(H, np.r_[beta, np.zeros(m)], rcond=None)[0]
        cons = []
        for _ in range(nc):
            S = rng.standard_normal((m, m)); T2 = 0.01 * (S + S.T)
            t1 = rng.standard_normal(m)
            t0 = -(t1 @ y_ls + y_ls @ T2 @ y_ls) + 1e-4 * rng.standard_normal()
            cons.append(smallsolve.ReducedConstraint(t0, t1, T2))
        Hbig = np.zeros((m + 5, m + 3)); Hbig[: m + 1, :m] = H                   # a strided view, as the solvers pass it
        view = Hbig[: m + 1, :m]
        try:
            smallsolve.NATIVE_KKT = False
            ref = smallsolve.kkt(view, beta, np.zeros(m), cons)
            smallsolve.NATIVE_KKT = True
            assert smallsolve._kkt_native(view, beta, cons) is not None          # the plain case is taken natively
            out = smallsolve.kkt(view, beta, np.zeros(m), cons)
        finally:
            smallsolve.NATIVE_KKT = True
        assert ref.success and out.success
        assert np.linalg.norm(out.x - ref.x) <= 1e-12 * np.linalg.norm(ref.x)
        assert max(abs(c.fun(out.x)) for c in cons) <= 1e-12
    # declined: opaque callbacks, and an infeasible pair of constraints (Newton cannot converge)
    cb = smallsolve.ReducedConstraint(callbacks={"func": lambda y, x0, Z: y[0] - 1.0, "jac": lambda y, x0, Z: np.eye(1, m)[0]}, x0=None, Z=None)
    assert smallsolve._kkt_native(H, beta, [cb]) is None
    t1 = rng.standard_normal(m)
    bad = [smallsolve.ReducedConstraint(1.0, t1, np.zeros((m, m))), smallsolve.ReducedConstraint(-1.0, t1, np.zeros((m, m)))]
    assert smallsolve._kkt_native(H, beta, bad) is None


# ---- round 2: the device-resident loop (solvers._Pipeline) on the numpy stand-in ---------------------------------
_PIPE_CASES = ["lkdv_cg_tol6", "lkdv_cg_tol8_n1500", "heat_tol7", "heat_tol7_jacobi", "lkdv_cg_kcap", "lkdv_cg_x0",
               "swe_rt_tol7", "swe_rt_h08_n10800", "lkdv_cg_gmres_n1500", "heat_gmres_jacobi", "lkdv_dg1_gmres"]


def _solve_case(name, pipeline, **ext):
    spec, dic, prob, x0, pre = cases.instantiate(name)
    wrap = getattr(wrappers, spec["exp"])
    cl = wrap.conlist(dic, x0) if spec["kind"] == "cgmres" else []
    solvers.configure(pipeline=pipeline)
    try:
        sess = solvers.DeviceSession(dic["A"], dic["b"], x0, spec["k"], conlist=cl, pre=pre, ctx_factory=FakeKrylovContext)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if spec["kind"] == "cgmres":
                x, info = solvers.cgmres(dic["A"], dic["b"], x0, spec["k"], tol=spec["tol"], contol=spec.get("contol", 10), conlist=cl,
                                         pre=pre, session=sess, small_solver="kkt", timing=spec.get("timing"), **ext)
            else:
                x, info = solvers.gmres(dic["A"], dic["b"], x0, spec["k"], tol=spec["tol"], pre=pre, session=sess, **ext)
    finally:
        solvers.configure(pipeline=True)
    return x, info, sess.ctx.log, dic


@pytest.mark.parametrize("name", _PIPE_CASES)
def test_pipelined_loop_follows_the_reference_flow(name, golden):
    """Givens update, y_j, x_j and the phase test computed "on the device" (at queueing time in the stand-in), the host
    only reading records: same step counts as the reference, same final iterate as the host-driven loop, and every
    unconstrained iterate but the stragglers formed by a step that was queued ahead."""
    x, info, log, dic = _solve_case(name, True)
    xh, infoh, logh, _ = _solve_case(name, False)
    assert any(e[0] == "pipe_begin" for e in log) and not any(e[0] == "pipe_begin" for e in logh)
    assert info["steps"] == infoh["steps"] == int(golden[f"{name}/steps"])
    assert helpers.rel_diff(x, golden[f"{name}/x_last"]) <= tolerance(name)
    assert helpers.rel_diff(x, xh) <= max(1e-12, tolerance(name))
    np.testing.assert_array_equal(info["x"][-1], x)
    helpers.check_histories(name, info, dic, golden)
    steps = [e for e in log if e[0] == "step"]
    assert [e[1] for e in steps] == list(range(len(steps)))                    # queued in order, each once
    assert info["steps"] <= len(steps) <= info["steps"] + 1                    # at most one step the loop did not use
    formed = [e[4] for e in steps if e[4] is not None]
    assert formed == sorted(set(formed))                                       # every device-formed iterate once, in order
    host_formed = [e for e in log if e[0] == "iterate"]
    assert len(formed) + len(host_formed) >= info["steps"]
    if info["name"] == "gmres":
        assert len(formed) >= info["steps"] - 1
    # residuals of device-formed iterates ride on the SpMV of a later step, except at the end of the solve
    assert sum(1 for e in steps if e[2]) >= len(formed) - 2


def test_pipelined_loop_hands_over_at_the_phase_switch():
    """Once a measured residual is <= contol*tol the stand-in device stops forming iterates BY ITSELF (phase word), for
    steps that were queued before the host knew; the host forms every constrained iterate and nothing queued ahead
    overwrites it."""
    x, info, log, dic = _solve_case("lkdv_cg_tol6", True)
    steps = [e for e in log if e[0] == "step"]
    asked = [e for e in steps if e[3]]
    refused = [e for e in asked if e[4] is None and e[1] >= 1]
    assert refused, "a step queued ahead with `want_iterate` must have met the phase word"
    first_host = next(i for i, e in enumerate(log) if e[0] == "iterate")
    assert not any(e[0] == "step" and e[4] is not None for e in log[first_host:])
    r = np.linalg.norm(dic["A"] @ x - dic["b"])
    assert abs(r - info["res"][-1]) <= 1e-10 * np.linalg.norm(dic["b"])


def test_last_iterate_is_downloaded_while_it_is_checked():
    """cgmres ends on a constrained iterate whose residual is < tol (solvers.py:296-297).  When the small solve
    predicts that, the iterate is launched through iterate_residual_launch_dl (its download overlaps its formation
    and check) and the array that call filled IS the returned x; with early_download off the plain download gives
    the same bits.  An early download whose iterate was not the last one is joined and dropped."""
    x, info, log, dic = _solve_case("lkdv_cg_tol6", True)
    early = [i for i, e in enumerate(log) if e[0] == "early_download"]
    assert len(early) >= 1
    assert sum(1 for e in log if e[0] == "download_join") >= len(early)      # every early download is joined
    assert not any(e[0] == "iterate_residual_launch" for e in log[early[-1] + 2:])   # nothing was formed after the last one
    solvers.configure(early_download=False)
    try:
        x2, info2, log2, _ = _solve_case("lkdv_cg_tol6", True)
    finally:
        solvers.configure(early_download=True)
    assert not any(e[0] == "early_download" for e in log2)
    np.testing.assert_array_equal(x, x2)
    np.testing.assert_array_equal(info["x"][-1], x)
    assert info["res"] == info2["res"]


def test_pipelined_loop_invalid_record_falls_back_to_the_host():
    """A device record with `valid = 0` (a vanishing or tiny Givens pivot: hess_kernel also sets the phase word) makes the
    host take the general least-squares route (solvers.py:113) and form that and every later iterate itself."""
    class Flaky(FakeKrylovContext):
        def step_enqueue(self, j, want_residual, want_iterate):
            t = super().step_enqueue(j, want_residual, want_iterate)
            if j == 2:
                col, y, info = self._pipe["rec"][j]
                self._pipe["rec"][j] = (col, np.full_like(y, np.nan), dict(info, valid=False, ls=0.0, phase=1))
                self._pipe["phase"] = 1
            return t

    spec, dic, prob, x0, pre = cases.instantiate("lkdv_cg_gmres_n1500")
    out = []
    for factory in (Flaky, FakeKrylovContext):
        sess = solvers.DeviceSession(dic["A"], dic["b"], x0, spec["k"], ctx_factory=factory)
        x, info = solvers.gmres(dic["A"], dic["b"], x0, spec["k"], tol=spec["tol"], session=sess)
        out.append((x, info, sess.ctx.log))
    (xf, inf, logf), (xr, inr, logr) = out
    assert inf["steps"] == inr["steps"]
    assert helpers.rel_diff(xf, xr) <= 1e-10
    np.testing.assert_allclose(inf["res"], inr["res"], rtol=1e-8)
    host_formed = [e[1] for e in logf if e[0] == "iterate"]
    # the phase word is set before step 2's last sweep, which would have formed x_1: x_1 and everything after it are the host's
    assert host_formed[0] == 2 and len(host_formed) == inf["steps"] - 1
    assert not any(e[0] == "step" and e[4] is not None and e[4] >= 2 for e in logf)
    assert not any(e[0] == "step" and e[3] and e[1] > 2 + solvers._Pipeline.DEPTH for e in logf)       # and the host stops asking (steps queued ahead excepted)


def _oracle_loop_swe(N, M, k, tol, steps):
    from oracle import cgmres_oracle as orc
    from structurepreservingiterativesolvers_b200.problems import swe
    forms, _ = swe.linforms(N=N, M=M)
    sol = [forms["z0"].copy()]
    for i in range(1, steps + 1):
        forms, _ = swe.linforms(N=N, M=M, zinit=sol[-1])                       # swe/Evolve.py:39
        x0 = np.zeros_like(forms["b"])
        z, _info = orc.cgmres(forms["A"], forms["b"], x0, k, tol=tol, conlist=wrappers.swe.conlist(forms, x0))
        sol.append(np.array(z))
    return sol


def _oracle_loop_lkdvrk(N, M, k, tol, steps):
    import scipy.sparse.linalg as spsla
    from oracle import cgmres_oracle as orc
    from structurepreservingiterativesolvers_b200.problems import lkdvRK
    forms, prob = lkdvRK.linforms(N=N, M=M, T=1)
    sol = [forms["z0"].copy()]
    z = np.tile(forms["z0"], prob.ns)                                          # lkdvRK/Evolve.py:37
    P = spsla.spilu(forms["A"].tocsc(), drop_tol=1e-4, fill_factor=10)         # lkdvRK/Evolve.py:51-52
    for i in range(1, steps + 1):
        forms, _ = lkdvRK.linforms(N=N, M=M, T=1, zinit=sol[-1])
        z, _info = orc.cgmres(forms["A"], forms["b"], z, k, tol=tol, contol=10, conlist=wrappers.lkdvRK.conlist(forms, z, prob), pre=P)
        z = np.array(z)
        sol.append(lkdvRK.z1calc(prob, z, sol[-1]))
    return sol


def test_evolve_mirrors_of_swe_and_lkdvrk_follow_the_oracle_loops():
    """wrappers.evolve_swe (swe/Evolve.py:18-60) and wrappers.evolve_lkdvRK (lkdvRK/Evolve.py:19-93: previous stage vector
    as the initial guess, one ILU factorisation for all steps, z <- z1calc) with the session resident across the steps,
    against the same loops driven by the oracle."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = wrappers.evolve_swe(N=100, M=6, k=40, tol=1e-7, steps=4, ctx_factory=FakeKrylovContext, small_solver="kkt")
        ref = _oracle_loop_swe(100, 6, 40, 1e-7, 4)
        for i, (a, b) in enumerate(zip(out["sol"], ref)):
            assert helpers.rel_diff(a, b) <= 1e-10 * max(i, 1), i
        assert max(out["dm"].max(), out["de"].max()) <= 1e-12 * (abs(ref[0]).sum())
        ref = _oracle_loop_lkdvrk(10, 20, 30, 1e-6, 4)
        for structured in (True, False):
            out = wrappers.evolve_lkdvRK(N=10, M=20, k=30, tol=1e-6, steps=4, structured=structured,
                                         ctx_factory=FakeKrylovContext, small_solver="kkt")
            for i, (a, b) in enumerate(zip(out["sol"], ref)):
                assert helpers.rel_diff(a, b) <= 1e-10 * max(i, 1), (structured, i)
            assert max(out["dm"].max(), out["dmo"].max(), out["de"].max()) <= 1e-12 * abs(ref[0]).sum()


def test_host_side_row_pattern_detection():
    """spis_host_find_patterns (no device): rows with identical (column - row, value bits) lists share an id, the ids of
    every thread count describe the same partition, a value that differs in its last bit is its own stencil, ghost
    columns move by col_shift, and matrices whose neighbouring rows do not share stencils are given up early."""
    import ctypes as C
    import scipy.sparse as sps
    from structurepreservingiterativesolvers_b200 import _native as nat
    from structurepreservingiterativesolvers_b200.problems import lkdv
    lib = nat.load_library()

    def detect(A, nthreads, n_local=None, shift=0):
        A = A.tocsr(); n = A.shape[0]
        ip = np.ascontiguousarray(A.indptr, dtype=np.int32); ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        da = np.ascontiguousarray(A.data, dtype=np.float64)
        pid = np.zeros(n, dtype=np.uint16); rep = np.zeros(4096, dtype=np.int32)
        npat, ml, ch = C.c_int(0), C.c_int(0), C.c_int64(0)
        rc = lib.spis_host_find_patterns(ip.ctypes.data_as(C.POINTER(C.c_int32)), ci.ctypes.data_as(C.POINTER(C.c_int32)), nat.dptr(da), n,
                                         n if n_local is None else n_local, shift, nthreads, pid.ctypes.data_as(C.POINTER(C.c_uint16)),
                                         rep.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(npat), C.byref(ml), C.byref(ch))
        assert rc == 0
        return pid, rep[: npat.value], npat.value, ml.value, ch.value

    def signature(A, r, n_local=None, shift=0):
        c = A.indices[A.indptr[r]: A.indptr[r + 1]].astype(np.int64)
        if n_local is not None:
            c = np.where(c >= n_local, c + shift, c)
        return tuple(c - r) + tuple(A.data[A.indptr[r]: A.indptr[r + 1]].view(np.int64))

    A = lkdv.linforms(space="CG", M=7_001, mlength=0.8 * 7_001)[0]["A"].tocsr()
    n = A.shape[0]
    sigs = [signature(A, r) for r in range(n)]
    truth = len(set(sigs))
    for nt in (1, 2, 5):
        pid, rep, npat, ml, ch = detect(A, nt)
        assert npat == truth and ml == np.diff(A.indptr).max()
        assert all(sigs[r] == sigs[rep[pid[r]]] for r in range(n))
        assert len({sigs[q] for q in rep}) == npat                         # no stencil twice in the table
        assert ch == int(np.sum(pid[1:] != pid[:-1]))
    B = A.copy(); B.data[B.indptr[1234]] = np.nextafter(B.data[B.indptr[1234]], np.inf)
    assert detect(B, 3)[2] == truth + 1
    # ghost columns: the last 5 columns belong to a neighbour and live 11 entries further out in the device vector
    S = sps.lil_matrix((60, 65))
    for r in range(60):
        S[r, r] = 1.0; S[r, r + 1] = -2.0                                # row 59 reaches ghost column 60
    S[0, 64] = 3.0
    S = S.tocsr()
    pid, rep, npat, ml, ch = detect(S, 2, n_local=60, shift=11)
    sg = [signature(S, r, 60, 11) for r in range(60)]
    assert npat == len(set(sg)) and all(sg[r] == sg[rep[pid[r]]] for r in range(60))
    # every row its own stencil: given up (npat = 0) without reading the matrix to the end
    rng = np.random.default_rng(5)
    V = sps.diags([rng.standard_normal(49_999), rng.standard_normal(50_000)], [-1, 0], format="csr")
    assert detect(V, 4)[2] == 0
