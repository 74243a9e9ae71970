"""The numpy oracle must reproduce the outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

import cases
import helpers


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_matches_reference_output(name, golden):
    spec, dic, prob, x0, pre = cases.instantiate(name)
    # the committed fixture was generated from exactly these inputs
    np.testing.assert_allclose(helpers.fingerprint(dic), golden[f"{name}/fingerprint"], rtol=1e-13)
    x, info = helpers.run_oracle(name)
    # same library calls in the same order as the reference -> round-off agreement.  The lkdvRK
    # cases go through a re-written z1calc (vectorised sum over stages), hence 1e-10 there.
    tol = 1e-10 if spec["exp"] == "lkdvRK" else 1e-13
    assert helpers.rel_diff(x, golden[f"{name}/x_last"]) <= tol
    ref_res = golden[f"{name}/res"]
    assert len(info["res"]) == len(ref_res)
    np.testing.assert_allclose(info["res"], ref_res, rtol=1e-6, atol=1e-12)
    assert info.get("steps", -1) == int(golden[f"{name}/steps"])
    X = golden[f"{name}/X"]
    assert len(info["x"]) == len(X)
    assert helpers.rel_diff(info["x"][0], X[0]) <= 1e-14          # quirk Q1: x[0] is r0


def test_oracle_timing_keys():
    x, info = helpers.run_oracle("lkdv_dg1_tol6_timing")
    assert sorted(info["timings"]) == ["constrained_steps", "constraint_building", "iter_time_constrained",
                                       "iter_time_unconstrained", "runtime"]
    assert info["timings"]["constrained_steps"] == 1


def test_structured_lkdvrk_constraints_reduce_to_the_reference_callbacks(golden):
    """wrappers.lkdvRK.conlist_structured (class-form quadratics in the stage vector) against the dict-form
    callbacks of lkdvRK/LinearSolver.py:29-76: same values and gradients on a random basis, and the oracle
    driven with them reproduces the reference's solve."""
    import warnings
    from oracle import cgmres_oracle as orc
    from structurepreservingiterativesolvers_b200 import wrappers
    name = "lkdvrk_tol6"
    spec, dic, prob, x0, pre = cases.instantiate(name)
    cb = wrappers.lkdvRK.conlist(dic, x0, prob)
    st = wrappers.lkdvRK.conlist_structured(dic, x0, prob)
    rng = np.random.default_rng(0)
    Z, y = rng.standard_normal((x0.size, 5)), rng.standard_normal(5)
    X = x0 + Z @ y
    for a, b in zip(cb, st):
        assert abs(a["func"](y, x0, Z) - (0.5 * X @ (b.M @ X) + b.v @ X + b.c)) <= 1e-13
        np.testing.assert_allclose((b.M @ X + b.v) @ Z, a["jac"](y, x0, Z), rtol=0, atol=1e-13)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x, info = orc.cgmres(dic["A"], dic["b"], x0, spec["k"], tol=spec["tol"], contol=spec["contol"], conlist=st, pre=pre)
    assert info["steps"] == int(golden[f"{name}/steps"])
    assert helpers.rel_diff(x, golden[f"{name}/x_last"]) <= 1e-9
