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
