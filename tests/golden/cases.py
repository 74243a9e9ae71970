"""Parity cases shared by make_golden.py (runs the REFERENCE) and the tests (oracle / CUDA path).

Every case is a deterministic linear system from the package's numpy re-assemblies plus the call
the reference's own <exp>/LinearSolver.py wrapper would make.  No RNG is involved except where a
seed is written down.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

from structurepreservingiterativesolvers_b200.problems import heat, lkdv, lkdvRK, swe


def _lkdv(space, M, fixed_h=False, **kw):
    return lkdv.linforms(space=space, M=M, mlength=(0.8 * M if fixed_h else None), **kw)


def _swe_like(dic):
    """A two-constraint (mass + energy) dictionary in the layout swe/LinearSolver.py reads,
    built from the lkdv operators: energy form L_swe := L - M (no RT assembly needed)."""
    out = dict(dic)
    out["L"] = (dic["L"] - dic["M"]).tocsr()
    return out


CASES = {
    # cfg1: lkdv/SingleSolve.py default -> prototypical solver (tol = 1e-50)
    "lkdv_dg1_proto": dict(exp="lkdv", build=lambda: _lkdv("DG", 50), kind="cgmres", k=20, tol=1e-50),
    "lkdv_dg1_gmres": dict(exp="lkdv", build=lambda: _lkdv("DG", 50), kind="gmres", k=20, tol=1e-50),
    # Evolve-style tolerance solves
    "lkdv_cg_tol6": dict(exp="lkdv", build=lambda: _lkdv("CG", 50), kind="cgmres", k=50, tol=1e-6, contol=10),
    "lkdv_cg_tol8_n1500": dict(exp="lkdv", build=lambda: _lkdv("CG", 500, fixed_h=True), kind="cgmres", k=50, tol=1e-8, contol=10),
    "lkdv_cg_gmres_n1500": dict(exp="lkdv", build=lambda: _lkdv("CG", 500, fixed_h=True), kind="gmres", k=30, tol=1e-7),
    "lkdv_dg1_tol6_timing": dict(exp="lkdv", build=lambda: _lkdv("DG", 50), kind="cgmres", k=40, tol=1e-6, contol=10, timing=True),
    # k reached before the tolerance: last iteration is the (only) constrained one
    "lkdv_cg_kcap": dict(exp="lkdv", build=lambda: _lkdv("CG", 200, fixed_h=True), kind="cgmres", k=8, tol=1e-9, contol=10),
    # heat: energy constraint with non-zero linear term; Jacobi through the `pre @ vec` branch
    "heat_tol7": dict(exp="heat", build=lambda: heat.linforms(M=12), kind="cgmres", k=20, tol=1e-7),
    "heat_tol7_jacobi": dict(exp="heat", build=lambda: heat.linforms(M=12), kind="cgmres", k=20, tol=1e-7, pre="jacobi"),
    "heat_proto": dict(exp="heat", build=lambda: heat.linforms(M=8), kind="cgmres", k=10, tol=1e-50),
    "heat_gmres_jacobi": dict(exp="heat", build=lambda: heat.linforms(M=12), kind="gmres", k=20, tol=1e-7, pre="jacobi"),
    # swe wrapper (mass + energy) with an ILU object exposing .solve (swe/TimedSolve.py:23)
    "swe_like_ilu": dict(exp="swe", build=lambda: _swe_wrap(_lkdv("CG", 100, fixed_h=True)), kind="cgmres", k=20, tol=1e-7, pre="ilu"),
    "swe_like_tol7": dict(exp="swe", build=lambda: _swe_wrap(_lkdv("CG", 100, fixed_h=True)), kind="cgmres", k=40, tol=1e-7),
    # cfg3: swe RT_2 x DG_0 re-assembly (problems/swe.py), the calls of swe/TimedSolve.py:17-41 (M = 2**3,
    # tol = 1e-7, k = 20, spilu(drop_tol=1e-2)) and swe/SingleSolve.py:16-37 (tol = 1e-50 -> prototypical)
    "swe_rt_tol7": dict(exp="swe", build=lambda: swe.linforms(M=8), kind="cgmres", k=20, tol=1e-7),
    "swe_rt_tol7_ilu": dict(exp="swe", build=lambda: swe.linforms(M=8), kind="cgmres", k=20, tol=1e-7, pre="ilu_swe", timing=True),
    "swe_rt_gmres_ilu": dict(exp="swe", build=lambda: swe.linforms(M=8), kind="gmres", k=20, tol=1e-7, pre="ilu_swe"),
    "swe_rt_proto": dict(exp="swe", build=lambda: swe.linforms(M=12), kind="cgmres", k=20, tol=1e-50),
    "swe_rt_h08_n10800": dict(exp="swe", build=lambda: swe.linforms(M=30, mlength=24.0), kind="cgmres", k=40, tol=1e-7),
    # lkdvRK: dict-form callbacks, non-zero initial guess, ILU preconditioner (lkdvRK/Evolve.py:51-61)
    "lkdvrk_tol6": dict(exp="lkdvRK", build=lambda: lkdvRK.linforms(M=20), kind="cgmres", k=30, tol=1e-6, contol=10, x0="stage", pre="ilu"),
    "lkdvrk_proto": dict(exp="lkdvRK", build=lambda: lkdvRK.linforms(M=10), kind="cgmres", k=8, tol=1e-50, x0="stage"),
    # non-zero x0 with class-form constraints (term0 / x0.MZ paths)
    "lkdv_cg_x0": dict(exp="lkdv", build=lambda: _lkdv("CG", 50), kind="cgmres", k=50, tol=1e-6, contol=10, x0="perturbed"),
}


def _swe_wrap(pair):
    dic, prob = pair
    return _swe_like(dic), prob


def make_pre(spec, A):
    if spec is None:
        return None
    if spec == "jacobi":
        return sps.diags(1.0 / A.diagonal())            # taken through `pre @ vec` (solvers.py:156-161)
    if spec == "ilu":
        return spsla.spilu(A.tocsc(), drop_tol=1e-4, fill_factor=10)    # lkdvRK/SingleSolve.py:19
    if spec == "ilu_swe":
        return spsla.spilu(A.tocsc(), drop_tol=1e-2, fill_factor=10)    # swe/TimedSolve.py:23-24
    raise KeyError(spec)


def make_x0(spec, dic, prob):
    n = dic["b"].size
    if spec is None:
        return np.zeros(n)
    if spec == "stage":                                  # lkdvRK/Evolve.py:37 starts from tile(z0, ns);
        return 0.01 * np.tile(dic["z0"], prob.ns)        # scaled: the unknowns are stage derivatives
    if spec == "perturbed":
        x = spsla.spsolve(dic["A"].tocsc(), dic["b"])
        return x * (1.0 + 1e-3 * np.cos(np.arange(n)))
    raise KeyError(spec)


def instantiate(name):
    """Returns (spec, dic, prob, x0, pre) for a case name."""
    spec = CASES[name]
    dic, prob = spec["build"]()
    x0 = make_x0(spec.get("x0"), dic, prob)
    pre = make_pre(spec.get("pre"), dic["A"])
    return spec, dic, prob, x0, pre


def wrapper_kwargs(spec, x0, pre, prob):
    """Keyword arguments in the form each reference wrapper accepts."""
    exp, kind = spec["exp"], spec["kind"]
    kw = dict(x0=x0, k=spec["k"], tol=spec["tol"])
    if kind == "cgmres":
        if exp in ("lkdv", "lkdvRK") and "contol" in spec:
            kw["contol"] = spec["contol"]
        if exp != "lkdvRK" and spec.get("timing"):
            kw["timing"] = True
    if exp != "lkdv" and pre is not None:
        kw["pre"] = pre
    if exp == "lkdvRK":
        kw["prob"] = prob
    return kw
