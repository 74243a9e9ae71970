"""Drop-in test: the REFERENCE's own <exp>/LinearSolver.py files, loaded from their place under $SPIS_REFERENCE
(default /root/reference), run against this package standing in for the reference's `solvers` module.

Each wrapper does `sys.path.insert(0, '../'); import solvers` (lkdv/LinearSolver.py:13-14, swe/LinearSolver.py:9-10,
heat/LinearSolver.py:9-10, lkdvRK/LinearSolver.py:11-12) and calls `solvers.cgmres / cgmres_p / gmres` with keywords
(lkdv/LinearSolver.py:50-59,70).  Here `sys.modules['solvers']` is this package's solver module before the wrappers
are executed, so the reference's constraint classes, dict-form callbacks, tolerance dispatch and keyword spelling are
exactly the reference's; the outputs must be the committed outputs of the unmodified reference.

On CPU the device is the numpy stand-in (tests/fake_ctx.py) behind a shim with the reference's signatures; with a GPU
the package's `solvers` module itself is the shim.  The reference tree does not travel to the GPU box: skipped there.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import pytest

import cases
import helpers
from conftest import has_gpu
from tolerances import tolerance
from structurepreservingiterativesolvers_b200 import solvers as product

REF = os.environ.get("SPIS_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "solvers.py")), reason="reference tree not present")


def _cpu_shim():
    """A module with the reference's three entry points whose device is the numpy stand-in."""
    from fake_ctx import FakeKrylovContext
    shim = types.ModuleType("solvers")

    def session(A, b, x0, k, conlist, pre):
        return product.DeviceSession(A, b, x0, k, conlist=conlist, pre=pre, ctx_factory=FakeKrylovContext)

    def gmres(A, b, x0, k, tol=1e-50, pre=None):
        return product.gmres(A, b, x0, k, tol=tol, pre=pre, session=session(A, b, x0, k, (), pre))

    def cgmres(A, b, x0, k, tol=1e-8, contol=10, conlist=[], pre=None, timing=None):
        return product.cgmres(A, b, x0, k, tol=tol, contol=contol, conlist=conlist, pre=pre, timing=timing,
                              session=session(A, b, x0, k, conlist, pre))

    def cgmres_p(A, b, x0, k, conlist=[], pre=None):
        return product.cgmres_p(A, b, x0, k, conlist=conlist, pre=pre, session=session(A, b, x0, k, conlist, pre))

    shim.gmres, shim.cgmres, shim.cgmres_p = gmres, cgmres, cgmres_p
    shim.constraint_container, shim.constraint_checker = product.constraint_container, product.constraint_checker
    return shim


@pytest.fixture(scope="module")
def reference_wrappers(request):
    """{exp: the reference's LinearSolver module}, executed with `solvers` = the shim."""
    shim = product if request.param == "gpu" else _cpu_shim()
    saved = {k: sys.modules.get(k) for k in ("solvers", "firedrake", "matplotlib", "matplotlib.pylab", "irksome", "lkdvRK", "refd")}
    fd = types.ModuleType("firedrake")                      # the wrappers' own imports (lkdvRK.py needs these names)
    fd.warning = lambda msg: warnings.warn(msg)
    fd.Constant = lambda *a, **k: None
    fd.pi = np.pi
    fd.__all__ = ["warning", "Constant", "pi"]
    mpl, pylab = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pylab")
    mpl.pylab = pylab
    sys.modules.update({"firedrake": fd, "matplotlib": mpl, "matplotlib.pylab": pylab, "irksome": types.ModuleType("irksome"),
                        "solvers": shim})

    def load(name, path, extra_path=None):
        if extra_path:
            sys.path.insert(0, extra_path)
        try:
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
        finally:
            if extra_path:
                sys.path.remove(extra_path)
        return mod

    out = {}
    try:
        for exp in ("lkdv", "swe", "heat", "lkdvRK"):
            d = os.path.join(REF, exp)
            if exp == "lkdvRK":                             # its wrapper imports lkdvRK for z1calc / dz1calc (lkdvRK/lkdvRK.py:162-189)
                sys.modules.pop("refd", None)
                load("lkdvRK", os.path.join(d, "lkdvRK.py"), extra_path=d)
            out[exp] = load(f"_dropin_{exp}_LinearSolver", os.path.join(d, "LinearSolver.py"))
            assert out[exp].solvers is shim                 # the wrapper really bound OUR module
        yield out
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for exp in ("lkdv", "swe", "heat", "lkdvRK"):
            sys.modules.pop(f"_dropin_{exp}_LinearSolver", None)


def _run(wrappers, name, golden):
    spec, dic, prob, x0, pre = cases.instantiate(name)
    wrap = wrappers[spec["exp"]]
    fn = wrap.cgmresWrapper if spec["kind"] == "cgmres" else wrap.gmresWrapper
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x, info = fn(dic, **cases.wrapper_kwargs(spec, x0, pre, prob))
    assert info.get("steps", -1) == int(golden[f"{name}/steps"])
    assert len(info["res"]) == len(golden[f"{name}/res"])
    assert len(info["x"]) == len(golden[f"{name}/X"])
    assert helpers.rel_diff(x, golden[f"{name}/x_last"]) <= tolerance(name)
    helpers.check_r0(info, dic, x0, golden, name)
    helpers.check_histories(name, info, dic, golden)


@pytest.mark.parametrize("reference_wrappers", ["cpu"], indirect=True)
@pytest.mark.parametrize("name", list(cases.CASES))
def test_reference_wrappers_on_this_package_cpu(reference_wrappers, name, golden):
    _run(reference_wrappers, name, golden)


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")
@pytest.mark.parametrize("reference_wrappers", ["gpu"], indirect=True)
@pytest.mark.parametrize("name", list(cases.CASES))
def test_reference_wrappers_on_this_package_gpu(reference_wrappers, name, golden):
    _run(reference_wrappers, name, golden)
