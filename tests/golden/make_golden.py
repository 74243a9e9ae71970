"""Generate tests/golden/reference_outputs.npz by running the UNMODIFIED reference.

    python tests/golden/make_golden.py            (needs /root/reference; not available on the GPU box)

The reference's solvers.py imports firedrake (only for `warning`) and matplotlib (unused); neither is
installed, so two stub modules are registered before the import.  Each case of
cases.py is then solved through the reference's own <exp>/LinearSolver.py wrapper, loaded from its
file, on the numpy re-assembled operators.  Outputs: last iterate, residual history, step count,
every iterate, and (cgmres with timing) the timing keys.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SPIS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def import_reference():
    """Returns (solvers module, {exp: LinearSolver module}) of the unmodified reference."""
    fd = types.ModuleType("firedrake")
    fd.warning = lambda msg: warnings.warn(msg)
    fd.Constant = lambda *a, **k: None          # default argument of lkdvRK.linforms (lkdvRK/lkdvRK.py:48)
    fd.pi = np.pi
    fd.__all__ = ["warning", "Constant", "pi"]
    mpl, pylab = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pylab")
    mpl.pylab = pylab
    irk = types.ModuleType("irksome")
    sys.modules.update({"firedrake": fd, "matplotlib": mpl, "matplotlib.pylab": pylab, "irksome": irk})

    def load(name, path, extra_path=None):
        if extra_path:
            sys.path.insert(0, extra_path)
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        if extra_path:
            sys.path.remove(extra_path)
        return mod

    solvers = load("solvers", os.path.join(REF, "solvers.py"))
    wrappers = {}
    for exp in ("lkdv", "swe", "heat", "lkdvRK"):
        d = os.path.join(REF, exp)
        if exp == "lkdvRK":
            # the wrapper imports lkdvRK for z1calc / dz1calc (lkdvRK/lkdvRK.py:162-189)
            for stale in ("refd", "lkdvRK"):
                sys.modules.pop(stale, None)
            load("lkdvRK", os.path.join(d, "lkdvRK.py"), extra_path=d)
        wrappers[exp] = load(f"_ref_{exp}_LinearSolver", os.path.join(d, "LinearSolver.py"))
    return solvers, wrappers


def main():
    import cases
    ref_solvers, ref_wrappers = import_reference()
    out = {}
    manifest = {"numpy": np.__version__, "scipy": __import__("scipy").__version__, "cases": {}}
    for name in cases.CASES:
        spec, dic, prob, x0, pre = cases.instantiate(name)
        wrap = ref_wrappers[spec["exp"]]
        kw = cases.wrapper_kwargs(spec, x0, pre, prob)
        fn = wrap.cgmresWrapper if spec["kind"] == "cgmres" else wrap.gmresWrapper
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, info = fn(dic, **kw)
        out[f"{name}/x_last"] = np.asarray(x)
        out[f"{name}/res"] = np.asarray(info["res"], dtype=float)
        out[f"{name}/X"] = np.asarray(info["x"])
        out[f"{name}/steps"] = np.asarray(info.get("steps", -1))
        # fingerprint of the inputs so that a changed generator is detected, not silently compared
        out[f"{name}/fingerprint"] = np.array([dic["A"].data.sum(), np.abs(dic["A"].data).sum(),
                                               dic["b"].sum(), np.abs(dic["b"]).sum(), float(dic["A"].nnz)])
        meta = {"name": info["name"], "n": int(dic["b"].size), "steps": int(info.get("steps", -1)),
                "final_res": float(info["res"][-1]) if len(info["res"]) else None}
        if info.get("timings"):
            meta["timing_keys"] = sorted(info["timings"].keys())
            meta["constrained_steps"] = int(info["timings"]["constrained_steps"])
        manifest["cases"][name] = meta
        print(name, meta)
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    with open(os.path.join(HERE, "reference_manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
