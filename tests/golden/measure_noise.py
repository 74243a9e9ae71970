"""Measure how reproducible the REFERENCE ALGORITHM's own output is at round-off level.

    python tests/golden/measure_noise.py          -> tests/golden/self_noise.json

For every golden case the numpy oracle (bit-identical to the reference on the unperturbed inputs,
tests/test_oracle_golden.py) is re-run on inputs that differ from the golden ones only by round-off:

  * `ulp`  : every stored entry of A and b multiplied by (1 + s*2^-52), s in {-1, 0, +1} (seeded);
  * `perm` : a symmetric permutation of the unknowns (class-form cases without ILU only).

The spread of the final iterate over these runs is the reference's SELF-NOISE: any other correct
implementation (different summation order, CGS2 instead of MGS) differs from the golden output by
about this much.  north_star's 1e-10 bar is met where the reference itself is reproducible to 1e-10;
tests/golden/tolerances.py uses max(1e-10, 3 x self-noise) per case.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.join(HERE, ".."), os.path.join(HERE, "..", "..")]

import cases  # noqa: E402
import helpers  # noqa: E402
from oracle import cgmres_oracle as orc  # noqa: E402
from structurepreservingiterativesolvers_b200 import wrappers  # noqa: E402

N_SAMPLES = 6


class _Q:
    def __init__(self, M, v, c):
        self.M, self.v, self.c = M, v, c


def _solve(spec, A, b, x0, conlist, pre):
    if spec["kind"] == "gmres":
        return orc.fgmres(A, b, x0, spec["k"], tol=spec["tol"], pre=pre)[0]
    proto = (spec["tol"] <= 1e-20) if spec["exp"] in ("lkdv", "lkdvRK") else (spec["tol"] < 1e-20)
    if proto:
        return orc.cgmres_prototype(A, b, x0, spec["k"], conlist=conlist, pre=pre)[0]
    kw = {"contol": spec["contol"]} if ("contol" in spec and spec["exp"] in ("lkdv", "lkdvRK")) else {}
    return orc.cgmres(A, b, x0, spec["k"], tol=spec["tol"], conlist=conlist, pre=pre, **kw)[0]


def main():
    warnings.simplefilter("ignore")
    g = np.load(os.path.join(HERE, "reference_outputs.npz"))
    out = {}
    for name in cases.CASES:
        spec, dic, prob, x0, pre = cases.instantiate(name)
        wrap = getattr(wrappers, spec["exp"])
        ref = g[name + "/x_last"]
        n = dic["b"].size
        samples = []
        for seed in range(1, N_SAMPLES + 1):
            rng = np.random.default_rng(seed)
            A = dic["A"].copy()
            A.data = A.data * (1.0 + rng.integers(-1, 2, A.nnz) * 2.0 ** -52)
            b = dic["b"] * (1.0 + rng.integers(-1, 2, n) * 2.0 ** -52)
            d2 = dict(dic, A=A, b=b)
            cl = []
            if spec["kind"] == "cgmres":
                cl = wrap.conlist(d2, x0, prob) if spec["exp"] == "lkdvRK" else wrap.conlist(d2, x0)
            x = _solve(spec, A, b, x0, cl, cases.make_pre(spec.get("pre"), A))
            samples.append(("ulp", helpers.rel_diff(x, ref)))
            if spec["exp"] != "lkdvRK" and spec.get("pre") not in ("ilu", "ilu_swe"):
                p = rng.permutation(n)
                P = sps.csr_matrix((np.ones(n), (np.arange(n), p)), shape=(n, n))
                Ap = (P @ dic["A"] @ P.T).tocsr()
                prep = None if pre is None else (P @ pre @ P.T)
                clp = []
                if spec["kind"] == "cgmres":
                    clp = [_Q((P @ c.M @ P.T).tocsr(), P @ np.asarray(c.v).reshape(-1), c.c) for c in wrap.conlist(dic, x0)]
                xp = _solve(spec, Ap, P @ dic["b"], P @ x0, clp, prep)
                samples.append(("perm", helpers.rel_diff(P.T @ xp, ref)))
        worst = max(s[1] for s in samples)
        out[name] = {"max": worst, "median": float(np.median([s[1] for s in samples])), "samples": len(samples)}
        print(f"{name:24s} self-noise max {worst:.1e} median {out[name]['median']:.1e} over {len(samples)} runs", flush=True)
    with open(os.path.join(HERE, "self_noise.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
