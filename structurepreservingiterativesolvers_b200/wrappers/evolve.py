"""Time loop of the linear KdV experiment with the system RESIDENT on the GPU (lkdv/Evolve.py:18-65).

The reference re-assembles the forms every step (`lkdv.linforms(..., zinit=sol[-1])`, Evolve.py:41) and hands
them to `cgmresWrapper` / `gmresWrapper`, which rebuild everything from scratch.  On a fixed mesh with a fixed
time step only the right-hand side and the invariant values of the step's initial state change: A, the three
constraint matrices and vectors and the Krylov workspace are uploaded ONCE here (`solvers.DeviceSession`) and every
step sends just b (and the three scalars) -- `DeviceSession.update`.  `resident=False` is the reference's call
pattern (a fresh upload per step) and gives the same numbers.
"""
from __future__ import annotations

import numpy as np

from .. import solvers
from ..problems import lkdv as lkdv_problem
from . import lkdv as lkdv_wrapper


def evolve(N=100, M=50, degree=1, k=50, tol=1e-6, contol=10, solver="cgmres", *, space="DG", mlength=None,
           steps=None, resident=True, ctx_factory=None, **ext):
    """Returns the reference's dict: 'sol' (list of state vectors), 'time', 'dm', 'dmo', 'de' (absolute
    deviations of mass / momentum / energy from their initial values, Evolve.py:58-62), plus 'steps' (Krylov
    iterations per time step).  `solver`: 'cgmres' or 'gmres' (the two wrappers Evolve.py:72-86 compares);
    `steps` limits the number of time steps (default N - 1, as the reference)."""
    if solver not in ("cgmres", "gmres"):
        raise ValueError("solver must be 'cgmres' or 'gmres'")
    forms, prob = lkdv_problem.linforms(N=N, M=M, degree=degree, space=space, mlength=mlength)
    sol = [forms["z0"].copy()]
    time = [0.0]
    inv = lkdv_problem.compute_invariants(forms, forms["z0"])
    mass, momentum, energy = [forms["m0"]], [forms["mo0"]], [forms["e0"]]
    its = []
    nsteps = N - 1 if steps is None else int(steps)
    x0 = np.zeros_like(forms["b"])
    sess = None
    skw = {} if ctx_factory is None else {"ctx_factory": ctx_factory}      # (tests: a CPU stand-in for the context)
    try:
        for i in range(1, nsteps + 1):
            forms, _ = lkdv_problem.linforms(N=N, M=M, degree=degree, space=space, mlength=mlength, zinit=sol[-1])   # Evolve.py:41
            if solver == "cgmres":
                cl = lkdv_wrapper.conlist(forms, x0)
                if resident:
                    if sess is None:
                        sess = solvers.DeviceSession(forms["A"], forms["b"], x0, k, conlist=cl, **skw)
                    else:
                        sess.update(b=forms["b"], constants=[c.c for c in cl])
                    z, info = solvers.cgmres(forms["A"], forms["b"], x0, k, tol=tol, contol=contol, conlist=cl,
                                             session=sess, **ext)
                elif ctx_factory is not None:
                    with_sess = solvers.DeviceSession(forms["A"], forms["b"], x0, k, conlist=cl, **skw)
                    z, info = solvers.cgmres(forms["A"], forms["b"], x0, k, tol=tol, contol=contol, conlist=cl,
                                             session=with_sess, **ext)
                    z = np.array(z, copy=True); with_sess.close()
                else:
                    z, info = lkdv_wrapper.cgmresWrapper(forms, x0=x0, k=k, tol=tol, contol=contol, **ext)     # Evolve.py:45
            else:
                if resident:
                    if sess is None:
                        sess = solvers.DeviceSession(forms["A"], forms["b"], x0, k, **skw)
                    else:
                        sess.update(b=forms["b"])
                    z, info = solvers.gmres(forms["A"], forms["b"], x0, k, tol=tol, session=sess, **ext)
                else:
                    z, info = lkdv_wrapper.gmresWrapper(forms, x0=x0, k=k, tol=tol, **ext)
            z = np.array(z, dtype=np.float64, copy=True)          # the session's result buffers are recycled
            inv = lkdv_problem.compute_invariants(forms, z)                                                   # Evolve.py:49-53
            mass.append(inv["mass"]); momentum.append(inv["momentum"]); energy.append(inv["energy"])
            sol.append(z)
            time.append(forms["T"] / N * i)
            its.append(info.get("steps", k))
    finally:
        if sess is not None:
            sess.close()
    mass, momentum, energy = np.asarray(mass), np.asarray(momentum), np.asarray(energy)
    return {"sol": sol, "time": time, "dm": np.abs(mass - mass[0]), "dmo": np.abs(momentum - momentum[0]),
            "de": np.abs(energy - energy[0]), "steps": its}
