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
from ..problems import lkdvRK as lkdvrk_problem
from ..problems import swe as swe_problem
from . import lkdv as lkdv_wrapper
from . import lkdvRK as lkdvrk_wrapper
from . import swe as swe_wrapper


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


def evolve_swe(N=100, M=50, degree=1, k=50, tol=1e-6, solver="cgmres", *, mlength=None, steps=None, resident=True,
               ctx_factory=None, **ext):
    """swe/Evolve.py:18-60: the time loop of the linearised rotating shallow-water experiment.  Returns the reference's dict
    ('sol', 'time', 'dm', 'de') plus 'steps' (Krylov iterations per time step).  As in the reference every step starts
    from x0 = 0 (Evolve.py:43); with `resident=True` the operator, the two constraint matrices and the Krylov workspace are
    uploaded once and each step sends only its right-hand side and the two invariant values (`DeviceSession.update`)."""
    if solver not in ("cgmres", "gmres"):
        raise ValueError("solver must be 'cgmres' or 'gmres'")
    forms, prob = swe_problem.linforms(N=N, M=M, degree=degree, mlength=mlength)
    sol, time = [forms["z0"].copy()], [0.0]
    mass, energy, its = [forms["m0"]], [forms["e0"]], []
    nsteps = N - 1 if steps is None else int(steps)
    x0 = np.zeros_like(forms["b"])
    skw = {} if ctx_factory is None else {"ctx_factory": ctx_factory}
    sess = None
    try:
        for i in range(1, nsteps + 1):
            forms, _ = swe_problem.linforms(N=N, M=M, degree=degree, mlength=mlength, zinit=sol[-1])      # Evolve.py:39
            cl = swe_wrapper.conlist(forms, x0) if solver == "cgmres" else []
            if resident or ctx_factory is not None:
                if sess is None or not resident:
                    if sess is not None:
                        sess.close()
                    sess = solvers.DeviceSession(forms["A"], forms["b"], x0, k, conlist=cl, **skw)
                else:
                    sess.update(b=forms["b"], constants=[c.c for c in cl] if cl else None)
                if solver == "cgmres":
                    z, info = solvers.cgmres(forms["A"], forms["b"], x0, k, tol=tol, conlist=cl, session=sess, **ext)
                else:
                    z, info = solvers.gmres(forms["A"], forms["b"], x0, k, tol=tol, session=sess, **ext)
            elif solver == "cgmres":
                z, info = swe_wrapper.cgmresWrapper(forms, x0=x0, k=k, tol=tol, **ext)                       # Evolve.py:43
            else:
                z, info = swe_wrapper.gmresWrapper(forms, x0=x0, k=k, tol=tol, **ext)
            z = np.array(z, dtype=np.float64, copy=True)
            inv = swe_problem.compute_invariants(forms, z)                                                   # Evolve.py:47-49
            mass.append(inv["mass"]); energy.append(inv["energy"])
            sol.append(z)
            time.append(forms["T"] / N * i)
            its.append(info.get("steps", k))
    finally:
        if sess is not None:
            sess.close()
    mass, energy = np.asarray(mass), np.asarray(energy)
    return {"sol": sol, "time": time, "dm": np.abs(mass - mass[0]), "de": np.abs(energy - energy[0]), "steps": its}


def evolve_lkdvRK(N=10, M=50, degree=1, tstages=2, T=1, k=50, tol=1e-6, contol=10, solver="cgmres", *, space="DG",
                  mlength=None, steps=None, resident=True, structured=True, pre="ilu", ctx_factory=None, **ext):
    """lkdvRK/Evolve.py:19-93: Gauss-Legendre time stepping of the linear KdV system.  The unknown of every step is the
    stacked stage vector; the loop feeds the PREVIOUS stage vector back as the initial guess (Evolve.py:37,61: the only
    caller with a non-zero x0), preconditions with one ILU factorisation of the first step's matrix (Evolve.py:51-52) and
    advances z <- z1calc(stages) (Evolve.py:68).  Returns the reference's dict ('sol', 'time', 'dm', 'dmo', 'de'; 'err' is
    left out: it needs the Firedrake error norm) plus 'steps'.
    `resident=True` keeps A, the constraint data and the workspace on the GPU and updates b, x0 and -- for the structured
    class-form constraints -- the linear terms and constants, which depend on the step's initial state; the reference's
    dict-form callbacks (structured=False) are rebuilt per step and need the session rebuilt with them.
    `pre`: 'ilu' (SuperLU through the host bridge, as the reference), 'block' (6x6 node-block Jacobi on the device) or None."""
    import scipy.sparse.linalg as spsla
    from ..preconditioners import BlockJacobiPreconditioner
    if solver not in ("cgmres", "gmres"):
        raise ValueError("solver must be 'cgmres' or 'gmres'")
    forms, prob = lkdvrk_problem.linforms(N=N, M=M, degree=degree, tstages=tstages, T=T, space=space, mlength=mlength)
    z0 = forms["z0"].copy()
    sol, time = [z0], [0.0]
    mass, momentum, energy, its = [forms["m0"]], [forms["mo0"]], [forms["e0"]], []
    z = np.tile(forms["z0"], prob.ns)                                                                     # Evolve.py:37
    if pre == "ilu":
        P = spsla.spilu(forms["A"].tocsc(), drop_tol=1e-4, fill_factor=10)                                 # Evolve.py:51-52
    elif pre == "block":
        P = BlockJacobiPreconditioner(forms["A"], 3 * prob.ns, "field")
    else:
        P = pre
    nsteps = N if steps is None else int(steps)
    skw = {} if ctx_factory is None else {"ctx_factory": ctx_factory}
    sess = None
    try:
        for i in range(1, nsteps + 1):
            forms, _ = lkdvrk_problem.linforms(N=N, M=M, degree=degree, tstages=tstages, T=T, space=space, mlength=mlength,
                                               zinit=sol[-1])                                              # Evolve.py:56-58
            if solver == "cgmres":
                cl = (lkdvrk_wrapper.conlist_structured if structured else lkdvrk_wrapper.conlist)(forms, z, prob)
            else:
                cl = []
            reuse = resident and sess is not None and (structured or solver == "gmres")
            if reuse:
                sess.update(b=forms["b"], x0=z, constants=[c.c for c in cl] if cl else None,
                            vectors=[c.v for c in cl] if cl else None)
            else:
                if sess is not None:
                    sess.close()
                sess = solvers.DeviceSession(forms["A"], forms["b"], z, k, conlist=cl, pre=P, **skw)
            if solver == "cgmres":
                stages, info = solvers.cgmres(forms["A"], forms["b"], z, k, tol=tol, contol=contol, conlist=cl, pre=P,
                                              session=sess, **ext)                                         # Evolve.py:61
            else:
                stages, info = solvers.gmres(forms["A"], forms["b"], z, k, tol=tol, pre=P, session=sess, **ext)
            z = np.array(stages, dtype=np.float64, copy=True)
            znew = lkdvrk_problem.z1calc(prob, z, sol[-1])                                                 # Evolve.py:67-68
            nd = znew.size // 3
            u, w = znew[:nd], znew[2 * nd:]
            Mm = forms["M"][:nd, :nd]
            mass.append(float(forms["omega"][:nd] @ u))                                                    # Evolve.py:73-75
            momentum.append(float(0.5 * u @ (Mm @ u)))
            energy.append(float(0.5 * w @ (Mm @ w) - 0.5 * u @ (Mm @ u)))
            sol.append(znew)
            time.append(forms["T"] / N * i)
            its.append(info.get("steps", k))
            if not resident and sess is not None:
                sess.close()
                sess = None
    finally:
        if sess is not None:
            sess.close()
    mass, momentum, energy = np.asarray(mass), np.asarray(momentum), np.asarray(energy)
    return {"sol": sol, "time": time, "dm": np.abs(mass - mass[0]), "dmo": np.abs(momentum - momentum[0]),
            "de": np.abs(energy - energy[0]), "steps": its}
