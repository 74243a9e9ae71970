"""numpy re-assemblies of the reference experiments' linear systems (test-matrix generators).

Firedrake / Irksome / PETSc are not installable here, so each module restates the
weak form of <exp>/<exp>.py with structured-mesh finite elements and returns the same dictionary
keys as the reference's ``linforms``.  They feed the parity tests, bench.py and the wrappers.
"""
from . import lkdv, heat, lkdvRK, swe  # noqa: F401
