"""Mirrors of the reference's per-experiment ``LinearSolver.py`` wrappers (SURVEY layer L2).

Each module exposes ``cgmresWrapper`` / ``gmresWrapper`` (and ``exact`` where the reference has
one) with the reference's argument names, builds the same constraint list and dispatches to the
B200 solvers in ``..solvers``.  ``conlist(...)`` returns just the constraint list so that tests can
hand identical constraints to the oracle.
"""
from . import lkdv, swe, heat, lkdvRK  # noqa: F401
from .evolve import evolve, evolve_lkdvRK, evolve_swe  # noqa: F401
