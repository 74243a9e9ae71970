"""Per-case parity tolerances (relative 2-norm difference of the final iterate vs the golden
reference output).

north_star's bar is 1e-10.  For a few cases the REFERENCE ITSELF is not reproducible to 1e-10:
its SLSQP calls run with ftol = 1e-24 and stop on `maxiter`, so round-off level changes of H move
the returned y by up to 1e-9 relative.  SELF_NOISE is the oracle-vs-golden difference when the
same system is solved after a symmetric permutation (tests/golden/measure_noise.py, seed 1).
A case's tolerance is max(1e-10, 3 x self-noise); cases not listed use 1e-10.
"""
SELF_NOISE = {
    "lkdv_cg_tol6": 3.0e-11,
    "lkdv_dg1_tol6_timing": 1.4e-10,
    "lkdv_cg_kcap": 2.3e-09,
    "swe_like_tol7": 4.9e-11,
    # dict-form prototype case: step 3 is infeasible (3 constraints, 4 unknowns) and the last steps
    # have an ill-conditioned reduced problem; measured CGS2-vs-MGS sensitivity 4.5e-10
    "lkdvrk_proto": 4.5e-10,
}


def tolerance(name):
    return max(1e-10, 3.0 * SELF_NOISE.get(name, 0.0))
