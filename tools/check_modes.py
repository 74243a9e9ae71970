import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np
import cases
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.device import KrylovContext
for name in ("heat_tol7_jacobi", "lkdv_cg_tol6", "swe_rt_tol7"):
    spec, dic, prob, x0, pre = cases.instantiate(name)
    A, b = dic["A"].tocsr(), dic["b"]
    n = b.size
    rng = np.random.default_rng(0)
    x0 = rng.standard_normal(n)
    for fmt in (nat.FMT_SELL, nat.FMT_AUTO):
        with KrylovContext(n, 4) as ctx:
            ctx.set_option("spmv_format", fmt)
            ctx.upload_matrix(nat.SLOT_A, A)
            ctx.upload_vec(nat.VEC_B, b); ctx.upload_vec(nat.VEC_X0, x0)
            ctx.set_precond(nat.PRE_NONE)
            beta = ctx.solve_begin()
            r0 = ctx.download(nat.VEC_R0)
            ref = b - A @ x0
            y = np.array([1.0])
            ctx.arnoldi_step(0)
            res = ctx.iterate_residual(np.array([0.37]))
            q0 = ref / np.linalg.norm(ref)
            xr = x0 + 0.37 * q0
            print(name, "fmt", ctx.info("fmt:0"), "npat", ctx.info("npat:0"), "beta rel err %.2e" % (abs(beta - np.linalg.norm(ref)) / np.linalg.norm(ref)),
                  "r0 err %.2e" % (np.abs(r0 - ref).max() / np.abs(ref).max()), "res rel err %.2e" % (abs(res - np.linalg.norm(A @ xr - b)) / np.linalg.norm(A @ xr - b)))
