#!/usr/bin/env python
"""ncu target: the SpMV kernels of one benchmark matrix on resident data, one launch set per (variant, mode).

    ncu --set full -k regex:spmv_ ... python tools/ncu_spmv.py swe 0,1
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.device import KrylovContext

wl = sys.argv[1] if len(sys.argv) > 1 else "lkdv"
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1").split(",")]
dic, x0, cl, _ = bench.build_system(10_000_000, wl)
A = dic["A"]
with KrylovContext(A.shape[0], 4) as ctx:
    ctx.upload_vec(nat.VEC_B, dic["b"])
    ctx.upload_matrix(nat.SLOT_A, A)
    for var in variants:
        ctx.set_option("spmv_variant", var)
        for mode in (0, 2):
            ms, by = ctx.bench_kernel(nat.PROF_SPMV, mode, reps=1)
            print(wl, "variant", var, "mode", mode, "%.1f us" % (ms * 1e3), flush=True)
