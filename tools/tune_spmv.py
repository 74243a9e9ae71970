#!/usr/bin/env python
"""SpMV formats x modes x grid sizes on the two benchmark matrices (n = 1e7), resident data."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.device import KrylovContext

rows = []
for wl in (sys.argv[1:] or ["lkdv", "swe"]):
    dic, x0, cl, _ = bench.build_system(10_000_000, wl.split("_")[0])
    A = dic["L"] if wl.endswith("_L") else dic["A"]; n = A.shape[0]
    with KrylovContext(n, 4) as ctx:
        ctx.upload_vec(nat.VEC_B, dic["b"])
        for fmt, name in ((nat.FMT_SELL, "sell"), (nat.FMT_SELLD, "selld"), (nat.FMT_PATTERN, "pattern")):
            ctx.set_option("spmv_format", fmt)
            try:
                ctx.upload_matrix(nat.SLOT_A, A)
            except nat.SpisError as exc:
                print(name, 'not applicable:', exc, flush=True)
                continue
            print(name, 'npat', ctx.info('npat:0'), flush=True)
            # (variant, spmv_ctas_per_sm, spmv_pipe_ctas_per_sm); variant 0 = first-generation kernels
            if name == 'pattern':
                sweeps = [(0, 8, 0), (1, 8, 0)]
            elif name == 'sell':
                sweeps = [(0, 8, 0), (1, 8, 4)]
            else:
                sweeps = [(0, 8, 0), (0, 6, 0)]
            print(name, 'ndict', ctx.info('ndict:0'), flush=True)
            for var, ctas, pipe in sweeps:
                ctx.set_option("spmv_variant", var)
                ctx.set_option("spmv_ctas_per_sm", ctas)
                ctx.set_option("spmv_pipe_ctas_per_sm", pipe)
                for mode in (0, 2):
                    ms, by = ctx.bench_kernel(nat.PROF_SPMV, mode, reps=20)
                    rows.append(dict(workload=wl, fmt=name, variant=var, ctas=ctas, pipe=pipe, mode=mode, us=ms * 1e3, gbs=by / ms * 1e-6,
                                     nnz_padded=ctx.info("nnz_padded:0")))
                    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_spmv.json"), "w"), indent=1)
