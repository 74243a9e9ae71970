#!/usr/bin/env python
"""Kernel tuning sweep on resident data: GB/s of every kernel class for each variant / grid size.

    python tools/tune.py [--n 10000050] [--out gpurun_out/tune.json]

Uses spis_bench_kernel (CUDA events around `reps` back-to-back launches, operands >> L2).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from structurepreservingiterativesolvers_b200 import _native as nat  # noqa: E402
from structurepreservingiterativesolvers_b200.device import KrylovContext  # noqa: E402
from structurepreservingiterativesolvers_b200.problems import lkdv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--kmax", type=int, default=50)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tune.json"))
    ap.add_argument("--peak", type=float, default=6547.5)
    args = ap.parse_args()
    M = lkdv.benchmark_size(args.n)
    t0 = time.time()
    dic, _ = lkdv.linforms(space="CG", M=M, mlength=0.8 * M)
    n = dic["b"].size
    print(f"assembled n={n} in {time.time()-t0:.1f}s", flush=True)
    rows = []
    with KrylovContext(n, args.kmax) as ctx:
        for fmt in (nat.FMT_SELL, nat.FMT_CSR):
            ctx.set_option("spmv_format", fmt)
            t0 = time.time()
            ctx.upload_matrix(nat.SLOT_A, dic["A"])
            up = time.time() - t0
            for ctas in (2, 4, 8, 16):
                ctx.set_option("spmv_ctas_per_sm", ctas)
                ms, by = ctx.bench_kernel(nat.PROF_SPMV, 0, reps=20)
                rows.append(dict(kernel="spmv", fmt=fmt, ctas=ctas, ms=ms, gbs=by / ms * 1e-6, upload_s=up))
                print(rows[-1], flush=True)
        ctx.set_option("spmv_format", nat.FMT_SELL)
        ctx.upload_matrix(nat.SLOT_A, dic["A"])
        for m in (1, 4, 10, 25, 50):
            for variant in (2, 4, 8):
                for ctas in (1, 2, 4, 8):
                    ctx.set_option("ctas_per_sm", ctas)
                    ctx.set_option("mdot_variant", variant)
                    ctx.set_option("lincomb_variant", variant)
                    reps = 20 if m <= 10 else 8
                    ms, by = ctx.bench_kernel(nat.PROF_MDOT, m, reps=reps)
                    rows.append(dict(kernel="mdot", m=m, variant=variant, ctas=ctas, ms=ms, gbs=by / ms * 1e-6))
                    ms, by = ctx.bench_kernel(nat.PROF_LINCOMB, m, reps=reps)
                    rows.append(dict(kernel="lincomb", m=m, variant=variant, ctas=ctas, ms=ms, gbs=by / ms * 1e-6))
            best_d = max((r for r in rows if r["kernel"] == "mdot" and r["m"] == m), key=lambda r: r["gbs"])
            best_l = max((r for r in rows if r["kernel"] == "lincomb" and r["m"] == m), key=lambda r: r["gbs"])
            print(f"m={m}: best mdot {best_d['gbs']:.0f} GB/s ({best_d['gbs']/args.peak:.2f}) v{best_d['variant']} c{best_d['ctas']}"
                  f" | best lincomb {best_l['gbs']:.0f} GB/s ({best_l['gbs']/args.peak:.2f}) v{best_l['variant']} c{best_l['ctas']}", flush=True)
        ms, by = ctx.bench_kernel(nat.PROF_SCALE, 1, reps=20)
        rows.append(dict(kernel="scale", ms=ms, gbs=by / ms * 1e-6))
        print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(rows, fh, indent=1)
    # compact table
    for kern in ("mdot", "lincomb"):
        print(f"== {kern}: GB/s by (m) x (variant, ctas)")
        for m in (1, 4, 10, 25, 50):
            cells = [f"v{r['variant']}c{r['ctas']}:{r['gbs']:.0f}" for r in rows if r["kernel"] == kern and r.get("m") == m]
            print(f"m={m:2d} " + " ".join(cells))


if __name__ == "__main__":
    main()
