#!/usr/bin/env python
"""GB/s of the fused CGS2-middle kernel (orth_mid_kernel) vs the two kernels it replaces, per m.

    python tools/tune_orthmid.py [--n 10000050]

`eff GB/s` charges every variant with the UNFUSED algorithmic bytes (2m+3)*8n so that the columns
compare time, not byte accounting.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from structurepreservingiterativesolvers_b200 import _native as nat  # noqa: E402
from structurepreservingiterativesolvers_b200.device import KrylovContext  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_050)
    ap.add_argument("--kmax", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tune_orthmid.json"))
    args = ap.parse_args()
    n = args.n
    rows = []
    with KrylovContext(n, args.kmax) as ctx:
        for m in (1, 2, 3, 4, 6, 8, 10, 12, 16, 21, 25, 32, 40, 50, 64):
            reps = 20 if m <= 12 else 8
            ms_d, by_d = ctx.bench_kernel(nat.PROF_MDOT, m, reps=reps)
            ms_l, by_l = ctx.bench_kernel(nat.PROF_LINCOMB, m, reps=reps)
            row = dict(m=m, unfused_ms=ms_d + ms_l, mdot_gbs=by_d / ms_d * 1e-6, lincomb_gbs=by_l / ms_l * 1e-6)
            for stages in (2, 3, 4, 8):
                ctx.set_option("orth_mid_max_stages", stages)
                try:
                    ms, by = ctx.bench_kernel(nat.PROF_ORTHMID, m, reps=reps)
                    row[f"fused_ms_s{stages}"] = ms
                    row[f"fused_gbs_s{stages}"] = by / ms * 1e-6
                except nat.SpisError as exc:
                    row[f"fused_ms_s{stages}"] = None
            rows.append(row)
            print(json.dumps(row), flush=True)
    with open(args.out, "w") as fh:
        json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
