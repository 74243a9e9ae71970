#!/usr/bin/env python
"""Where does orth_mid_kernel lose bandwidth?  probe=1 streams tiles through smem without consuming them."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.device import KrylovContext
n = 10_000_050
with KrylovContext(n, 64) as ctx:
    for m in (1, 2, 4, 8, 16, 21, 26, 27, 32, 50, 54):
        for E in (1, 2):
            for stages in (2, 8):
                for probe in (0, 1):
                    ctx.set_option("orth_mid_force_e", E); ctx.set_option("orth_mid_max_stages", stages); ctx.set_option("orth_mid_probe", probe)
                    try:
                        ms, by = ctx.bench_kernel(nat.PROF_ORTHMID, m, reps=10)
                        print(f"m={m:2d} E={E} S<={stages} probe={probe}: {ms:.3f} ms {by/ms*1e-6:6.0f} GB/s", flush=True)
                    except nat.SpisError as e:
                        print(f"m={m:2d} E={E} S<={stages} probe={probe}: n/a")
