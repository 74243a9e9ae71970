#!/usr/bin/env python
"""ncu target: a few launches of single kernel classes on resident data (n = 1e7).

    ncu --set full -k regex:orth_mid ... python tools/ncu_kernels.py orthmid 16 21 50
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.device import KrylovContext

cls = {"orthmid": nat.PROF_ORTHMID, "mdot": nat.PROF_MDOT, "lincomb": nat.PROF_LINCOMB}[sys.argv[1]]
ms = [int(a) for a in sys.argv[2:]] or [21]
n = int(os.environ.get("SPIS_N", 10_000_050))
with KrylovContext(n, 64) as ctx:
    for m in ms:
        t, by = ctx.bench_kernel(cls, m, reps=1)
        print(sys.argv[1], m, "%.3f ms" % t, "%.0f GB/s" % (by / t * 1e-6))
