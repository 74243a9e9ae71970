"""Why is the Arnoldi mdot slower inside a solve than alone?  21 pipelined steps on the 1e7 lkdv system with and without
the residual measurement riding in the SpMV / mdot tail, per-launch durations from profile mode; then mdot alone."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from structurepreservingiterativesolvers_b200 import _native as nat, solvers
from structurepreservingiterativesolvers_b200.problems import lkdv
warnings.simplefilter("ignore")
M = lkdv.benchmark_size(10_000_000)
d, _ = lkdv.linforms(space="CG", M=M, mlength=0.8 * M)
x0 = np.zeros(d["b"].size)
out = {}
for ride in (1, 0):
    sess = solvers.DeviceSession(d["A"], d["b"], x0, 50, conlist=(), profile=True)
    ctx = sess.ctx
    for rep in range(2):
        beta = sess.begin()
        ctx.pipe_begin(1e-30, True)
        ctx.reset_profile()
        for j in range(22):
            ctx.step_enqueue(j, bool(ride) and j >= 2, True)
            if j >= 3:
                ctx.step_wait(j - 3)
                if ride and j - 3 >= 2:
                    pass
        ctx.sync()
        ctx.profile()
        tr = ctx.profile_trace()
    by = {}
    for c, t0, dt in tr:
        by.setdefault(c, []).append(round(dt * 1e3, 1))
    out["ride" if ride else "plain"] = {k: v for k, v in by.items()}
    sess.close()
from structurepreservingiterativesolvers_b200.device import KrylovContext
with KrylovContext(d["b"].size, 44) as ctx:
    alone = {}
    for m in (18, 19, 20, 21, 22):
        ms, by = ctx.bench_kernel(nat.PROF_MDOT, m, reps=20)
        alone[m] = round(ms * 1e3, 1)
    out["mdot_alone_us"] = alone
    alone = {}
    for m in (18, 19, 20, 21, 22):
        ms, by = ctx.bench_kernel(nat.PROF_ORTHMID, m, reps=20)
        alone[m] = round(ms * 1e3, 1)
    out["orthmid_alone_us"] = alone
print(json.dumps(out))
