import os, sys, time, warnings
sys.path.insert(0, '/root/repo')
import bench
from structurepreservingiterativesolvers_b200 import solvers
dic, x0, conlist, _ = bench.build_system(10_000_000)
mats = bench.pin_inputs(dic, x0, conlist)
warnings.simplefilter("ignore")
for rep in range(8):
    t0 = time.perf_counter()
    x, info = solvers.cgmres(mats[0], mats[1], mats[2], 50, tol=1e-6, conlist=mats[3], timing=True, small_solver="kkt")
    t1 = time.perf_counter()
    print(f"== e2e #{rep}: {1e3*(t1-t0):.1f} ms, loop runtime {1e3*info['timings']['runtime']:.1f} ms", file=sys.stderr)
    del x, info
