import os, sys, time, warnings
sys.path.insert(0, '/root/repo')
import bench
from structurepreservingiterativesolvers_b200 import solvers
dic, x0, conlist, _ = bench.build_system(10_000_000)
sess = solvers.DeviceSession(dic["A"], dic["b"], x0, 50, conlist=conlist)
warnings.simplefilter("ignore")
def solve():
    return solvers.cgmres(dic["A"], dic["b"], x0, 50, tol=1e-6, contol=10, conlist=conlist, timing=True, small_solver="kkt", session=sess)
for rep in range(4):
    t0 = time.perf_counter(); x, info = solve(); t1 = time.perf_counter()
    print(f"plain #{rep}: {1e3*(t1-t0):.2f} ms  runtime {1e3*info['timings']['runtime']:.2f}")
with bench.ClockSampler(0) as clk:
    time.sleep(1.0)
    for rep in range(6):
        t0 = time.perf_counter(); x, info = solve(); t1 = time.perf_counter()
        print(f"sampler #{rep}: {1e3*(t1-t0):.2f} ms  runtime {1e3*info['timings']['runtime']:.2f}")
print(clk.summary())
for rep in range(3):
    t0 = time.perf_counter(); x, info = solve(); t1 = time.perf_counter()
    print(f"plain again #{rep}: {1e3*(t1-t0):.2f} ms  runtime {1e3*info['timings']['runtime']:.2f}")
