import os, time, numpy as np, threading, ctypes, sys
sys.path.insert(0, '/root/repo')
print("cpu.max:", open('/sys/fs/cgroup/cpu.max').read().strip() if os.path.exists('/sys/fs/cgroup/cpu.max') else 'n/a')
print("affinity:", len(os.sched_getaffinity(0)), "nproc", os.cpu_count())
from structurepreservingiterativesolvers_b200 import _native as nat
lib = nat.load_library()
a = np.zeros(60_000_000)
for rep in range(3):
    t=time.perf_counter(); r=nat.any_nonzero(a); dt=time.perf_counter()-t
    print("host scan 480MB (16 thr): %.1f ms -> %.1f GB/s"%(dt*1e3, 0.48/dt))
import torch
ap = torch.from_numpy(a).pin_memory().numpy()
from structurepreservingiterativesolvers_b200.device import KrylovContext
with KrylovContext(1000, 2) as ctx:
    for rep in range(3):
        t=time.perf_counter(); r=ctx.any_nonzero(ap); dt=time.perf_counter()-t
        print("pinned DMA scan 480MB: %.1f ms -> %.1f GB/s"%(dt*1e3, 0.48/dt))
    def bg():
        ctx.use_aux_stream(True)
        t=time.perf_counter(); r=ctx.any_nonzero(a); dt=time.perf_counter()-t
        print("helper-thread host scan 480MB (4 thr): %.1f ms"%(dt*1e3))
        ctx.use_aux_stream(False)
    th=threading.Thread(target=bg); th.start(); th.join()
