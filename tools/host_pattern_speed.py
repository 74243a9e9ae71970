"""Speed of the host-side row-pattern detection (spis_host_find_patterns) on the 1e7 lkdv operator, per thread count."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200.problems import lkdv
lib = nat.load_library()
M = lkdv.benchmark_size(int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000)
d, _ = lkdv.linforms(space="CG", M=M, mlength=0.8 * M)
A = d["A"].tocsr()
n = A.shape[0]
ip = np.ascontiguousarray(A.indptr, dtype=np.int32); ci = np.ascontiguousarray(A.indices, dtype=np.int32); da = np.ascontiguousarray(A.data)
pid = np.zeros(n, dtype=np.uint16); rep = np.zeros(4096, dtype=np.int32)
npat = C.c_int(0); ml = C.c_int(0); ch = C.c_int64(0)
print("cpus", os.cpu_count(), "bytes", ip.nbytes + ci.nbytes + da.nbytes)
for nt in (1, 2, 4, 8, 16, 32):
    best = 1e9
    for rep_ in range(3):
        t = time.perf_counter()
        lib.spis_host_find_patterns(ip.ctypes.data_as(C.POINTER(C.c_int32)), ci.ctypes.data_as(C.POINTER(C.c_int32)), nat.dptr(da), n, n, 0, nt,
                                    pid.ctypes.data_as(C.POINTER(C.c_uint16)), rep.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(npat), C.byref(ml), C.byref(ch))
        best = min(best, time.perf_counter() - t)
    print("threads", nt, "npat", npat.value, "ms", round(best * 1e3, 2), "GB/s", round((ip.nbytes + ci.nbytes + da.nbytes) / best / 1e9, 1), flush=True)
