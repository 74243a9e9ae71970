#!/usr/bin/env python
"""Multi-GPU parity check (torchrun, NCCL): row-sharded CGMRES vs the single-GPU solve.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/dist_gpu_check.py [n_target] [transport] [lkdv|swe]

swe: every rank assembles only its own strip of the RT_2 x DG_0 system (problems/swe.py rows=...).
"""
import os
import sys
import time
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from structurepreservingiterativesolvers_b200 import solvers, wrappers  # noqa: E402
from structurepreservingiterativesolvers_b200.distributed import DistributedSession, TorchComm, cgmres_distributed  # noqa: E402
from structurepreservingiterativesolvers_b200.partition import FieldBlockPartition, StripPartition  # noqa: E402
from structurepreservingiterativesolvers_b200.problems import lkdv, swe  # noqa: E402


class Inv:
    def __init__(self, M, v, c):
        self.M, self.v, self.c = M, v, c


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = TorchComm(device=local)
    rank, world = comm.rank, comm.world
    warnings.simplefilter("ignore")
    n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    workload = sys.argv[3] if len(sys.argv) > 3 else "lkdv"
    if workload == "swe":
        M = swe.benchmark_size(n_target)
        part = StripPartition((swe.NU * M, swe.NR * M), M, world)
        dl, _ = swe.linforms(M=M, mlength=0.8 * M, rows=part.block_range(rank))
        ids = part.global_ids(rank)
        n = 12 * M * M
        tol = 1e-7
        cl_loc = wrappers.swe.conlist(dl, np.zeros(dl["b"].size))
        A_loc, b_loc, x0_loc = dl["A"], dl["b"], np.zeros(dl["b"].size)
        problem_mod = swe
        if rank == 0:
            d, _ = swe.linforms(M=M, mlength=0.8 * M)
            x0 = np.zeros(n)
            cl = wrappers.swe.conlist(d, x0)
    else:
        M = lkdv.benchmark_size(n_target)
        d, _ = lkdv.linforms(space="CG", M=M, mlength=0.8 * M)
        n = d["b"].size
        x0 = np.zeros(n)
        cl = wrappers.lkdv.conlist(d, x0)
        tol = 1e-6 * np.sqrt(n / 150)
        part = FieldBlockPartition(3, M, world)
        ids = part.global_ids(rank)
        cl_loc = [Inv(c.M.tocsr()[ids], np.asarray(c.v).reshape(-1)[ids], c.c) for c in cl]
        A_loc, b_loc, x0_loc = d["A"][ids], d["b"][ids], x0[ids]
        problem_mod = lkdv
    transport = sys.argv[2] if len(sys.argv) > 2 else "auto"
    if len(sys.argv) > 4 and sys.argv[4] == "soak":
        sys.exit(soak(comm, part, A_loc, b_loc, x0_loc, cl_loc, tol, transport, int(sys.argv[5]) if len(sys.argv) > 5 else 200))
    sess = DistributedSession(A_loc, b_loc, x0_loc, 50, part, comm, conlist=cl_loc, profile=True,
                              transport=transport)
    for rep in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        x_loc, info = cgmres_distributed(None, b_loc, x0_loc, 50, part, comm, tol=tol, contol=10, conlist=cl_loc,
                                         small_solver="kkt", timing=True, session=sess)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    xg = sess.gather(x_loc)
    ok = True
    if rank == 0:
        s1 = solvers.DeviceSession(d["A"], d["b"], x0, 50, conlist=cl, device=local)
        for _ in range(3):
            t0 = time.perf_counter()
            xs, infos = solvers.cgmres(d["A"], d["b"], x0, 50, tol=tol, contol=10, conlist=cl, small_solver="kkt", timing=True, session=s1)
            t_single = time.perf_counter() - t0
        print("single-GPU solve %.1f ms, timings %s" % (t_single * 1e3, {k: round(float(v), 5) for k, v in infos["timings"].items()}), flush=True)
        print("sharded timings %s" % {k: round(float(v), 5) for k, v in info["timings"].items()}, flush=True)
        rel = np.linalg.norm(xg - xs) / np.linalg.norm(xs)
        inv = problem_mod.compute_invariants(d, xg)
        dev = max(abs(inv["mass"] - d["m0"]) / abs(d["m0"]), abs(inv["energy"] - d["e0"]) / max(abs(d["e0"]), abs(d.get("mo0", 0.0))))
        ok = (info["steps"] == infos["steps"]) and rel <= 1e-10 and dev <= 1e-11
        print(f"world={world} n={n} steps={info['steps']} (single {infos['steps']}) rel.diff={rel:.2e} invariant dev={dev:.2e} "
              f"solve={dt*1e3:.1f} ms transport={sess.transport} collectives={comm.counts} halo={sess.plan.n_halo} -> {'OK' if ok else 'FAIL'}", flush=True)
        prof = sess.ctx.profile()
        print({k: (round(v['ms'], 2), v['launches']) for k, v in prof.items() if v['launches']}, flush=True)
    sess.close()
    # ---- the per-call path: a fresh DistributedSession per solve (the public call pattern) must not rebuild the
    # communicator or the sharding plan, and a constraint whose v is zero on every rank but one must still be reduced
    # on all of them (include/spis_b200.h, spis_constraint_define)
    if workload == "lkdv":
        vmask = np.zeros(n)
        own0 = part.global_ids(0)
        sel = own0[: own0.size // 3]                                # non-zero only on rank 0's nodes of the first field;
        vmask[sel] = np.where(np.arange(sel.size) % 2 == 0, 1.0, -1.0)   # alternating signs: v.x is O(1), so one ulp of the invariant is tiny
                                                                        # (the 'kkt' engine settles signs by ulps of |c|, smallsolve._settle_signs)
        # a constraint the solution (nearly) satisfies already, so that it does not amplify rounding: v.x = v.x_gmres
        box = [float(-(vmask @ xs)) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        cz = box[0]
        # (two LINEAR invariants: quadratic ones add their own SLSQP-level sensitivity, tests/golden/self_noise.json)
        cl2_loc = [cl_loc[0], Inv(0 * A_loc, vmask[ids], cz)]
        times = []
        for rep in range(3):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            x2_loc, info2 = cgmres_distributed(A_loc, b_loc, x0_loc, 50, part, comm, tol=tol, contol=10, conlist=cl2_loc,
                                               small_solver="kkt", timing=True, transport=transport)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        s2 = DistributedSession(A_loc, b_loc, x0_loc, 50, part, comm, conlist=cl2_loc, transport=transport)
        x2g = s2.gather(x2_loc)
        s2.close()
        if rank == 0:
            cl2 = [cl[0], Inv(0 * d["A"], vmask, cz)]
            xs2, infos2 = solvers.cgmres(d["A"], d["b"], x0, 50, tol=tol, contol=10, conlist=cl2, small_solver="kkt", timing=True, device=local)
            rel2 = np.linalg.norm(x2g - xs2) / np.linalg.norm(xs2)
            ok2 = info2["steps"] == infos2["steps"] and rel2 <= 1e-10 and comm.counts["plans_built"] <= 2 and comm.counts["peer_comms_built"] == (1 if sess.transport == "p2p" else 0)
            print(f"per-call sessions + zero-v-slice constraint: steps={info2['steps']} (single {infos2['steps']}) rel.diff={rel2:.2e} "
                  f"|v.x + c|={abs(vmask @ x2g + cz):.2e} e2e per call {[round(t * 1e3, 1) for t in times]} ms counts={comm.counts} -> {'OK' if ok2 else 'FAIL'}", flush=True)
            ok = ok and ok2
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def soak(comm, part, A_loc, b_loc, x0_loc, cl_loc, tol, transport, reps):
    """Many back-to-back sharded solves through fresh sessions (the NVLink flag protocol, the persistent communicator
    and the plan cache under repetition): every solve must return the bits of the first one."""
    ref, t0 = None, time.perf_counter()
    bad = 0
    for rep in range(reps):
        x_loc, info = cgmres_distributed(A_loc, b_loc, x0_loc, 50, part, comm, tol=tol, contol=10, conlist=cl_loc,
                                         small_solver="kkt", timing=True, transport=transport)
        sig = (info["steps"], float(np.asarray(x_loc) @ np.asarray(x_loc)), float(info["res"][-1]))
        if ref is None:
            ref = sig
        bad += sig != ref
    dt = time.perf_counter() - t0
    flag = torch.tensor([float(bad)], device="cuda")
    dist.all_reduce(flag)
    if comm.rank == 0:
        print(f"soak: {reps} sharded solves on {comm.world} GPUs, {dt / reps * 1e3:.2f} ms each, mismatching solves summed over ranks: {int(flag.item())}, "
              f"counts={comm.counts} -> {'OK' if flag.item() == 0 else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if flag.item() == 0 else 1


if __name__ == "__main__":
    main()
