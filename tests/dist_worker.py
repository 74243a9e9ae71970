"""Worker for tests/test_distributed_cpu.py (launched by torch.distributed.run, gloo backend).

Each rank owns a mesh block of the system, uses the numpy stand-in context for its local kernels and
the REAL distributed.py / partition.py for everything between ranks; the gathered result must match
the single-process solve of the same system."""
import os
import sys
import warnings

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)

from fake_ctx import FakeKrylovContext  # noqa: E402
from structurepreservingiterativesolvers_b200 import solvers, wrappers  # noqa: E402
from structurepreservingiterativesolvers_b200.distributed import (DistributedSession, TorchComm,  # noqa: E402
                                                                  cgmres_distributed, gmres_distributed)
from structurepreservingiterativesolvers_b200.partition import ArrayPartition, FieldBlockPartition, StripPartition, take_rows  # noqa: E402
from structurepreservingiterativesolvers_b200.preconditioners import JacobiPreconditioner  # noqa: E402
from structurepreservingiterativesolvers_b200.problems import heat, lkdv, swe  # noqa: E402


class Inv:
    def __init__(self, M, v, c):
        self.M, self.v, self.c = M, v, c


def shard(conlist, part, rank):
    ids = part.global_ids(rank)
    return [Inv(take_rows(c.M, part, rank), np.asarray(c.v).reshape(-1)[ids], c.c) for c in conlist]


def main():
    dist.init_process_group("gloo")
    comm = TorchComm(device=None)
    rank, world = comm.rank, comm.world
    warnings.simplefilter("ignore")
    checks = []

    # ---- case 1: lkdv P1, field-blocked mesh partition, 3 constraints, x0 = 0
    d, _ = lkdv.linforms(space="CG", M=60)
    n = d["b"].size
    x0 = np.zeros(n)
    cl = wrappers.lkdv.conlist(d, x0)
    part = FieldBlockPartition(3, 60, world)
    ids = part.global_ids(rank)
    sess1 = solvers.DeviceSession(d["A"], d["b"], x0, 40, conlist=cl, ctx_factory=FakeKrylovContext)
    xs, infos = solvers.cgmres(d["A"], d["b"], x0, 40, tol=1e-6, conlist=cl, session=sess1, small_solver="kkt")
    xd, infod = cgmres_distributed(take_rows(d["A"], part, rank), d["b"][ids], x0[ids], 40, part, comm, tol=1e-6,
                                   conlist=shard(cl, part, rank), gather=True, small_solver="kkt",
                                   session=DistributedSession(take_rows(d["A"], part, rank), d["b"][ids], x0[ids], 40, part, comm,
                                                              conlist=shard(cl, part, rank), ctx_factory=FakeKrylovContext))
    checks.append(("lkdv steps", infod["steps"] == infos["steps"]))
    checks.append(("lkdv x", np.linalg.norm(xd - xs) <= 1e-9 * np.linalg.norm(xs)))
    checks.append(("lkdv res", abs(infod["res"][-1] - infos["res"][-1]) <= 1e-9 * np.linalg.norm(d["b"])))

    # ---- case 2: heat P1 (2-D coupling), arbitrary striped ownership, non-zero x0, Jacobi, v != 0
    h, _ = heat.linforms(M=10)
    n = h["b"].size
    x0 = 0.05 * np.sin(np.arange(n))
    cl = wrappers.heat.conlist(h, x0)
    owner = (np.arange(n) * world) // n
    owner = np.roll(owner, 7)                                   # non-contiguous ownership
    part = ArrayPartition(owner, world)
    ids = part.global_ids(rank)
    pre = JacobiPreconditioner(h["A"])
    sess2 = solvers.DeviceSession(h["A"], h["b"], x0, 25, conlist=cl, pre=pre, ctx_factory=FakeKrylovContext)
    xs, infos = solvers.cgmres(h["A"], h["b"], x0, 25, tol=1e-7, conlist=cl, pre=pre, session=sess2, small_solver="kkt")
    pre_loc = JacobiPreconditioner(diag=h["A"].diagonal()[ids])
    dsess = DistributedSession(take_rows(h["A"], part, rank), h["b"][ids], x0[ids], 25, part, comm,
                               conlist=shard(cl, part, rank), pre=pre_loc, ctx_factory=FakeKrylovContext)
    xd, infod = cgmres_distributed(None, h["b"][ids], x0[ids], 25, part, comm, tol=1e-7, conlist=shard(cl, part, rank),
                                   pre=pre_loc, gather=True, small_solver="kkt", session=dsess)
    checks.append(("heat steps", infod["steps"] == infos["steps"]))
    checks.append(("heat x", np.linalg.norm(xd - xs) <= 1e-9 * np.linalg.norm(xs)))
    checks.append(("heat halo used", world == 1 or dsess.plan.n_halo > 0))
    checks.append(("heat collectives", world == 1 or (comm.counts["allreduce"] > 0 and comm.counts["halo"] > 0)))

    # ---- case 3: plain FGMRES
    xs, infos = solvers.gmres(h["A"], h["b"], x0, 15, tol=1e-9, session=solvers.DeviceSession(h["A"], h["b"], x0, 15, ctx_factory=FakeKrylovContext))
    xd, infod = gmres_distributed(None, h["b"][ids], x0[ids], 15, part, comm, tol=1e-9, gather=True,
                                  session=DistributedSession(take_rows(h["A"], part, rank), h["b"][ids], x0[ids], 15, part, comm,
                                                             ctx_factory=FakeKrylovContext))
    checks.append(("gmres x", np.linalg.norm(xd - xs) <= 1e-10 * np.linalg.norm(xs)))

    # ---- case 4: swe RT_2 x DG_0, strips of squares, every rank assembles ONLY its own rows
    Ms = 10
    dg, _ = swe.linforms(M=Ms, mlength=0.8 * Ms)
    x0 = np.zeros(dg["b"].size)
    clg = wrappers.swe.conlist(dg, x0)
    xs, infos = solvers.cgmres(dg["A"], dg["b"], x0, 40, tol=1e-7, conlist=clg, small_solver="kkt",
                               session=solvers.DeviceSession(dg["A"], dg["b"], x0, 40, conlist=clg, ctx_factory=FakeKrylovContext))
    part = StripPartition((swe.NU * Ms, swe.NR * Ms), Ms, world)
    dl, _ = swe.linforms(M=Ms, mlength=0.8 * Ms, rows=part.block_range(rank))
    x0l = np.zeros(dl["b"].size)
    cll = wrappers.swe.conlist(dl, x0l)
    dsess = DistributedSession(dl["A"], dl["b"], x0l, 40, part, comm, conlist=cll, ctx_factory=FakeKrylovContext)
    xd, infod = cgmres_distributed(None, dl["b"], x0l, 40, part, comm, tol=1e-7, conlist=cll, gather=True,
                                   small_solver="kkt", session=dsess)
    inv = swe.compute_invariants(dg, xd)
    checks.append(("swe steps", infod["steps"] == infos["steps"]))
    checks.append(("swe x", np.linalg.norm(xd - xs) <= 1e-9 * np.linalg.norm(xs)))
    checks.append(("swe invariants", abs(inv["mass"] - dg["m0"]) <= 1e-11 * abs(dg["m0"]) and abs(inv["energy"] - dg["e0"]) <= 1e-11 * abs(dg["e0"])))
    checks.append(("swe halo is a strip boundary", world == 1 or 0 < dsess.plan.n_halo <= 2 * 12 * Ms))

    bad = [name for name, ok in checks if not ok]
    print(f"rank {rank}/{world}: {len(checks) - len(bad)} ok, failed: {bad}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
