"""Row-sharded CGMRES over several B200s: one process per GPU, torch.distributed for the plumbing.

Every rank owns a mesh block of rows (partition.py) and a KrylovContext for them.  The C library
calls back into this module at its two communication points (include/spis_b200.h):

  * all-reduce of a small device buffer after each fused dot-product block (m, m, 1 doubles per
    Arnoldi step; 1 per residual; the constraint Gram rows) -> `torch.distributed.all_reduce` (NCCL,
    NVLS/NVSwitch), enqueued on the context's own CUDA stream so no host synchronisation is added;
  * halo exchange before each SpMV: the library packs the entries the neighbours need
    (halo_pack_kernel) and the callback moves them with batched NCCL send/recv straight into the
    ghost region of the input vector.

The Hessenberg column, residual norms and constraint terms come out identical on every rank (NCCL
all-reduce is bitwise identical across ranks), so each rank runs the k-dimensional host solve
redundantly and the control flow of solvers.cgmres stays rank-symmetric without broadcasting y.

The same code runs on CPU with the gloo backend and a numpy stand-in for the context
(tests/test_distributed_cpu.py) -- that covers the partitioning, the halo plan and the collective
ordering; the kernels themselves are covered by the single-GPU tests.
"""
from __future__ import annotations

import numpy as np

from . import _native as nat
from . import solvers
from .device import KrylovContext
from .partition import localize


class _DevBuf:
    """Lets torch view `count` doubles at a raw device pointer (no copy, no ownership)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8",
                                         "data": (int(ptr), False), "version": 3, "strides": None}


class TorchComm:
    """torch.distributed-backed collectives for one rank (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, device=None, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.cuda = device is not None
        self.device = device
        self.stream = None
        if self.cuda:
            torch.cuda.set_device(device)
            self.stream = torch.cuda.Stream(device=device)      # non-default: its handle is not NULL
        self._views = {}
        self.counts = {"allreduce": 0, "halo": 0}

    # -- views ----------------------------------------------------------------------------------
    def _view(self, buf, count):
        if isinstance(buf, np.ndarray):                          # CPU stand-in: share memory with numpy
            return self.torch.from_numpy(buf.reshape(-1)[:count])
        key = (int(buf), int(count))
        t = self._views.get(key)
        if t is None:
            t = self.torch.as_tensor(_DevBuf(buf, count), device=f"cuda:{self.device}")
            self._views[key] = t
        return t

    def stream_handle(self):
        return self.stream.cuda_stream if self.stream is not None else None

    def _on_stream(self):
        import contextlib
        return self.torch.cuda.stream(self.stream) if self.cuda else contextlib.nullcontext()

    # -- collectives ------------------------------------------------------------------------------
    def allreduce(self, buf, count):
        if self.world == 1:
            return
        self.counts["allreduce"] += 1
        with self._on_stream():
            self.dist.all_reduce(self._view(buf, count), op=self.dist.ReduceOp.SUM, group=self.group)

    def make_halo(self, plan):
        send_counts = [int(c) for c in plan.send_counts]
        recv_counts = [int(c) for c in plan.recv_counts]
        n_send, n_recv = sum(send_counts), sum(recv_counts)
        s_off = np.concatenate([[0], np.cumsum(send_counts)]).astype(int)
        r_off = np.concatenate([[0], np.cumsum(recv_counts)]).astype(int)
        dist = self.dist

        def halo(send_buf, recv_buf):
            if self.world == 1 or (n_send == 0 and n_recv == 0):
                return
            self.counts["halo"] += 1
            with self._on_stream():
                st = self._view(send_buf, n_send) if n_send else None
                rt = self._view(recv_buf, n_recv) if n_recv else None
                ops = []
                for peer in range(self.world):
                    if recv_counts[peer]:
                        ops.append(dist.P2POp(dist.irecv, rt[r_off[peer]:r_off[peer + 1]], peer, group=self.group))
                for peer in range(self.world):
                    if send_counts[peer]:
                        ops.append(dist.P2POp(dist.isend, st[s_off[peer]:s_off[peer + 1]], peer, group=self.group))
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        return halo

    def any_rank(self, flag):
        if self.world == 1:
            return bool(flag)
        t = self.torch.tensor([1.0 if flag else 0.0], dtype=self.torch.float64,
                              device=(f"cuda:{self.device}" if self.cuda else "cpu"))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(t.item() > 0)

    def exchange_requests(self, requests):
        """requests[s] = ids I need from rank s  ->  wanted_by[s] = ids rank s needs from me."""
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, [np.asarray(r, dtype=np.int64) for r in requests], group=self.group)
        return [gathered[s][self.rank] for s in range(self.world)]

    def setup_peer_memory(self, ctx, plan):
        """Switch a context to the NVLink peer-memory collectives (spis_xcomm_*): all-reduces fused
        into the reducing kernels, halo pushed straight into the neighbours' buffers.  Only the
        one-off exchange of IPC handles and ghost offsets goes through torch.distributed."""
        handle = ctx.xcomm_create(self.rank, self.world, plan.n_halo)
        info = [None] * self.world
        self.dist.all_gather_object(info, (handle, [int(c) for c in plan.recv_counts]), group=self.group)
        ctx.xcomm_connect(b"".join(h for h, _ in info))
        dest_rank, dest_off = [], []
        for peer in range(self.world):
            cnt = int(plan.send_counts[peer])
            if cnt:
                # my entries land after those of lower-ranked sources in the peer's ghost ordering
                base = sum(info[peer][1][:self.rank])
                dest_rank.append(np.full(cnt, peer, dtype=np.int32))
                dest_off.append(base + np.arange(cnt, dtype=np.int32))
        dest_rank = np.concatenate(dest_rank) if dest_rank else np.zeros(0, dtype=np.int32)
        dest_off = np.concatenate(dest_off) if dest_off else np.zeros(0, dtype=np.int32)
        ctx.xcomm_set_halo(dest_rank, dest_off, (plan.send_counts > 0).astype(np.int32),
                           (plan.recv_counts > 0).astype(np.int32))
        self.dist.barrier(group=self.group)          # every rank connected before the first collective

    def allgather_vec(self, local):
        parts = [None] * self.world
        self.dist.all_gather_object(parts, np.asarray(local), group=self.group)
        return parts


class DistributedSession(solvers.DeviceSession):
    """DeviceSession for one rank of a row-sharded system.

    A_rows / the constraints' M hold THIS RANK'S ROWS with GLOBAL column ids (local row order =
    part.global_ids(rank)); b, x0 and the constraints' v are the local pieces; the constraint
    constants c are global.  Dict-form (opaque callback) constraints and host-side preconditioners
    need the full Z on one host and are not supported here.
    """

    def __init__(self, A_rows, b_loc, x0_loc, k, part, comm, conlist=(), pre=None, *, orth=None,
                 spmv_format=None, profile=None, ctx_factory=KrylovContext, transport="auto"):
        """transport: 'p2p' = NVLink peer-memory collectives inside the kernels (GPUs on one node),
        'nccl' = torch.distributed callbacks, 'auto' = p2p on GPUs, callbacks otherwise."""
        self.comm, self.part = comm, part
        if transport == "auto":
            transport = "p2p" if (comm.cuda and comm.world > 1 and ctx_factory is KrylovContext) else "nccl"
        self.transport = transport
        rank = comm.rank
        conlist = list(conlist)
        for c in conlist:
            if solvers._classify_constraint(c) != "class":
                raise NotImplementedError("row-sharded solves support class-form constraints only")
        if pre is not None and not isinstance(pre, (solvers.JacobiPreconditioner, solvers.BlockJacobiPreconditioner)):
            raise NotImplementedError("row-sharded solves support None / Jacobi / block-Jacobi preconditioners")
        mats = [A_rows] + [c.M for c in conlist]
        local, plan = localize(mats, part, rank)
        plan.set_send_side(comm.exchange_requests(plan.requests), part)
        self.plan = plan
        A_loc = local[0]
        cons_loc = []
        for c, M_loc in zip(conlist, local[1:]):
            cons_loc.append(type("ShardedInvariant", (), {})())
            cons_loc[-1].M, cons_loc[-1].v, cons_loc[-1].c = M_loc, c.v, c.c

        def factory(n, kk, device=None):
            ctx = ctx_factory(n, kk, device=(comm.device if comm.cuda else 0), n_halo=plan.n_halo,
                              stream=comm.stream_handle())
            ctx.halo_set_plan(plan.send_idx)
            if transport == "p2p":
                comm.setup_peer_memory(ctx, plan)
            else:
                ctx.set_collectives(comm.allreduce, comm.make_halo(plan))
            return ctx

        super().__init__(A_loc, b_loc, x0_loc, k, conlist=cons_loc, pre=pre, orth=orth,
                         spmv_format=spmv_format, profile=profile, ctx_factory=factory)

    def _any_rank(self, flag):
        return self.comm.any_rank(flag)

    def gather(self, x_loc):
        """Assemble the global vector (global ordering) from the local pieces on every rank."""
        parts = self.comm.allgather_vec(x_loc)
        out = np.empty(self.part.n)
        for r, p in enumerate(parts):
            out[self.part.global_ids(r)] = p
        return out


def cgmres_distributed(A_rows, b_loc, x0_loc, k, part, comm, tol=1e-8, contol=10, conlist=(), pre=None,
                       timing=None, gather=False, **ext):
    """solvers.cgmres on a row-sharded system; returns this rank's piece of x (or the gathered
    global vector with gather=True) and the usual info dict (identical on every rank)."""
    sess = ext.pop("session", None) or DistributedSession(A_rows, b_loc, x0_loc, k, part, comm, conlist=conlist, pre=pre,
                                                          orth=ext.pop("orth", None), profile=ext.pop("profile", None))
    x_loc, info = solvers.cgmres(A_rows, b_loc, x0_loc, k, tol=tol, contol=contol, conlist=conlist, pre=pre,
                                 timing=timing, session=sess, **ext)
    return (sess.gather(x_loc) if gather else x_loc), info


def gmres_distributed(A_rows, b_loc, x0_loc, k, part, comm, tol=1e-50, pre=None, gather=False, **ext):
    sess = ext.pop("session", None) or DistributedSession(A_rows, b_loc, x0_loc, k, part, comm, pre=pre,
                                                          orth=ext.pop("orth", None), profile=ext.pop("profile", None))
    x_loc, info = solvers.gmres(A_rows, b_loc, x0_loc, k, tol=tol, pre=pre, session=sess, **ext)
    return (sess.gather(x_loc) if gather else x_loc), info
