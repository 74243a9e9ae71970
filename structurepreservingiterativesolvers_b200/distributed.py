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

import os
import collections
import ctypes as C

import numpy as np
import scipy.sparse as sps

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
        self.counts = {"allreduce": 0, "halo": 0, "plans_built": 0, "peer_comms_built": 0}
        self._peer = None                    # persistent NVLink communicator: (handle, red_cap, halo_cap)
        self._plans = collections.OrderedDict()

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

    def any_rank_many(self, flags):
        """Logical OR over all ranks of several yes/no decisions, ONE collective for all of them."""
        flags = [bool(f) for f in flags]
        if self.world == 1 or not flags:
            return flags
        if self._peer is not None:
            vals = np.array([1.0 if f else 0.0 for f in flags], dtype=np.float64)
            lib = nat.load_library()
            rc = lib.spis_comm_allreduce(self._peer[0], nat.dptr(vals), vals.size)
            if rc != nat.OK:
                raise nat.SpisError(rc, lib.spis_last_global_error().decode())
            return [v > 0 for v in vals]
        t = self.torch.tensor([1.0 if f else 0.0 for f in flags], dtype=self.torch.float64,
                              device=(f"cuda:{self.device}" if self.cuda else "cpu"))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return [bool(v > 0) for v in t.tolist()]

    # -- persistent peer-memory communicator --------------------------------------------------------
    def peer_comm(self, red_need, halo_need):
        """The process-wide NVLink communicator (spis_comm_*): comm buffer, IPC mappings and flag counters are set
        up ONCE; every later session only attaches its context.  `red_need` / `halo_need` must be the same on all
        ranks (k_max + 5 and the largest ghost count of any rank, which the sharding plan records): the communicator
        is rebuilt, collectively, only when one of them outgrows it."""
        lib = nat.load_library()
        if self._peer is not None and self._peer[1] >= red_need and self._peer[2] >= halo_need:
            return self._peer[0]
        if self._peer is not None:
            self.dist.barrier(group=self.group)              # nobody is inside a collective of the old buffer
            lib.spis_comm_destroy(self._peer[0])
            self._peer = None
        red_cap = max(4096, int(red_need))
        halo_cap = max(1 << 16, 2 * int(halo_need))
        handle = C.create_string_buffer(64)
        comm = C.c_void_p()
        rc = lib.spis_comm_create(self.device, self.rank, self.world, red_cap, halo_cap, handle, 64, C.byref(comm))
        if rc != nat.OK:
            raise nat.SpisError(rc, lib.spis_last_global_error().decode())
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle.raw, group=self.group)
        rc = lib.spis_comm_connect(comm, C.c_char_p(b"".join(handles)))
        if rc != nat.OK:
            raise nat.SpisError(rc, lib.spis_last_global_error().decode())
        self.dist.barrier(group=self.group)                  # every rank connected before the first collective
        self._peer = (comm, red_cap, halo_cap)
        self.counts["peer_comms_built"] += 1
        return comm

    def close(self):
        if self._peer is not None:
            nat.load_library().spis_comm_destroy(self._peer[0])
            self._peer = None

    # -- sharding plan: column localisation + halo lists, cached per system ---------------------------
    def sharded_plan(self, mats, part):
        """Local matrices (owned columns first, then ghosts) and the halo plan for THIS rank's rows of `mats`.
        Building it costs host work on every index array plus two exchanges of index lists; the reference's call pattern
        is one solver call per time step on the SAME sparsity structure (lkdv/Evolve.py:39-56), so the plan is kept,
        keyed by the identity of the index arrays, and a later call only pairs the cached local column indices with
        the new value arrays."""
        mats = [sps.csr_matrix(m) for m in mats]
        key = (id(part),) + tuple((m.shape, int(m.nnz), m.indptr.ctypes.data, m.indices.ctypes.data) for m in mats)
        entry = self._plans.get(key)
        if entry is None:
            local, plan = localize(mats, part, self.rank)
            plan.set_send_side(self.exchange_requests(plan.requests), part)
            info = [None] * self.world
            self.dist.all_gather_object(info, [int(c) for c in plan.recv_counts], group=self.group)
            dest_rank, dest_off = [], []
            for peer in range(self.world):
                cnt = int(plan.send_counts[peer])
                if cnt:
                    base = sum(info[peer][:self.rank])      # my entries land after those of lower-ranked sources
                    dest_rank.append(np.full(cnt, peer, dtype=np.int32))
                    dest_off.append(base + np.arange(cnt, dtype=np.int32))
            entry = {"indices": [m.indices for m in local], "ncols": local[0].shape[1] if local else 0, "plan": plan,
                     "dest_rank": np.concatenate(dest_rank) if dest_rank else np.zeros(0, dtype=np.int32),
                     "dest_off": np.concatenate(dest_off) if dest_off else np.zeros(0, dtype=np.int32),
                     "send_to": (plan.send_counts > 0).astype(np.int32), "recv_from": (plan.recv_counts > 0).astype(np.int32),
                     "halo_max": max(sum(c) for c in info), "keep": [(m.indptr, m.indices) for m in mats]}
            self._plans[key] = entry
            self.counts["plans_built"] += 1
            while len(self._plans) > 4:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        n_r = part.n_local(self.rank)
        local = [sps.csr_matrix((m.data, idx, m.indptr), shape=(n_r, entry["ncols"]), copy=False)
                 for m, idx in zip(mats, entry["indices"])]
        return local, entry

    def clear_plans(self):
        self._plans.clear()

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
        local, entry = comm.sharded_plan(mats, part)          # cached per sparsity structure
        plan = entry["plan"]
        self.plan = plan
        A_loc = local[0]
        cons_loc = []
        for c, M_loc in zip(conlist, local[1:]):
            cons_loc.append(type("ShardedInvariant", (), {})())
            cons_loc[-1].M, cons_loc[-1].v, cons_loc[-1].c = M_loc, c.v, c.c
        peer = comm.peer_comm(int(k) + 5, entry["halo_max"]) if transport == "p2p" else None
        # every decision that changes the sequence of device reductions is taken by ALL ranks together, and all of
        # them in ONE collective: is x0 zero? which constraint matrices are identically zero (`0*A`)? which v?
        keys, flags = ["x0"], [nat.any_nonzero(nat.as_f64(x0_loc))]
        for idx, c in enumerate(cons_loc):
            M = c.M
            keys.append(("M", idx)); flags.append(bool(M.nnz != 0 and nat.any_nonzero(M.data)))
            keys.append(("v", idx)); flags.append(nat.any_nonzero(nat.as_f64(c.v)))
        self._preflags = dict(zip(keys, comm.any_rank_many(flags)))

        def factory(n, kk, device=None):
            ctx = ctx_factory(n, kk, device=(comm.device if comm.cuda else 0), n_halo=plan.n_halo,
                              stream=comm.stream_handle())
            ctx.halo_set_plan(plan.send_idx)
            # the ranks of one node share its cores: the host threads that look for row patterns are divided among them
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", comm.world) or comm.world)
            ctx.set_option("host_threads", max(1, min(16, (os.cpu_count() or 4) // max(1, local_world))))
            if transport == "p2p":
                ctx.attach_comm(peer)
                ctx.xcomm_set_halo(entry["dest_rank"], entry["dest_off"], entry["send_to"], entry["recv_from"])
            else:
                ctx.set_collectives(comm.allreduce, comm.make_halo(plan))
            return ctx

        super().__init__(A_loc, b_loc, x0_loc, k, conlist=cons_loc, pre=pre, orth=orth,
                         spmv_format=spmv_format, profile=profile, ctx_factory=factory)

    def _any_rank(self, flag, key=None):
        if key is not None and key in self._preflags:
            return self._preflags[key]
        return self.comm.any_rank(flag() if callable(flag) else flag)

    def gather(self, x_loc):
        """Assemble the global vector (global ordering) from the local pieces on every rank."""
        parts = self.comm.allgather_vec(x_loc)
        out = np.empty(self.part.n)
        for r, p in enumerate(parts):
            out[self.part.global_ids(r)] = p
        return out


def _own_session(ext, make):
    sess = ext.pop("session", None)
    if sess is not None:
        return sess, False
    sess = make()
    sess._owned_by_solver = True          # solvers._finish_history / _abandon manage its lifetime
    return sess, True


def cgmres_distributed(A_rows, b_loc, x0_loc, k, part, comm, tol=1e-8, contol=10, conlist=(), pre=None,
                       timing=None, gather=False, transport="auto", **ext):
    """solvers.cgmres on a row-sharded system; returns this rank's piece of x (or the gathered
    global vector with gather=True) and the usual info dict (identical on every rank).  Without `session=` a
    DistributedSession is built for the call: the communicator and the sharding plan it needs are cached in `comm`,
    so from the second call on a system with the same sparsity structure this is upload time only."""
    orth, profile = ext.pop("orth", None), ext.pop("profile", None)
    sess, own = _own_session(ext, lambda: DistributedSession(A_rows, b_loc, x0_loc, k, part, comm, conlist=conlist, pre=pre,
                                                             orth=orth, profile=profile, transport=transport))
    try:
        x_loc, info = solvers.cgmres(A_rows, b_loc, x0_loc, k, tol=tol, contol=contol, conlist=conlist, pre=pre,
                                     timing=timing, session=sess, **ext)
        return (sess.gather(x_loc) if gather else x_loc), info
    except BaseException:
        if own:
            solvers._abandon(sess)
        raise


def gmres_distributed(A_rows, b_loc, x0_loc, k, part, comm, tol=1e-50, pre=None, gather=False, transport="auto", **ext):
    orth, profile = ext.pop("orth", None), ext.pop("profile", None)
    sess, own = _own_session(ext, lambda: DistributedSession(A_rows, b_loc, x0_loc, k, part, comm, pre=pre, orth=orth,
                                                             profile=profile, transport=transport))
    try:
        x_loc, info = solvers.gmres(A_rows, b_loc, x0_loc, k, tol=tol, pre=pre, session=sess, **ext)
        return (sess.gather(x_loc) if gather else x_loc), info
    except BaseException:
        if own:
            solvers._abandon(sess)
        raise
