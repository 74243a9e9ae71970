"""Row-block sharding of the Krylov loop over several GPUs: pure numpy, no devices.

The unknowns of the reference's systems are ordered field-blocked, [u; v; w] (lkdv/refd.py:17,
swe/refd.py:24-25), so a naive contiguous row split would put every coupling of a row on another
GPU.  `FieldBlockPartition` gives rank r the same contiguous node range of EVERY
field; locally the unknowns stay field-blocked, so the SELL slices keep their uniform row lengths.

`localize` turns the rows a rank owns (global column numbering) into the local matrix the C ABI
expects -- owned columns first, then ghost columns ordered by (owner rank, global id) -- and
`HaloPlan` records who sends what.  The send side of the plan needs one exchange of index lists
between ranks, done by `distributed.py` with torch.distributed.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps


class ArrayPartition:
    """Arbitrary ownership given as an array owner[i] in [0, P); local order = increasing global id."""

    def __init__(self, owner, P):
        self.owner = np.asarray(owner, dtype=np.int64)
        self.P = int(P)
        self.n = self.owner.size
        self._local = np.empty(self.n, dtype=np.int64)
        self._globals = []
        for r in range(self.P):
            g = np.flatnonzero(self.owner == r)
            self._globals.append(g)
            self._local[g] = np.arange(g.size)

    def owner_of(self, g):
        return self.owner[np.asarray(g, dtype=np.int64)]

    def local_of(self, g):
        return self._local[np.asarray(g, dtype=np.int64)]

    def global_ids(self, r):
        return self._globals[r]

    def n_local(self, r):
        return self._globals[r].size


class FieldBlockPartition:
    """nfields fields of N nodes each, global index = f*N + node; rank r owns nodes
    [starts[r], starts[r+1]) of every field; local index = f*n_nodes_r + (node - starts[r])."""

    def __init__(self, nfields, N, P):
        self.nfields, self.N, self.P = int(nfields), int(N), int(P)
        self.n = self.nfields * self.N
        self.starts = (np.arange(self.P + 1, dtype=np.int64) * self.N) // self.P

    def _node(self, g):
        g = np.asarray(g, dtype=np.int64)
        return g // self.N, g % self.N

    def owner_of(self, g):
        _, node = self._node(g)
        return np.searchsorted(self.starts, node, side="right") - 1

    def local_of(self, g):
        f, node = self._node(g)
        r = np.searchsorted(self.starts, node, side="right") - 1
        width = self.starts[r + 1] - self.starts[r]
        return f * width + (node - self.starts[r])

    def global_ids(self, r):
        nodes = np.arange(self.starts[r], self.starts[r + 1], dtype=np.int64)
        return (np.arange(self.nfields, dtype=np.int64)[:, None] * self.N + nodes[None, :]).reshape(-1)

    def n_local(self, r):
        return int(self.nfields * (self.starts[r + 1] - self.starts[r]))

    def owned_ranges(self, r):
        """[(global_lo, global_hi, local_start)]: the owned unknowns as contiguous global ranges."""
        w = int(self.starts[r + 1] - self.starts[r])
        return [(f * self.N + int(self.starts[r]), f * self.N + int(self.starts[r + 1]), f * w) for f in range(self.nfields)]


class StripPartition:
    """Several fields of DIFFERENT densities over the same `nblocks` mesh blocks (2-D strips).

    Field f stores block_sizes[f] unknowns per mesh block, block-major; global index =
    field_off[f] + block*block_sizes[f] + within.  Rank r owns blocks [starts[r], starts[r+1]) of
    every field; its local order is field-blocked again.  swe (problems/swe.py): mesh block = one
    row of squares, block_sizes = (10 M, 2 M) for (velocity, density); FieldBlockPartition is the
    special case block_sizes = (1, ..., 1)."""

    def __init__(self, block_sizes, nblocks, P):
        self.bs = np.asarray(block_sizes, dtype=np.int64)
        self.nblocks, self.P = int(nblocks), int(P)
        self.field_off = np.concatenate([[0], np.cumsum(self.bs * self.nblocks)]).astype(np.int64)
        self.n = int(self.field_off[-1])
        self.starts = (np.arange(self.P + 1, dtype=np.int64) * self.nblocks) // self.P

    def _decode(self, g):
        g = np.asarray(g, dtype=np.int64)
        f = np.searchsorted(self.field_off, g, side="right") - 1
        loc = g - self.field_off[f]
        return f, loc // self.bs[f], loc % self.bs[f]

    def owner_of(self, g):
        _, blk, _ = self._decode(g)
        return np.searchsorted(self.starts, blk, side="right") - 1

    def local_of(self, g):
        f, blk, within = self._decode(g)
        r = np.searchsorted(self.starts, blk, side="right") - 1
        width = self.starts[r + 1] - self.starts[r]
        before = np.concatenate([[0], np.cumsum(self.bs)])[f]          # unknowns per block of earlier fields
        return before * width + (blk - self.starts[r]) * self.bs[f] + within

    def global_ids(self, r):
        b0, b1 = self.starts[r], self.starts[r + 1]
        return np.concatenate([self.field_off[f] + np.arange(b0 * self.bs[f], b1 * self.bs[f], dtype=np.int64)
                               for f in range(self.bs.size)])

    def n_local(self, r):
        return int(self.bs.sum() * (self.starts[r + 1] - self.starts[r]))

    def block_range(self, r):
        return int(self.starts[r]), int(self.starts[r + 1])

    def owned_ranges(self, r):
        b0, b1 = int(self.starts[r]), int(self.starts[r + 1])
        out, lstart = [], 0
        for f in range(self.bs.size):
            lo = int(self.field_off[f] + b0 * self.bs[f])
            hi = int(self.field_off[f] + b1 * self.bs[f])
            out.append((lo, hi, lstart))
            lstart += hi - lo
        return out


class HaloPlan:
    """Ghost layout of one rank: ghost global ids ordered by (owner, id), counts per source rank;
    the send side (send_idx, send_counts) is filled in after the ranks have exchanged requests."""

    def __init__(self, rank, P, ghost_gids, ghost_owner):
        self.rank, self.P = rank, P
        self.ghost_gids = ghost_gids
        self.recv_counts = np.bincount(ghost_owner, minlength=P).astype(np.int64)
        self.requests = [ghost_gids[ghost_owner == s] for s in range(P)]      # what I need from s
        self.send_idx = np.zeros(0, dtype=np.int32)
        self.send_counts = np.zeros(P, dtype=np.int64)

    @property
    def n_halo(self):
        return int(self.ghost_gids.size)

    def set_send_side(self, wanted_by, part):
        """wanted_by[s] = global ids rank s needs from this rank."""
        self.send_counts = np.array([len(w) for w in wanted_by], dtype=np.int64)
        if self.send_counts.sum():
            self.send_idx = np.concatenate([part.local_of(np.asarray(w, dtype=np.int64)) for w in wanted_by]).astype(np.int32)
        else:
            self.send_idx = np.zeros(0, dtype=np.int32)


def localize(mats, part, rank):
    """mats: list of CSR matrices holding THIS RANK'S ROWS (local row order = part.global_ids(rank))
    with GLOBAL column ids.  Returns (local matrices with shape (n_r, n_r + n_halo), HaloPlan); the
    ghost numbering is shared by all matrices (union of their off-rank columns)."""
    n_r = part.n_local(rank)
    mats = [sps.csr_matrix(m) for m in mats]
    for m in mats:
        if m.shape[0] != n_r:
            raise ValueError(f"rank {rank} owns {n_r} rows, matrix has {m.shape[0]}")
    # Column classification.  Partitions that own contiguous global ranges (mesh blocks) are handled with
    # range comparisons on the int32 index arrays as they are -- a 1e8-unknown strip has 6e8 entries, and
    # the generic path (int64 copies, searchsorted per entry) costs minutes and tens of GB there.
    ranges = part.owned_ranges(rank) if hasattr(part, "owned_ranges") else None
    off, news = [], []
    for m in mats:
        cols = m.indices
        if ranges is not None:
            new = np.full(cols.size, -1, dtype=np.int32)
            for lo, hi, lstart in ranges:
                sel = (cols >= lo) & (cols < hi)
                new[sel] = cols[sel] - (lo - lstart)
            own = new >= 0
        else:
            c64 = cols.astype(np.int64)
            own = part.owner_of(c64) == rank
            new = np.full(cols.size, -1, dtype=np.int64)
            new[own] = part.local_of(c64[own])
        news.append((new, own))
        off.append(np.unique(cols[~own]).astype(np.int64))
    ghost = np.unique(np.concatenate(off)) if off else np.zeros(0, dtype=np.int64)
    gowner = part.owner_of(ghost) if ghost.size else np.zeros(0, dtype=np.int64)
    order = np.lexsort((ghost, gowner))                   # by owner, then global id
    ghost, gowner = ghost[order], gowner[order]
    by_gid = np.argsort(ghost, kind="stable")
    ghost_sorted = ghost[by_gid]
    out = []
    for m, (new, own) in zip(mats, news):
        if (~own).any():
            pos = np.searchsorted(ghost_sorted, m.indices[~own].astype(np.int64))
            new[~own] = n_r + by_gid[pos]
        out.append(sps.csr_matrix((m.data, new.astype(np.int32, copy=False), m.indptr), shape=(n_r, n_r + ghost.size)))
    return out, HaloPlan(rank, part.P, ghost, gowner)


def take_rows(A, part, rank):
    """Rows of a GLOBAL matrix owned by `rank`, in local order, global column ids (small cases/tests;
    large systems assemble their local rows directly)."""
    return sps.csr_matrix(A)[part.global_ids(rank)]
