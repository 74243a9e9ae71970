"""Row-sharded solver logic on CPU: partition / halo plan unit tests, and a world_size-2 gloo run."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sps

from structurepreservingiterativesolvers_b200.partition import (ArrayPartition, FieldBlockPartition, StripPartition,
                                                                localize, take_rows)
from structurepreservingiterativesolvers_b200.problems import heat, lkdv, swe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_field_block_partition_is_a_bijection(P):
    part = FieldBlockPartition(3, 50, P)
    seen = np.concatenate([part.global_ids(r) for r in range(P)])
    assert sorted(seen) == list(range(150))
    for r in range(P):
        g = part.global_ids(r)
        assert np.all(part.owner_of(g) == r)
        np.testing.assert_array_equal(part.local_of(g), np.arange(g.size))
        assert part.n_local(r) == g.size
    sizes = [part.n_local(r) for r in range(P)]
    assert max(sizes) - min(sizes) <= 3                              # balanced to one node per field


@pytest.mark.parametrize("P", [1, 2, 3, 5])
def test_strip_partition_matches_swe_numbering(P):
    M = 7
    part = StripPartition((swe.NU * M, swe.NR * M), M, P)
    assert part.n == 12 * M * M
    seen = np.concatenate([part.global_ids(r) for r in range(P)])
    assert sorted(seen) == list(range(part.n))
    for r in range(P):
        g = part.global_ids(r)
        np.testing.assert_array_equal(g, swe.strip_ids(M, *part.block_range(r)))     # what linforms(rows=...) assembles
        assert np.all(part.owner_of(g) == r)
        np.testing.assert_array_equal(part.local_of(g), np.arange(g.size))
        assert part.n_local(r) == g.size
    fb, sp = FieldBlockPartition(3, 11, P), StripPartition((1, 1, 1), 11, P)
    for r in range(P):
        np.testing.assert_array_equal(fb.global_ids(r), sp.global_ids(r))


def test_swe_rows_assembled_locally_equal_global_rows():
    M, P = 9, 4
    d, _ = swe.linforms(M=M, mlength=0.8 * M)
    part = StripPartition((swe.NU * M, swe.NR * M), M, P)
    for r in range(P):
        dl, _ = swe.linforms(M=M, mlength=0.8 * M, rows=part.block_range(r))
        ids = part.global_ids(r)
        assert abs(dl["A"] - d["A"][ids]).max() == 0 and abs(dl["L"] - d["L"][ids]).max() == 0
        np.testing.assert_array_equal(dl["b"], d["b"][ids])
        np.testing.assert_array_equal(dl["omega"], d["omega"][ids])
        assert dl["m0"] == d["m0"] and dl["e0"] == d["e0"]
        # a strip needs one row of squares from each neighbouring strip: its halo is O(M), not O(M^2)
        (A_loc,), plan = localize([dl["A"]], part, r)
        assert 0 < plan.n_halo <= 2 * 12 * M


@pytest.mark.parametrize("P", [2, 4])
def test_localized_spmv_reproduces_global_product(P):
    """Emulate the halo exchange in-process: every rank's local matrix times [owned | ghosts]."""
    for A, part in ((lkdv.linforms(space="CG", M=40)[0]["A"], FieldBlockPartition(3, 40, P)),
                    (heat.linforms(M=9)[0]["A"], ArrayPartition(np.arange(100) % P, P)),
                    (swe.linforms(M=8)[0]["A"], StripPartition((swe.NU * 8, swe.NR * 8), 8, P))):
        n = A.shape[0]
        rng = np.random.default_rng(0)
        x = rng.standard_normal(n)
        y = np.zeros(n)
        plans, locs = [], []
        for r in range(P):
            (A_loc,), plan = localize([take_rows(A, part, r)], part, r)
            plans.append(plan); locs.append(A_loc)
        for r in range(P):
            plans[r].set_send_side([plans[s].requests[r] for s in range(P)], part)
        for r in range(P):
            ids = part.global_ids(r)
            ghosts = np.concatenate([x[part.global_ids(s)][plans[s].send_idx[
                int(plans[s].send_counts[:r].sum()): int(plans[s].send_counts[:r + 1].sum())]] for s in range(P)])
            assert ghosts.size == plans[r].n_halo
            np.testing.assert_array_equal(ghosts, x[plans[r].ghost_gids])   # ordered by (owner, id)
            y[ids] = locs[r] @ np.concatenate([x[ids], ghosts])
        np.testing.assert_allclose(y, A @ x, rtol=1e-13, atol=1e-13)
        # 1-D lkdv: each rank needs one node per field from each neighbour
        if isinstance(part, FieldBlockPartition) and P == 2:
            assert plans[0].n_halo == 6


def test_shared_ghost_numbering_for_constraint_matrices():
    d, _ = lkdv.linforms(space="CG", M=30)
    part = FieldBlockPartition(3, 30, 3)
    mats = [take_rows(d["A"], part, 1), take_rows(d["L"] - d["M"], part, 1)]
    (A_loc, E_loc), plan = localize(mats, part, 1)
    assert A_loc.shape == E_loc.shape == (30, 30 + plan.n_halo)
    assert E_loc.indices.max() < A_loc.shape[1]


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_process():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=280)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "failed: []" in res.stdout
