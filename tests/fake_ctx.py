"""numpy stand-in for KrylovContext, for CPU-only tests of the HOST logic in solvers.py.

Test infrastructure only: it lets `-m "not gpu"` runs exercise the control flow of
gmres / cgmres / cgmres_p (phase switching, quirks, history, timing dict, lookahead ordering,
constraint bookkeeping) without a GPU.  It mirrors the arithmetic the CUDA library performs
(CGS2 orthogonalisation, incremental constraint terms), not the reference's.
The product never imports this file.
"""
import numpy as np
import scipy.sparse as sps

from structurepreservingiterativesolvers_b200 import _native as nat


class FakeKrylovContext:
    def __init__(self, n, k_max, device=0, n_halo=0, stream=None):
        self.n, self.k_max = n, k_max
        self.generation = 0
        self.closed = False
        self.opts = {"orth": nat.ORTH_CGS2}
        self.mats = {}
        self.vecs = {}
        self.pre_kind = nat.PRE_NONE
        self.blocks = None
        self.cons = {}
        self.V = np.zeros((k_max + 1, n))
        self.Zs = np.zeros((k_max, n))
        self.log = []
        self._pending = None
        self.allreduce = None

    # ---- plumbing
    def close(self):
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key, value):
        self.opts[key] = value

    def info(self, key):
        return {"n": self.n, "k_max": self.k_max}.get(key, 0)

    def upload_matrix(self, slot, A):
        self.mats[slot] = sps.csr_matrix(A)

    def upload_vec(self, which, v):
        self.vecs[which] = np.array(v, dtype=float)

    def upload_blocks(self, blocks, sb, sf):
        self.blocks = (np.array(blocks), sb, sf)

    def set_precond(self, kind):
        self.pre_kind = kind

    def sync(self):
        pass

    # ---- Krylov
    def _Z(self):
        return self.V if self.pre_kind == nat.PRE_NONE else self.Zs

    def _apply_pre(self, q):
        if self.pre_kind == nat.PRE_JACOBI:
            return self.vecs[nat.VEC_PRE_DIAG] * q
        if self.pre_kind == nat.PRE_CSR:
            return self.mats[nat.SLOT_PRE] @ q
        if self.pre_kind == nat.PRE_BLOCK:
            blocks, sb, sf = self.blocks
            nblk, bs, _ = blocks.shape
            idx = np.arange(nblk)[:, None] * sb + np.arange(bs)[None, :] * sf
            z = np.zeros_like(q)
            z[idx] = np.einsum("irc,ic->ir", blocks, q[idx])
            return z
        raise AssertionError

    def solve_begin(self):
        A, b, x0 = self.mats[nat.SLOT_A], self.vecs[nat.VEC_B], self.vecs[nat.VEC_X0]
        self.r0 = b - A @ x0
        beta = np.sqrt(self.r0 @ self.r0)
        self.V[:] = 0
        self.V[0] = self.r0 / beta
        self.generation += 1
        for c in self.cons.values():
            c["done"] = 0
            c["T1"] = np.zeros(self.k_max)
            c["T2"] = np.zeros((self.k_max, self.k_max))
            c["MZ"] = np.zeros((self.k_max, self.n))
        self.log.append(("begin",))
        return float(beta)

    def arnoldi_launch(self, j):
        assert self._pending is None
        self.log.append(("launch", j))
        m = j + 1
        if self.pre_kind not in (nat.PRE_NONE, nat.PRE_HOST):
            self.Zs[j] = self._apply_pre(self.V[j])
        w = self.mats[nat.SLOT_A] @ self._Z()[j]
        Vm = self.V[:m]
        orth = self.opts.get("orth", nat.ORTH_CGS2)
        if orth == nat.ORTH_MGS:
            h = np.zeros(m)
            for i in range(m):
                h[i] = Vm[i] @ w
                w = w - h[i] * Vm[i]
        else:
            h = Vm @ w
            w = w - Vm.T @ h
            if orth == nat.ORTH_CGS2:
                h2 = Vm @ w
                w = w - Vm.T @ h2
                h = h + h2
        nrm = np.sqrt(w @ w)
        if nrm != 0:
            self.V[j + 1] = w / nrm
        else:
            self.V[j + 1] = w
        self._pending = (j, np.concatenate([h, [nrm]]))

    def arnoldi_wait(self, j):
        pj, col = self._pending
        assert pj == j
        self._pending = None
        self.log.append(("wait", j))
        return col

    def arnoldi_step(self, j):
        self.arnoldi_launch(j)
        return self.arnoldi_wait(j)

    def form_iterate(self, y):
        y = np.asarray(y, dtype=float)
        self.X = self.vecs[nat.VEC_X0] + self._Z()[: y.size].T @ y

    def iterate_residual(self, y):
        self.form_iterate(y)
        self.log.append(("iterate", len(y)))
        r = self.mats[nat.SLOT_A] @ self.X - self.vecs[nat.VEC_B]
        return float(np.sqrt(r @ r))

    # ---- constraints
    def constraint_define(self, c, slot, v, cc):
        self.cons[c] = {"slot": slot, "v": None if v is None else np.array(v, dtype=float), "c": cc, "done": 0,
                        "T1": np.zeros(self.k_max), "T2": np.zeros((self.k_max, self.k_max)),
                        "MZ": np.zeros((self.k_max, self.n))}

    def constraint_terms(self, c, m):
        C = self.cons[c]
        x0 = self.vecs[nat.VEC_X0]
        Z = self._Z()
        M = self.mats[C["slot"]] if C["slot"] >= 0 else None
        t0 = C["c"]
        if M is not None:
            t0 += 0.5 * x0 @ (M @ x0)
        if C["v"] is not None:
            t0 += C["v"] @ x0
        for col in range(C["done"], m):
            t1 = 0.0
            if M is not None:
                C["MZ"][col] = M @ Z[col]
                C["T2"][: col + 1, col] = 0.5 * (Z[: col + 1] @ C["MZ"][col])
                C["T2"][col, :col] = 0.5 * (C["MZ"][:col] @ Z[col])
                t1 += x0 @ C["MZ"][col]
            if C["v"] is not None:
                t1 += C["v"] @ Z[col]
            C["T1"][col] = t1
        C["done"] = max(C["done"], m)
        self.log.append(("terms", c, m))
        return float(t0), C["T1"][:m].copy(), C["T2"][:m, :m].copy()

    # ---- downloads
    def download(self, which, j=0, pinned=False):
        if which == nat.VEC_R0:
            return self.r0.copy()
        if which == nat.VEC_X:
            return self.X.copy()
        if which == nat.VEC_Q:
            return self.V[j].copy()
        if which == nat.VEC_Z:
            return self._Z()[j].copy()
        raise KeyError(which)

    def download_Z(self, j0, j1):
        return self._Z()[j0:j1].copy()

    def host_pre_get(self, j):
        return self.V[j].copy()

    def host_pre_put(self, j, z):
        self.Zs[j] = z

    def profile(self):
        return {}

    def reset_profile(self):
        pass
