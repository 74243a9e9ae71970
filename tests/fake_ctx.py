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
        self.n, self.k_max, self.n_halo = n, k_max, n_halo
        self.nt = n + n_halo              # vectors carry their ghost entries at [n, n + n_halo)
        self.send_idx = np.zeros(0, dtype=np.int64)
        self.halo_cb = None
        self.generation = 0
        self.closed = False
        self.opts = {"orth": nat.ORTH_CGS2}
        self.mats = {}
        self.vecs = {nat.VEC_X0: np.zeros(n)}
        self.pre_kind = nat.PRE_NONE
        self.blocks = None
        self.cons = {}
        self.V = np.zeros((k_max + 1, self.nt))
        self.Zs = np.zeros((k_max, self.nt))
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

    def use_aux_stream(self, on):
        """DeviceSession stages the constraints from a helper thread when the context offers this."""
        import threading
        if on:
            self.aux_switches = getattr(self, "aux_switches", 0) + 1          # one per helper thread
            self.aux_on_caller_thread = getattr(self, "aux_on_caller_thread", False) or \
                threading.current_thread() is threading.main_thread()

    def info(self, key):
        if key == "device_pipeline":
            return int(self.opts.get("orth", nat.ORTH_CGS2) == nat.ORTH_CGS2 and self.pre_kind != nat.PRE_HOST
                       and self.allreduce is None and self.halo_cb is None
                       and (self.opts.get("fuse_iterate", 1) or self.pre_kind != nat.PRE_NONE))
        if key == "pipe_lag":
            return int(self.pre_kind == nat.PRE_NONE)
        if key == "can_fuse_iterate":
            return int(self.pre_kind == nat.PRE_NONE and self.opts.get("orth", nat.ORTH_CGS2) == nat.ORTH_CGS2
                       and self.opts.get("fuse_iterate", 1))
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

    def set_collectives(self, allreduce, halo):
        self.allreduce, self.halo_cb = allreduce, halo

    def halo_set_plan(self, send_idx):
        self.send_idx = np.asarray(send_idx, dtype=np.int64)

    def _ar(self, arr):
        """all-reduce (sum over ranks) of a small array, in place"""
        arr = np.ascontiguousarray(np.atleast_1d(arr), dtype=float)
        if self.allreduce is not None:
            self.allreduce(arr, arr.size)
        return arr

    def _spmv(self, slot, vec):
        """vec has nt entries; refresh its ghosts, multiply"""
        if self.halo_cb is not None and (self.n_halo or self.send_idx.size):
            send = np.ascontiguousarray(vec[self.send_idx])
            recv = np.zeros(self.n_halo)
            self.halo_cb(send, recv)
            vec[self.n:] = recv
        return self.mats[slot] @ vec[: self.mats[slot].shape[1]]

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
        n = self.n
        b = self.vecs[nat.VEC_B]
        self.x0 = np.zeros(self.nt); self.x0[:n] = self.vecs[nat.VEC_X0]
        self.r0 = b - self._spmv(nat.SLOT_A, self.x0)
        beta = np.sqrt(self._ar(self.r0 @ self.r0)[0])
        self.V[:] = 0
        self.V[0, :n] = self.r0 / beta
        self.generation += 1
        for c in list(self.cons.values()):          # a helper thread may be defining constraints right now
            c["done"] = 0
            c["T1"] = np.zeros(self.k_max)
            c["T2"] = np.zeros((self.k_max, self.k_max))
            c["MZ"] = np.zeros((self.k_max, self.n))
            c["t0"] = None
        self.log.append(("begin",))
        return float(beta)

    def arnoldi_begin_residual(self, j):
        """Residual of the iterate in X measured by the SpMV of Arnoldi step j (one pass over A for both)."""
        r = self._spmv(nat.SLOT_A, self.Xfull) - self.vecs[nat.VEC_B]
        self._resid = float(np.sqrt(self._ar(r @ r)[0]))
        self.log.append(("residual_rides", j))
        self.arnoldi_begin(j)

    def arnoldi_begin(self, j):
        assert getattr(self, "_begun", None) is None       # (a finished step may still be waiting to be collected)
        self.log.append(("begin_step", j))
        m = j + 1
        n = self.n
        if self.pre_kind not in (nat.PRE_NONE, nat.PRE_HOST):
            self.Zs[j, :n] = self._apply_pre(self.V[j, :n])
        w = self._spmv(nat.SLOT_A, self._Z()[j])
        Vm = self.V[:m, :n]
        orth = self.opts.get("orth", nat.ORTH_CGS2)
        h2 = None
        if orth == nat.ORTH_MGS:
            h = np.zeros(m)
            for i in range(m):
                h[i] = self._ar(Vm[i] @ w)[0]
                w = w - h[i] * Vm[i]
        else:
            h = self._ar(Vm @ w)
            w = w - Vm.T @ h
            if orth == nat.ORTH_CGS2:
                h2 = self._ar(Vm @ w)
                h = h + h2
        self._begun = (j, h, h2, w)

    def arnoldi_finish(self, j, y_iterate=None):
        bj, h, h2, w = self._begun
        assert bj == j and self._pending is None
        self._begun = None
        n = self.n
        if y_iterate is not None:                      # the iterate of the previous step rides on the last sweep
            assert self.pre_kind == nat.PRE_NONE and self.opts.get("orth", nat.ORTH_CGS2) == nat.ORTH_CGS2
            assert len(y_iterate) <= j + 1
            self.form_iterate(y_iterate)
            self.log.append(("fused_iterate", len(y_iterate)))
        if h2 is not None:
            w = w - self.V[: j + 1, :n].T @ h2
        nrm = np.sqrt(self._ar(w @ w)[0])
        if nrm != 0:
            self.V[j + 1, :n] = w / nrm
        else:
            self.V[j + 1, :n] = w
        self._pending = (j, np.concatenate([h, [nrm]]))

    def arnoldi_launch(self, j):
        self.log.append(("launch", j))
        self.arnoldi_begin(j)
        self.arnoldi_finish(j)

    def residual_launch(self):
        r = self._spmv(nat.SLOT_A, self.Xfull) - self.vecs[nat.VEC_B]
        self._resid = float(np.sqrt(self._ar(r @ r)[0]))
        self.log.append(("residual_launch",))

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
        self.Xfull = np.zeros(self.nt)
        self.Xfull[: self.n] = self.vecs[nat.VEC_X0] + self._Z()[: y.size, : self.n].T @ y
        self.X = self.Xfull[: self.n]

    def iterate_residual_launch(self, y):
        self._resid = self.iterate_residual(y)
        self.log.append(("iterate_residual_launch", len(y)))

    def iterate_residual_launch_dl(self, y, chunks=4):
        """Early download of a final candidate: here the copy is complete at once."""
        self.iterate_residual_launch(y)
        self.log.append(("early_download", len(y)))
        self._early = self.X.copy()
        return self._early

    def download_join(self):
        self.log.append(("download_join",))

    def iterate_residual_wait(self):
        r, self._resid = self._resid, None
        assert r is not None
        return r

    def iterate_residual(self, y):
        self.form_iterate(y)
        self.log.append(("iterate", len(y)))
        r = self._spmv(nat.SLOT_A, self.Xfull) - self.vecs[nat.VEC_B]
        return float(np.sqrt(self._ar(r @ r)[0]))

    # ---- pipelined loop (spis_pipe_begin / spis_step_enqueue): same records, computed at enqueue time
    def pipe_begin(self, thr, phase0):
        k = self.k_max
        beta = float(np.sqrt(self.r0 @ self.r0))
        self._pipe = {"thr2": float(thr) * float(thr), "phase": 1 if phase0 else 0, "cs": np.zeros(k), "sn": np.zeros(k),
                      "gv": np.concatenate([[beta], np.zeros(k)]), "R": np.zeros((k, k)), "tracking": True,
                      "y": {}, "rec": {}, "res": []}
        self.log.append(("pipe_begin", bool(phase0)))

    def step_enqueue(self, j, want_residual, want_iterate):
        P = self._pipe
        n, m = self.n, j + 1
        nopre = self.pre_kind == nat.PRE_NONE
        ticket = -1
        if not nopre:
            self.Zs[j, :n] = self._apply_pre(self.V[j, :n])
        if want_residual:
            r = self._spmv(nat.SLOT_A, self.Xfull) - self.vecs[nat.VEC_B]
            res2 = float(self._ar(r @ r)[0])
            if not res2 > P["thr2"]:
                P["phase"] = 1
            P["res"].append((res2, P["phase"]))
            ticket = len(P["res"]) - 1
        w = self._spmv(nat.SLOT_A, self._Z()[j])
        Vm = self.V[:m, :n]
        h1 = self._ar(Vm @ w)
        w = w - Vm.T @ h1
        h2 = self._ar(Vm @ w)
        nw2 = float(self._ar(w @ w)[0])
        s2 = float(h2 @ h2)
        n2 = nw2 - s2
        if not n2 > 0.0:
            n2 = 0.0
        col = np.concatenate([h1 + h2, [np.sqrt(n2)]])
        # Givens update and back substitution (hess_kernel)
        valid, ls, y = P["tracking"], 0.0, np.zeros(m)
        if valid:
            r = col.copy()
            for i in range(j):
                a, b = r[i], r[i + 1]
                r[i] = P["cs"][i] * a + P["sn"][i] * b
                r[i + 1] = -P["sn"][i] * a + P["cs"][i] * b
            den = np.hypot(r[j], r[j + 1])
            if den > 0:
                P["cs"][j], P["sn"][j] = r[j] / den, r[j + 1] / den
                P["R"][:j, j] = r[:j]
                P["R"][j, j] = den
                gj = P["gv"][j]
                P["gv"][j], P["gv"][j + 1] = P["cs"][j] * gj, -P["sn"][j] * gj
                ls = abs(P["gv"][j + 1])
                d = np.abs(np.diag(P["R"][:m, :m]))
                valid = bool(d.min() > 1e-14 * d.max())
                if valid:
                    y = np.linalg.solve(P["R"][:m, :m], P["gv"][:m])
            else:
                P["tracking"] = valid = False
        if not valid:
            P["phase"] = 1
        P["y"][j] = y
        P["rec"][j] = (col, y.copy(), {"valid": valid, "ls": float(ls), "norm2": n2, "nw2": nw2, "s2": s2, "phase": P["phase"]})
        w = w - Vm.T @ h2
        self.V[j + 1, :n] = w / col[m] if col[m] > 0 else 0.0
        formed = None
        if want_iterate and P["phase"] == 0:
            if nopre and j >= 1:
                self.form_iterate(P["y"][j - 1]); formed = j - 1
            elif not nopre:
                self.form_iterate(P["y"][j]); formed = j
        self.log.append(("step", j, bool(want_residual), bool(want_iterate), formed))
        return ticket

    def step_wait(self, j):
        col, y, info = self._pipe["rec"][j]
        return col.copy(), y.copy(), dict(info)

    def resid_wait(self, ticket):
        res2, phase = self._pipe["res"][ticket]
        return float(np.sqrt(res2)), res2, phase == 0

    # ---- constraints
    def constraint_define(self, c, slot, v, cc):
        self.cons[c] = {"slot": slot, "v": None if v is None else np.array(v, dtype=float), "c": cc, "done": 0,
                        "T1": np.zeros(self.k_max), "T2": np.zeros((self.k_max, self.k_max)),
                        "MZ": np.zeros((self.k_max, self.n))}

    def constraint_set_vector(self, c, v):
        self.cons[c]["v"] = None if v is None else np.array(v, dtype=float)
        self.cons[c]["t0"] = None
        self.cons[c]["done"] = 0

    def constraint_set_constant(self, c, cc):
        self.cons[c]["c"] = cc
        self.cons[c]["t0"] = None

    def constraint_terms_batch(self, cs, m):
        self.log.append(("constraint_terms_batch", tuple(int(c) for c in cs), m))
        return [self.constraint_terms(int(c), m) for c in cs]

    def constraint_terms(self, c, m):
        C = self.cons[c]
        n = self.n
        x0 = self.vecs[nat.VEC_X0]
        x0nz = not self.opts.get("x0_is_zero", 0)
        Z = self._Z()
        slot = C["slot"]
        hasM = slot >= 0
        if C.get("t0") is None:
            t0 = C["c"]
            if x0nz and (hasM or C["v"] is not None):
                parts = []
                if hasM:
                    parts.append(x0 @ self._spmv(slot, self.x0))
                if C["v"] is not None:
                    parts.append(C["v"] @ x0)
                parts = self._ar(np.array(parts))
                i = 0
                if hasM:
                    t0 += 0.5 * parts[i]; i += 1
                if C["v"] is not None:
                    t0 += parts[i]
            C["t0"] = t0
        t0 = C["t0"]
        for col in range(C["done"], m):
            t1 = 0.0
            if hasM:
                # ghosts of z_col were filled when the Arnoldi step multiplied it by A
                C["MZ"][col] = self.mats[slot] @ Z[col][: self.mats[slot].shape[1]]
                a = [Z[i, :n] @ C["MZ"][col] for i in range(col + 1)]
                if x0nz:
                    a.append(x0 @ C["MZ"][col])
                a = self._ar(np.array(a))
                C["T2"][: col + 1, col] = 0.5 * a[: col + 1]
                if x0nz:
                    t1 += a[col + 1]
                if col > 0 or C["v"] is not None:
                    bb = [C["MZ"][i] @ Z[col, :n] for i in range(col)]
                    if C["v"] is not None:
                        bb.append(C["v"] @ Z[col, :n])
                    bb = self._ar(np.array(bb))
                    C["T2"][col, :col] = 0.5 * bb[:col]
                    if C["v"] is not None:
                        t1 += bb[col]
            elif C["v"] is not None:
                t1 += self._ar(C["v"] @ Z[col, :n])[0]
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
            return self.V[j, : self.n].copy()
        if which == nat.VEC_Z:
            return self._Z()[j, : self.n].copy()
        raise KeyError(which)

    def download_Z(self, j0, j1):
        return self._Z()[j0:j1, : self.n].copy()

    def host_pre_get(self, j):
        return self.V[j, : self.n].copy()

    def host_pre_put(self, j, z):
        self.Zs[j, : self.n] = z

    def profile(self):
        return {}

    def reset_profile(self):
        pass
