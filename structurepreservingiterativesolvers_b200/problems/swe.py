"""numpy re-assembly of swe/swe.py: linearised rotating shallow water, implicit midpoint step,
RT_2 x DG_0 mixed finite elements on a periodic M x M square mesh of right triangles.

Reference weak form (swe/swe.py:75-86), unknowns z = [u; rho]:

    F1 = <(u - u0)/dt, phi> + f <(-umid_y, umid_x), phi> - c^2 <rhomid, div phi>
    F2 = <(rho - rho0)/dt + div umid, psi>,          umid = (u + u0)/2, rhomid = (rho + rho0)/2

    A = [[Mu/dt + f/2 C,  -c^2/2 D^T],      b = [[Mu/dt - f/2 C,   c^2/2 D^T],  @ z0
         [ 1/2 D,          Mr/dt    ]]           [-1/2 D,          Mr/dt     ]]

with Mu the RT mass matrix, C the Coriolis form, D[K, j] = int_K div phi_j, Mr = diag(|K|).
Energy form L = blkdiag(Mu, c^2 Mr) (:95-97), omega = [0; |K|] (:99), m0 = int rho0, e0 =
1/2 z0^T L z0 (:102-103).  Mesh, spaces and initial condition: swe/swe.py:26-40.

What is restated and what is equivalent-but-not-identical [derived, Firedrake is not installable]:
  * the discrete SPACES are the reference's (RT_2 = [P1]^2 + x P1~, 8 functions per triangle, normal
    components continuous across edges; piecewise constants), so A is the same operator;
  * the BASIS of RT_2 is ours: two normal point values per edge at the edge's Gauss points and the
    two components of the cell mean (FIAT's "point" variant places its nodes differently), and the
    numbering is ours (square-major, [u; rho] field-blocked as swe/refd.py:24-25 combines them).
    A, b, L differ from the reference's by that change of basis / permutation only;
  * structural zeros are kept (PETSc keeps them, lkdv/lkdv.py:109-110): 16 / 9 / 9 entries per
    edge / interior / density row, 12.5 per row on average, n = 12 M^2.

The mesh is uniform, so every square carries the same 12 rows up to translation: the matrices are
assembled once on a 4 x 4 mesh with the general element-by-element loop and then replicated as
stencils -- any strip of rows of a 1e8-unknown system can be produced directly in CSR form, with
global column ids, without ever forming the global matrix (row-sharded runs assemble locally).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

NU = 10          # velocity unknowns per square: H0 H1 V0 V1 D0 D1 | T0x T0y | T1x T1y
NR = 2           # density unknowns per square: T0, T1
_G = (0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0))          # 2-point Gauss abscissae on [0,1]

# Dunavant degree-5 rule on the reference triangle (barycentric points, weights summing to 1)
_a1, _b1, _w1 = 0.059715871789770, 0.470142064105115, 0.132394152788506
_a2, _b2, _w2 = 0.797426985353087, 0.101286507323456, 0.125939180544827
_QB = np.array([[1 / 3, 1 / 3, 1 / 3],
                [_a1, _b1, _b1], [_b1, _a1, _b1], [_b1, _b1, _a1],
                [_a2, _b2, _b2], [_b2, _a2, _b2], [_b2, _b2, _a2]])
_QW = np.array([0.225, _w1, _w1, _w1, _w2, _w2, _w2])


class problem:
    """Parameters of swe/swe.py:12-40.  `mlength` defaults to the reference's 40; pass 0.8*M to keep
    the mesh width of the default run (h = 0.8) when scaling up."""

    def __init__(self, N=100, M=50, degree=1, T=10, mlength=None):
        if degree != 1:
            raise NotImplementedError("only degree=1 (RT_2 x DG_0, the reference's default) is re-assembled")
        self.mlength = 40.0 if mlength is None else float(mlength)
        self.degree = degree
        self.c = 1.0
        self.f = 0.1
        self.N, self.M, self.T = N, int(M), T
        self.dt = float(T) / N
        self.h = self.mlength / self.M
        self.n = (NU + NR) * self.M * self.M

    def ic_rho(self, x, y):
        """rho0 of swe/swe.py:34-40 (u0 = 0); the bump sits in the middle of the domain."""
        c0 = 0.5 * self.mlength if self.mlength != 40.0 else 20.0
        return 10.0 * np.exp(-((x - c0) ** 2 + (y - c0) ** 2) / 20.0 ** 2)


# ------------------------------------------------------------------------------------------------
# one element
# ------------------------------------------------------------------------------------------------
def _monomials(x, y):
    """The 8 spanning functions of RT_2 at points (x, y): array (8, npts, 2)."""
    o, z = np.ones_like(x), np.zeros_like(x)
    return np.stack([np.stack(c, -1) for c in ((o, z), (x, z), (y, z), (z, o), (z, x), (z, y),
                                                (x * x, x * y), (x * y, y * y))])


def _monomial_div(x, y):
    o, z = np.ones_like(x), np.zeros_like(x)
    return np.stack([z, o, z, z, z, o, 3 * x, 3 * y])


def _element(verts, edges):
    """Nodal basis and local matrices of one triangle.

    verts: (3, 2) corner coordinates; edges: three (pa, pb, normal) with GLOBAL direction pa -> pb
    and GLOBAL unit normal, in the local order of the edge unknowns.  Local unknowns: 2 per edge
    (normal component at the edge's two Gauss points, ordered along pa -> pb), then the two
    components of the cell mean.  Returns (mass 8x8, coriolis 8x8, div 8, area).
    """
    verts = np.asarray(verts, dtype=float)
    cen = verts.mean(axis=0)
    e1, e2 = verts[1] - verts[0], verts[2] - verts[0]
    area = 0.5 * abs(e1[0] * e2[1] - e1[1] * e2[0])
    scale = np.sqrt(2.0 * area)
    q = _QB @ verts                                               # quadrature points
    qx, qy = (q[:, 0] - cen[0]) / scale, (q[:, 1] - cen[1]) / scale
    mono_q = _monomials(qx, qy)                                   # (8, nq, 2)
    vand = np.zeros((8, 8))                                       # vand[k, i] = dof_i(mono_k)
    col = 0
    for pa, pb, nrm in edges:
        pa, pb, nrm = np.asarray(pa, float), np.asarray(pb, float), np.asarray(nrm, float)
        for t in _G:
            p = pa + t * (pb - pa)
            m = _monomials(np.array([(p[0] - cen[0]) / scale]), np.array([(p[1] - cen[1]) / scale]))[:, 0, :]
            vand[:, col] = m @ nrm
            col += 1
    vand[:, 6] = np.einsum("q,kq->k", _QW, mono_q[:, :, 0])
    vand[:, 7] = np.einsum("q,kq->k", _QW, mono_q[:, :, 1])
    coef = np.linalg.inv(vand)                                    # phi_i = sum_k coef[i, k] mono_k
    phi = np.einsum("ik,kqc->iqc", coef, mono_q)                  # (8, nq, 2)
    w = _QW * area
    mass = np.einsum("q,iqc,jqc->ij", w, phi, phi)
    cor = np.einsum("q,iq,jq->ij", w, phi[:, :, 1], phi[:, :, 0]) - np.einsum("q,iq,jq->ij", w, phi[:, :, 0], phi[:, :, 1])
    div = (coef @ _monomial_div(qx, qy)) @ w / scale              # d/dx of the scaled coordinate
    return mass, cor, div, area


def _elements(h):
    """The two triangle shapes of a square [0,h]^2 split along the diagonal a-c."""
    a, b, c, d = (0.0, 0.0), (h, 0.0), (h, h), (0.0, h)
    nH, nV, nD = (0.0, 1.0), (1.0, 0.0), (1.0 / np.sqrt(2.0), -1.0 / np.sqrt(2.0))
    t0 = _element([a, b, c], [(a, b, nH), (b, c, nV), (a, c, nD)])       # bottom H, right V, diagonal
    t1 = _element([a, c, d], [(a, c, nD), (d, c, nH), (a, d, nV)])       # diagonal, top H, left V
    return t0, t1


def _local_to_global(M):
    """(M*M, 8) velocity unknown ids of T0 and T1 of every square, and (M*M,) density ids (offset 0)."""
    i, j = np.meshgrid(np.arange(M), np.arange(M), indexing="xy")          # s = j*M + i
    i, j = i.reshape(-1), j.reshape(-1)
    s = j * M + i
    s_right = j * M + (i + 1) % M
    s_up = ((j + 1) % M) * M + i
    g0 = np.stack([NU * s + 0, NU * s + 1, NU * s_right + 2, NU * s_right + 3, NU * s + 4, NU * s + 5,
                   NU * s + 6, NU * s + 7], axis=1)
    g1 = np.stack([NU * s + 4, NU * s + 5, NU * s_up + 0, NU * s_up + 1, NU * s + 2, NU * s + 3,
                   NU * s + 8, NU * s + 9], axis=1)
    return g0, g1, s


def assemble_blocks(M, h):
    """Element-by-element assembly (the general loop): Mu, C (nu x nu), D (nr x nu), areas (nr)."""
    (m0, c0, d0, a0), (m1, c1, d1, a1) = _elements(h)
    g0, g1, s = _local_to_global(M)
    nu, nr = NU * M * M, NR * M * M
    rows = np.concatenate([np.repeat(g0, 8, axis=1).reshape(-1), np.repeat(g1, 8, axis=1).reshape(-1)])
    cols = np.concatenate([np.tile(g0, (1, 8)).reshape(-1), np.tile(g1, (1, 8)).reshape(-1)])
    ncell = M * M
    Mu = sps.coo_matrix((np.concatenate([np.tile(m0.reshape(-1), ncell), np.tile(m1.reshape(-1), ncell)]), (rows, cols)), shape=(nu, nu)).tocsr()
    C = sps.coo_matrix((np.concatenate([np.tile(c0.reshape(-1), ncell), np.tile(c1.reshape(-1), ncell)]), (rows, cols)), shape=(nu, nu)).tocsr()
    drows = np.concatenate([np.repeat(NR * s, 8), np.repeat(NR * s + 1, 8)])
    dcols = np.concatenate([g0.reshape(-1), g1.reshape(-1)])
    D = sps.coo_matrix((np.concatenate([np.tile(d0, ncell), np.tile(d1, ncell)]), (drows, dcols)), shape=(nr, nu)).tocsr()
    areas = np.empty(nr)
    areas[0::2], areas[1::2] = a0, a1
    return Mu, C, D, areas


def _system_from_blocks(Mu, C, D, areas, prob):
    """A, the right-hand-side operator B (b = B z0) and L as explicit-zero-preserving CSR matrices."""
    dt, f, c2 = prob.dt, prob.f, prob.c ** 2
    Mr = sps.diags(areas).tocsr()
    A = sps.bmat([[Mu / dt + 0.5 * f * C, -0.5 * c2 * D.T], [0.5 * D, Mr / dt]], format="csr")
    B = sps.bmat([[Mu / dt - 0.5 * f * C, 0.5 * c2 * D.T], [-0.5 * D, Mr / dt]], format="csr")
    L = sps.bmat([[Mu, None], [None, c2 * Mr]], format="csr")
    for X in (A, B, L):
        X.sum_duplicates()
        X.sort_indices()
    return A, B, L


# ------------------------------------------------------------------------------------------------
# stencil replication
# ------------------------------------------------------------------------------------------------
class _Stencil:
    """The 12 rows of one square of a translation-invariant operator on the periodic mesh."""

    M0 = 4

    def __init__(self, S):
        """S: the operator assembled on the M0 x M0 mesh (explicit zeros kept)."""
        M0 = self.M0
        S = sps.csr_matrix(S)
        nu0 = NU * M0 * M0
        ref_i = ref_j = 1
        sref = ref_j * M0 + ref_i
        self.rows = []                                    # per row type: (di, dj, is_rho, local, value) arrays
        for rt in range(NU + NR):
            g = NU * sref + rt if rt < NU else nu0 + NR * sref + (rt - NU)
            lo, hi = S.indptr[g], S.indptr[g + 1]
            cols, vals = S.indices[lo:hi].astype(np.int64), S.data[lo:hi].copy()
            is_rho = cols >= nu0
            per = np.where(is_rho, NR, NU)
            loc = np.where(is_rho, cols - nu0, cols)
            sq, local = loc // per, loc % per
            di = (sq % M0 - ref_i + 1) % M0 - 1            # wrapped into {-1, 0, 1, 2}; 2 never occurs
            dj = (sq // M0 - ref_j + 1) % M0 - 1
            if (np.abs(di) > 1).any() or (np.abs(dj) > 1).any():
                raise AssertionError("stencil wider than one square")
            self.rows.append((di, dj, is_rho, local, vals))
        self.len_u = sum(len(r[0]) for r in self.rows[:NU])
        self.len_r = sum(len(r[0]) for r in self.rows[NU:])

    def replicate(self, M, j0=0, j1=None, sort=None):
        """CSR rows of the squares in strips [j0, j1) of the M x M mesh, GLOBAL column ids, local row
        order = [u rows of those squares; rho rows of those squares] (square-major)."""
        j1 = M if j1 is None else j1
        if M < 3:
            raise ValueError("the periodic mesh needs M >= 3 (a square must not neighbour itself)")
        nsq = (j1 - j0) * M
        nu = NU * M * M
        i, j = np.meshgrid(np.arange(M, dtype=np.int64), np.arange(j0, j1, dtype=np.int64), indexing="xy")
        i, j = i.reshape(-1), j.reshape(-1)
        idx_dtype = np.int32 if (NU + NR) * M * M < 2**31 - 1 else np.int64
        blocks = []
        base_u, base_r = {}, {}                                   # first unknown of the neighbour square, per offset

        def bases(d_i, d_j):
            key = (int(d_i), int(d_j))
            if key not in base_u:
                sq = ((j + d_j) % M) * M + (i + d_i) % M
                base_u[key] = (NU * sq).astype(idx_dtype)
                base_r[key] = (nu + NR * sq).astype(idx_dtype)
            return base_u[key], base_r[key]

        for first, last, width in ((0, NU, self.len_u), (NU, NU + NR, self.len_r)):
            ind = np.empty((width, nsq), dtype=idx_dtype)         # filled row by row (contiguous), transposed once
            lens, vrow = [], []
            pos = 0
            for di, dj, is_rho, local, vals in self.rows[first:last]:
                for e in range(len(di)):
                    bu, br = bases(di[e], dj[e])
                    np.add(br if is_rho[e] else bu, idx_dtype(local[e]), out=ind[pos])
                    pos += 1
                lens.append(len(di))
                vrow.append(vals)
            blocks.append((np.ascontiguousarray(ind.T).reshape(-1), np.tile(np.concatenate(vrow), nsq),
                           np.tile(np.array(lens, dtype=np.int64), nsq)))
            del ind
        indices = np.concatenate([blocks[0][0], blocks[1][0]])
        data = np.concatenate([blocks[0][1], blocks[1][1]])
        lens = np.concatenate([blocks[0][2], blocks[1][2]])
        indptr = np.zeros(lens.size + 1, dtype=np.int64)
        np.cumsum(lens, out=indptr[1:])
        if indptr[-1] < 2**31 - 1:
            indptr = indptr.astype(np.int32)
        out = sps.csr_matrix((data, indices, indptr), shape=(lens.size, (NU + NR) * M * M))
        if sort or (sort is None and M <= 400):
            out.sort_indices()
        return out


def _stencil_apply(st, M, z, j0=0, j1=None):
    """(operator @ z) for the rows of strips [j0, j1) without forming the operator: one gather of z per
    stencil entry.  z is a FULL state vector.  Used for b = B z0 on 1e8-unknown strips, where B would be
    a second 7.5 GB matrix alive only for this product."""
    j1 = M if j1 is None else j1
    nsq = (j1 - j0) * M
    nu = NU * M * M
    i, j = np.meshgrid(np.arange(M, dtype=np.int64), np.arange(j0, j1, dtype=np.int64), indexing="xy")
    i, j = i.reshape(-1), j.reshape(-1)
    out_u = np.zeros((nsq, NU))
    out_r = np.zeros((nsq, NR))
    sq_cache = {}
    for rt, (di, dj, is_rho, local, vals) in enumerate(st.rows):
        acc = np.zeros(nsq)
        for e in range(len(di)):
            if vals[e] == 0.0:
                continue
            key = (int(di[e]), int(dj[e]))
            if key not in sq_cache:
                sq_cache[key] = ((j + dj[e]) % M) * M + (i + di[e]) % M
            sq = sq_cache[key]
            acc += vals[e] * (z[nu + NR * sq + local[e]] if is_rho[e] else z[NU * sq + local[e]])
        if rt < NU:
            out_u[:, rt] = acc
        else:
            out_r[:, rt - NU] = acc
    return np.concatenate([out_u.reshape(-1), out_r.reshape(-1)])


_STENCIL_CACHE = {}


def _stencils(prob):
    key = (prob.h, prob.dt, prob.f, prob.c)
    if key not in _STENCIL_CACHE:
        blocks = assemble_blocks(_Stencil.M0, prob.h)
        A, B, L = _system_from_blocks(*blocks, prob)
        _STENCIL_CACHE[key] = (_Stencil(A), _Stencil(B), _Stencil(L), blocks[3][:2].copy())
    return _STENCIL_CACHE[key]


def strip_ids(M, j0, j1):
    """Global ids of the unknowns of strips [j0, j1): the local row order of `linforms(rows=...)`."""
    u = np.arange(NU * j0 * M, NU * j1 * M, dtype=np.int64)
    r = NU * M * M + np.arange(NR * j0 * M, NR * j1 * M, dtype=np.int64)
    return np.concatenate([u, r])


def initial_state(prob, j0=0, j1=None):
    """z0 restricted to strips [j0, j1) (local order of strip_ids): u0 = 0, rho0 = the bump evaluated
    at the cell centroids (DG0 interpolation, swe/swe.py:61-62)."""
    M, h = prob.M, prob.h
    j1 = M if j1 is None else j1
    i, j = np.meshgrid(np.arange(M), np.arange(j0, j1), indexing="xy")
    x0, y0 = i.reshape(-1) * h, j.reshape(-1) * h
    rho = np.empty(NR * x0.size)
    rho[0::2] = prob.ic_rho(x0 + 2.0 * h / 3.0, y0 + h / 3.0)          # centroid of T0 = (a + b + c)/3
    rho[1::2] = prob.ic_rho(x0 + h / 3.0, y0 + 2.0 * h / 3.0)          # centroid of T1 = (a + c + d)/3
    return np.concatenate([np.zeros(NU * x0.size), rho])


def linforms(N=100, M=50, degree=1, T=10, zinit=None, mlength=None, rows=None, method="stencil", sort=None):
    """Same keys as swe/swe.py:46-120 ('A','b','omega','L','m0','e0','z0','T').

    rows=(j0, j1): only the rows of mesh strips [j0, j1) (for one rank of a row-sharded run): A and L
    then hold those rows with GLOBAL column ids, b / omega / z0 the matching local pieces, and m0 / e0
    the GLOBAL invariants.  method='loop' uses the element-by-element assembly for everything
    (small meshes; the cross-check of the stencil path).
    """
    prob = problem(N=N, M=M, degree=degree, T=T, mlength=mlength)
    M = prob.M
    j0, j1 = (0, M) if rows is None else (int(rows[0]), int(rows[1]))
    z0_full = initial_state(prob) if zinit is None else np.asarray(zinit, dtype=float)
    if z0_full.size != prob.n:
        raise ValueError("zinit must be a full state vector")
    ids = strip_ids(M, j0, j1)
    if method == "loop":
        A, B, L = _system_from_blocks(*assemble_blocks(M, prob.h), prob)
        if rows is not None:
            A, B, L = A[ids], B[ids], L[ids]
        areas2 = None
    else:
        sA, sB, sL, areas2 = _stencils(prob)
        A, L = (s.replicate(M, j0, j1, sort=sort) for s in (sA, sL))
        B = None
    b = B @ z0_full if B is not None else _stencil_apply(sB, M, z0_full, j0, j1)
    area0, area1 = (0.5 * prob.h ** 2, 0.5 * prob.h ** 2)
    omega_full_rho = np.empty(NR * M * M)
    omega_full_rho[0::2], omega_full_rho[1::2] = area0, area1
    nloc_sq = (j1 - j0) * M
    omega = np.concatenate([np.zeros(NU * nloc_sq), omega_full_rho[NR * j0 * M: NR * j1 * M]])
    rho_full = z0_full[NU * M * M:]
    m0 = float(omega_full_rho @ rho_full)
    # e0 = 1/2 z0^T L z0 over the whole mesh; u0 = 0 for the default initial state
    if zinit is None:
        e0 = 0.5 * prob.c ** 2 * float(rho_full @ (omega_full_rho * rho_full))
    else:
        Lfull = L if rows is None else _stencils(prob)[2].replicate(M, sort=False)
        e0 = 0.5 * float(z0_full @ (Lfull @ z0_full))
    out = {"A": A, "b": b, "omega": omega, "L": L, "m0": m0, "e0": e0,
           "z0": z0_full[ids] if rows is not None else z0_full, "T": T}
    return out, prob


def compute_invariants(dic, vec):
    """mass = int rho, energy = 1/2 int |u|^2 + c^2 rho^2 (swe/swe.py:123-137), full systems only."""
    vec = np.asarray(vec, dtype=float)
    return {"mass": float(dic["omega"] @ vec), "energy": 0.5 * float(vec @ (dic["L"] @ vec))}


def benchmark_size(n_target):
    """Smallest M with 12 M^2 >= n_target (M = 913 -> n = 10 002 828)."""
    M = int(np.ceil(np.sqrt(n_target / float(NU + NR))))
    return max(M, 3)
