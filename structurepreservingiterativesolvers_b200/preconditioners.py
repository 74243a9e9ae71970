"""Device-native preconditioner descriptions (Jacobi, block-Jacobi).

The reference accepts ``pre=None``, an object with ``.solve`` (SuperLU ILU, swe/TimedSolve.py:23)
or anything supporting ``pre @ vec`` (solvers.py:149-161).  north_star item (c) asks for a
Jacobi / block-diagonal preconditioner on the device; these two classes describe one.  They are
recognised by DeviceSession and applied by `jacobi_kernel` / `blockdiag_kernel`.

Both also implement ``@`` on host vectors so that the very same object can be handed to the
REFERENCE solver (or the oracle) when generating parity data; the GPU solvers never call it.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps


class JacobiPreconditioner:
    """z = D^{-1} q with D = diag(A)."""

    def __init__(self, A=None, diag=None):
        if diag is None:
            diag = sps.csr_matrix(A).diagonal()
        diag = np.asarray(diag, dtype=np.float64).reshape(-1)
        if not np.all(diag != 0):
            raise ValueError("Jacobi preconditioner needs a zero-free diagonal")
        self.dinv = 1.0 / diag
        self.shape = (diag.size, diag.size)
        self.dtype = np.dtype(np.float64)

    def __matmul__(self, vec):
        return self.dinv * np.asarray(vec)

    def tocsr(self):
        return sps.diags(self.dinv).tocsr()


class BlockJacobiPreconditioner:
    """z = blockdiag(B_i^{-1}) q for dense bs x bs diagonal blocks of A.

    layout='contiguous': block i holds unknowns i*bs .. i*bs+bs-1.
    layout='field'     : unknowns are field-blocked [u; v; w] (lkdv/refd.py:17) with bs fields of
                         nblk entries each; block i couples unknowns {i + f*nblk}: the 3 x 3
                         node-block Jacobi of SURVEY section 7.2 H-D.
    """

    def __init__(self, A, bs: int, layout: str = "contiguous"):
        A = sps.csr_matrix(A)
        n = A.shape[0]
        if not 1 <= bs <= 8:
            raise ValueError("block size must be in 1..8")
        if n % bs:
            raise ValueError("matrix size must be a multiple of the block size")
        nblk = n // bs
        blocks = np.zeros((nblk, bs, bs))
        if layout == "contiguous":
            self.stride_block, self.stride_field = bs, 1
            for r in range(bs):
                for c in range(bs):
                    d = A.diagonal(c - r)                   # A[p, p + c - r]
                    start = r if c >= r else c              # index into d of p = i*bs + r
                    # d[p'] with p' = p when c>=r (row index), p' = p + (c-r) when c<r (col index)
                    blocks[:, r, c] = d[start::bs][:nblk]
        elif layout == "field":
            self.stride_block, self.stride_field = 1, nblk
            for r in range(bs):
                for c in range(bs):
                    blocks[:, r, c] = A[r * nblk:(r + 1) * nblk, c * nblk:(c + 1) * nblk].diagonal()
        else:
            raise ValueError("layout must be 'contiguous' or 'field'")
        self.bs, self.nblk, self.layout = bs, nblk, layout
        self.inv_blocks = np.ascontiguousarray(np.linalg.inv(blocks))
        self.shape = (n, n)
        self.dtype = np.dtype(np.float64)

    def _gather(self, vec):
        vec = np.asarray(vec, dtype=np.float64)
        if self.layout == "contiguous":
            return vec.reshape(self.nblk, self.bs)
        return vec.reshape(self.bs, self.nblk).T

    def __matmul__(self, vec):
        x = self._gather(vec)
        z = np.einsum("irc,ic->ir", self.inv_blocks, x)
        if self.layout == "contiguous":
            return z.reshape(-1)
        return np.ascontiguousarray(z.T).reshape(-1)

    def tocsr(self):
        """The same operator as an explicit sparse matrix (for cross-checks)."""
        idx = (np.arange(self.nblk)[:, None] * self.stride_block
               + np.arange(self.bs)[None, :] * self.stride_field)        # (nblk, bs)
        rows = np.repeat(idx[:, :, None], self.bs, axis=2).reshape(-1)
        cols = np.repeat(idx[:, None, :], self.bs, axis=1).reshape(-1)
        return sps.csr_matrix((self.inv_blocks.reshape(-1), (rows, cols)), shape=self.shape)
