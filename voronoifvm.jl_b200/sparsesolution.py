"""Host mirror of the reference's sparse unknown storage (`unknown_storage = :sparse`, src/vfvm_system.jl:813-825,
src/vfvm_sparsesolution.jl:10-172): only the degrees of freedom of species that are defined at a node are stored, in the CSC
layout of the `node_dof` matrix (column = node, rows = the species enabled there, ascending).

The device twin always works on the dense n x N layout (dofs of undefined species are identity rows and stay zero); this class is
the host-side container with the reference's accessors -- `dof`, `dofs`, `a[i, K]` (NaN for an undefined dof, as
src/vfvm_sparsesolution.jl:156-166), assignment that ignores undefined dofs (:139-149), `+`/`-`, `copy`/`similar` -- and the two
conversions the drop-in needs: `dense()` (upload) and `from_dense()` (download).  Species and node arguments are 0-based here, like
every array index in this package.
"""
from __future__ import annotations

import numpy as np


class SparseSolutionArray:
    def __init__(self, node_dof: np.ndarray, values=None, history=None):
        mask = np.asfortranarray(node_dof, dtype=bool)  # (n, N)
        self.shape = mask.shape
        self._mask = mask
        cnt = mask.sum(axis=0)
        self.colptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)  # node_dof.colptr
        self.rowval = np.nonzero(mask.T)[1].astype(np.int32)  # species of every stored dof, node after node
        ndof = int(self.colptr[-1])
        self.nzval = np.zeros(ndof) if values is None else np.array(values, dtype=np.float64).reshape(ndof)
        self.history = history

    # -- src/vfvm_sparsesolution.jl:104-113
    def dof(self, ispec: int, inode: int) -> int:
        """position of dof (ispec, inode) in `dofs()`, -1 if the species is not defined at the node"""
        lo, hi = self.colptr[inode], self.colptr[inode + 1]
        k = lo + np.searchsorted(self.rowval[lo:hi], ispec)
        return int(k) if k < hi and self.rowval[k] == ispec else -1

    def dofs(self) -> np.ndarray:
        return self.nzval

    def __getitem__(self, idx):
        i, K = idx
        if isinstance(i, (int, np.integer)) and isinstance(K, (int, np.integer)):
            k = self.dof(int(i), int(K))
            return self.nzval[k] if k >= 0 else float("nan")
        return self.dense(fill=np.nan)[idx]

    def __setitem__(self, idx, val):
        i, K = idx
        if isinstance(i, (int, np.integer)) and isinstance(K, (int, np.integer)):
            k = self.dof(int(i), int(K))
            if k >= 0:  # undefined dofs are ignored, so that broadcasts work (:139-149)
                self.nzval[k] = val
            return
        d = self.dense()
        d[idx] = val
        self.nzval[:] = d.T[self._mask.T]

    def dense(self, fill=0.0) -> np.ndarray:
        """(n, N) Fortran-ordered array, `fill` at undefined dofs: what the device twin uploads"""
        d = np.full(self.shape, fill, order="F")
        d.T[self._mask.T] = self.nzval
        return d

    @classmethod
    def from_dense(cls, node_dof, dense, history=None):
        a = cls(node_dof, history=history)
        a.nzval[:] = np.asarray(dense).T[a._mask.T]
        return a

    def copy(self):
        return SparseSolutionArray(self._mask, self.nzval.copy(), self.history)

    def similar(self):
        return SparseSolutionArray(self._mask)

    def __add__(self, other):
        return SparseSolutionArray(self._mask, self.nzval + other.nzval)

    def __sub__(self, other):
        return SparseSolutionArray(self._mask, self.nzval - other.nzval)

    def __len__(self):
        return self.nzval.size

    def __array__(self, dtype=None, copy=None):
        return self.dense(fill=np.nan) if dtype is None else self.dense(fill=np.nan).astype(dtype)
