"""FEMatrix / FEVector: block overlays on one CSC matrix / one dense vector.

Mirrors /root/reference/src/fematrix.jl:75-79, :226-290 (block offsets: rows follow
the first FESpace list, columns the second) and /root/reference/src/fevector.jl:85-89,
:234-267.  The CSC image uses Julia's conventions so that it can be handed to / taken
from ``A.entries.cscmatrix`` unchanged: ``colptr`` int64 ``n+1``, ``rowval`` int64
``nnz`` ascending within a column, ``nzval`` float64, all **1-based**.
"""
from __future__ import annotations

import numpy as np


class ExtendableSparseMatrixCSC:
    """The ``entries`` object of an FEMatrix ([EXT E2] ExtendableSparse): only the
    flushed CSC part exists here; the linked-list overflow of the reference is
    replaced by the GPU pattern builder."""

    def __init__(self, m: int, n: int):
        self.m, self.n = int(m), int(n)
        self.colptr = np.ones(self.n + 1, dtype=np.int64)
        self.rowval = np.zeros(0, dtype=np.int64)
        self.nzval = np.zeros(0, dtype=np.float64)

    @property
    def cscmatrix(self):
        return self

    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def nnz(self):
        return int(self.colptr[-1] - 1)

    def set_csc(self, colptr, rowval, nzval):
        self.colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        self.rowval = np.ascontiguousarray(rowval, dtype=np.int64)
        self.nzval = np.ascontiguousarray(nzval, dtype=np.float64)

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))

    def find(self, i: int, j: int) -> int:
        """0-based position of 1-based entry (i,j) in nzval, or -1."""
        lo, hi = self.colptr[j - 1] - 1, self.colptr[j] - 1
        p = lo + np.searchsorted(self.rowval[lo:hi], i)
        return int(p) if p < hi and self.rowval[p] == i else -1

    def __getitem__(self, ij):
        p = self.find(*ij)
        return 0.0 if p < 0 else float(self.nzval[p])

    def __setitem__(self, ij, v):
        p = self.find(*ij)
        if p < 0:
            raise KeyError(f"entry {ij} not in pattern (insertions go through the pattern builder)")
        self.nzval[p] = v


class FEMatrixBlock:
    def __init__(self, name, fesx, fesy, offset, offsety, entries):
        self.name, self.FESX, self.FESY = name, fesx, fesy
        self.offset, self.offsetY = offset, offsety
        self.last_index, self.last_indexY = offset + fesx.ndofs, offsety + fesy.ndofs
        self.entries = entries


class FEMatrix:
    """``FEMatrix(FESX[, FESY])`` (fematrix.jl:226-290)."""

    def __init__(self, fesx, fesy=None):
        X = list(fesx) if isinstance(fesx, (list, tuple)) else [fesx]
        Y = X if fesy is None else (list(fesy) if isinstance(fesy, (list, tuple)) else [fesy])
        self.FESX, self.FESY = X, Y
        self.entries = ExtendableSparseMatrixCSC(sum(f.ndofs for f in X), sum(f.ndofs for f in Y))
        self.FEMatrixBlocks = []
        off = 0
        for fx in X:
            offy = 0
            for fy in Y:
                self.FEMatrixBlocks.append(FEMatrixBlock(f"{fx.name} x {fy.name}", fx, fy, off, offy, self.entries))
                offy += fy.ndofs
            off += fx.ndofs
        self.nbrow, self.nbcol = len(X), len(Y)

    def __getitem__(self, ij):
        if isinstance(ij, tuple):
            i, j = ij
            return self.FEMatrixBlocks[(i - 1) * self.nbcol + (j - 1)]
        return self.FEMatrixBlocks[ij - 1]

    @property
    def shape(self):
        return self.entries.shape


class FEVectorBlock:
    def __init__(self, name, fes, offset, entries):
        self.name, self.FES, self.offset, self.last_index, self.entries = name, fes, offset, offset + fes.ndofs, entries

    def view(self):
        return self.entries[self.offset : self.last_index]

    def __getitem__(self, i):
        return self.entries[self.offset + i - 1]

    def __setitem__(self, i, v):
        self.entries[self.offset + i - 1] = v


class FEVector:
    """``FEVector(FES)`` (fevector.jl:234-267): one dense Float64 vector with block views."""

    def __init__(self, fes):
        F = list(fes) if isinstance(fes, (list, tuple)) else [fes]
        self.entries = np.zeros(sum(f.ndofs for f in F), dtype=np.float64)
        self.FEVectorBlocks = []
        off = 0
        for f in F:
            self.FEVectorBlocks.append(FEVectorBlock(f.name, f, off, self.entries))
            off += f.ndofs

    def __getitem__(self, i):
        return self.FEVectorBlocks[i - 1]

    def __len__(self):
        return len(self.FEVectorBlocks)
