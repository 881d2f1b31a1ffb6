"""TEST INFRASTRUCTURE — oracle comparison of the multi-GPU device path (one rank per GPU, NCCL exchange).

A small slab problem is pushed through exactly the code the benchmark times (symbolic phase with halo rows, exchange plan,
femb_assemble_poisson with the fused / sequential interface exchange); every rank then compares its OWNED columns of the
matrix and its owned entries of the vector with the single-domain numpy oracle (orc.example200_assemble on the global
mesh).  Used by tests/test_multigpu.py (torchrun, >= 2 GPUs) and by bench.py at N > 1 (`checks.mg_parity_max_rel`) — as
the checker only, outside every timed region.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _owned_global(part, colptr, rowval, nz, n_global, l2g_dense):
    """the rank's owned columns as a global-numbered sparse matrix (n_global x n_global)"""
    nloc = part.l2g.shape[0]
    cols = np.repeat(np.arange(nloc, dtype=np.int64), np.diff(colptr))
    keep = part.owner[cols] == part.rank
    return sp.csc_matrix((nz[keep], (l2g_dense[rowval[keep] - 1], l2g_dense[cols[keep]])), shape=(n_global, n_global))


def _compare(part, colptr, rowval, nz, b, Aref, bref, l2g_dense):
    n_global = Aref.shape[0]
    mine = _owned_global(part, colptr, rowval, nz, n_global, l2g_dense)
    owned = np.nonzero(part.owner == part.rank)[0]
    mask = np.zeros(n_global)
    mask[l2g_dense[owned]] = 1.0
    want = Aref @ sp.diags(mask)
    D = (mine - want).tocoo()
    scale = np.abs(Aref.data).max()
    rel_a = float(np.abs(D.data).max() / scale) if D.nnz else 0.0
    rel_b = float(np.abs(b[owned] - bref[l2g_dense[owned]]).max() / max(np.abs(bref).max(), 1e-300))
    return rel_a, rel_b


def c2_reference(n, nranks, rows):
    """single-domain oracle for the 2-D P1 slab problem (global mesh = all slabs stacked in y)"""
    import femb_b200 as fb
    import femb_oracle as orc

    X = np.linspace(0.0, 1.0, n + 1)
    Y = np.arange(nranks * rows + 1) / float(n)
    grid = fb.simplexgrid(X, Y)
    el = orc.Element("H1Pk", 1, 1)
    cn = grid["CellNodes"]
    colptr, rowval, nzval, b = orc.example200_assemble(el, "Triangle2D", grid["Coordinates"], cn, grid["CellVolumes"], cn, grid.nnodes, 1.0,
                                                       lambda x: x[..., 0] - x[..., 1])
    return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(grid.nnodes,) * 2), b


def check_c2(ctx, part, n, rows):
    """after one assembly (+ exchange) of the slab problem on this rank's ctx: (max rel error of owned columns, of owned b)"""
    nloc = part.l2g.shape[0]
    colptr, rowval = ctx.pattern_get(nloc)
    nz, b = ctx.matrix_get(), ctx.vector_get(nloc)
    Aref, bref = c2_reference(n, part.nranks, rows)
    return _compare(part, colptr, rowval, nz, b, Aref, bref, part.l2g)


def c3_reference(m, nranks, layers, mu, rhs):
    """single-domain oracle for the 3-D P2 z-slab problem + the map from the partition's sparse global edge ids to the
    dense dof numbering of the global FESpace"""
    import femb_b200 as fb
    import femb_oracle as orc

    X = np.linspace(0.0, 1.0, m + 1)
    Z = np.arange(0, nranks * layers + 1) / float(m)
    grid = fb.simplexgrid(X, X, Z)
    fes = fb.FESpace(fb.H1P2(1, 3), grid)
    el = orc.Element("H1P2", 1, 2)
    colptr, rowval, nzval, b = orc.example200_assemble(el, "Tetrahedron3D", grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], fes["CellDofs"],
                                                       fes.ndofs, mu, rhs, None, None, 0.0)
    NN = grid.nnodes
    en = grid["EdgeNodes"].astype(np.int64) - 1
    ekey = NN + np.minimum(en[:, 0], en[:, 1]) * NN + np.maximum(en[:, 0], en[:, 1])
    order = np.argsort(ekey)
    return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(fes.ndofs,) * 2), b, NN, ekey[order], order


def check_c3(ctx, part, m, layers, mu, rhs):
    nloc = part.l2g.shape[0]
    colptr, rowval = ctx.pattern_get(nloc)
    nz, b = ctx.matrix_get(), ctx.vector_get(nloc)
    Aref, bref, NN, ekeys, eorder = c3_reference(m, part.nranks, layers, mu, rhs)
    dense = part.l2g.copy()
    is_edge = part.l2g >= NN
    pos = np.searchsorted(ekeys, part.l2g[is_edge])
    assert np.array_equal(ekeys[pos], part.l2g[is_edge]), "an edge dof of the partition is not an edge of the global mesh"
    dense[is_edge] = NN + eorder[pos]
    return _compare(part, colptr, rowval, nz, b, Aref, bref, dense)
