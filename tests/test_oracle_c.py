"""The C restatement of the reference loop (oracle/femb_oracle_c.c, the CPU baseline of bench.py) against the
numpy oracle: same pattern, same values — bit-identical in serial mode (same summation order)."""
import numpy as np
import pytest

import femb_oracle as orc
import oracle_c
import femb_b200 as fb


@pytest.mark.parametrize("dim,order,n,jit", [(2, 1, 24, False), (2, 1, 17, True), (2, 2, 12, True), (3, 1, 5, True), (3, 2, 3, True)])
def test_c_oracle_matches_numpy_oracle(dim, order, n, jit):
    X = np.linspace(0, 1, n + 1)
    grid = fb.simplexgrid(X, X) if dim == 2 else fb.simplexgrid(X, X, X)
    if jit:
        grid = fb.jitter(grid, 0.15)
    g = grid.geometry
    el = orc.Element("H1Pk", 1, order) if dim == 2 else orc.Element("H1P%d" % order, 1)
    ft = fb.H1Pk(1, 2, order) if dim == 2 else (fb.H1P1(1) if order == 1 else fb.H1P2(1, 3))
    fes = fb.FESpace(ft, grid)
    cd = fes["CellDofs"]
    coef = np.array([0.0, 1.0, -1.0] if dim == 2 else [0.25, 1.0, -1.0, 0.5])
    rhs = lambda x: coef[0] + x @ coef[1:]  # noqa: E731
    colptr, rowval, nzval, b = orc.example200_assemble(el, g, grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], cd, fes.ndofs, 1.3, rhs)
    xref, w = orc.quadrature(g, 2 * (order - 1))
    vals, der = el.tabulate(g, xref)
    for nthreads in (1, 4):
        nz2, b2 = np.zeros_like(nzval), np.zeros_like(b)
        miss = oracle_c.example200(grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], cd, w, xref, vals[:, :, 0], der[:, :, 0, :], 1.3, coef, 1e-15,
                                   colptr, rowval, nz2, b2, nthreads)
        assert miss == 0
        scale = np.abs(nzval).max()
        tol = 0.0 if nthreads == 1 and order == 1 else 1e-13
        assert np.abs(nz2 - nzval).max() <= tol * scale
        assert np.abs(b2 - b).max() <= 1e-13 * np.abs(b).max()
        if order == 1 or jit:
            assert np.all(nz2 != 0.0)  # every slot of the filtered pattern received a contribution
