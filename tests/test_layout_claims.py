"""CPU checks of the column-layout facts the fused owner-computes kernels rely on (femb_ns.cu k_th_divp_gather, femb_hdiv.cu
k_hdiv_mass_gather<4, true>), with the algorithms restated in numpy on the oracle's CSC pattern and compared with the oracle's
own local matrices.  The device kernels re-check the same facts once per pattern (k_th_layout_check, k_hdiv_xsrc)."""
import numpy as np

import femb_oracle as orc
import femb_b200 as fb


def _role_perm2():
    """femb_p2.cu p2_role_perm2: perm[k][role] = local dof (self / other vertices / incident edges / opposite edge; edges: endpoints, other vertex, self, ...)."""
    def edge_of2(a, b):
        return (0 if b == 1 else 2) if a == 0 else 1

    def E(a, b):
        return 3 + (edge_of2(a, b) if a < b else edge_of2(b, a))

    perm = np.zeros((6, 6), int)
    for k in range(3):
        o1, o2 = (1 if k == 0 else 0), (1 if k == 2 else 2)
        perm[k] = [k, o1, o2, E(k, o1), E(k, o2), E(o1, o2)]
    for e, (i, j) in enumerate([(0, 1), (1, 2), (0, 2)]):
        k = 3 - i - j
        perm[3 + e] = [i, j, k, E(i, j), E(i, k), E(j, k)]
    return perm


def _th_value(i, a, D):
    """-|T| int d_c phi_i lambda_a without the factor (femb_ns.cu th_value), D = |T| d_c lambda_{0,1,2}."""
    if i < 3:
        return D[a] / 3.0 if i == a else 0.0
    ea, eb = (0, 1) if i == 3 else ((1, 2) if i == 4 else (0, 2))
    ma = 1 / 6 if ea == a else 1 / 12
    mb = 1 / 6 if eb == a else 1 / 12
    return 4.0 * (ma * D[eb] + mb * D[ea])


def test_taylor_hood_pressure_entries_from_the_velocity_columns():
    """Example210:293-301,376-380 assembled by the column owners: the pressure rows of a velocity column sit behind its [ux | uy] rows at
    the positions of the column's vertex rows, and the pressure column of node v has the rows of the velocity column of vertex dof v."""
    X = np.linspace(0, 1, 6)
    grid = fb.jitter(fb.simplexgrid(X, X))
    FESu, FESp = fb.FESpace(fb.H1Pk(2, 2, 2), grid), fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    cdu, cdp = FESu["CellDofs"].astype(np.int64), FESp["CellDofs"].astype(np.int64)
    ndu, ndp = FESu.ndofs, FESp.ndofs
    n, nnodes = ndu // 2, grid.nnodes
    assert ndp == nnodes
    sol = np.zeros(ndu + ndp)
    Aloc, Bloc, _ = orc.example210_local(grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], cdu, sol, 1e-2, 0.0, True, False)
    nall = ndu + ndp
    I, J, V = orc.example210_triplets(np.ones_like(Aloc), Bloc, cdu, cdp, ndu)  # pattern of the three blocks
    I2, J2, V2 = orc.example210_triplets(np.zeros_like(Aloc), Bloc, cdu, cdp, ndu)
    colptr, rowval, want = orc.flush_csc(nall, nall, np.concatenate([I, I2]), np.concatenate([J, J2]), np.concatenate([0 * V, V2]))
    cp0 = colptr - 1
    perm = _role_perm2()
    coords, cn, vol = grid["Coordinates"], grid["CellNodes"].astype(np.int64) - 1, grid["CellVolumes"]
    out = np.zeros_like(want)
    for j in range(n):
        b0, e0 = cp0[j], cp0[j + 1]
        L = int(np.searchsorted(rowval[b0:e0], n, side="right"))  # run of the component-diagonal block
        npr = [int(cp0[j + c * n + 1] - cp0[j + c * n]) - 2 * L for c in range(2)]
        assert npr[0] == npr[1] >= 0
        run = rowval[b0:b0 + L]
        assert np.all(run[:npr[0]] <= nnodes) and np.all(run[npr[0]:] > nnodes)  # vertex rows first, as many as pressure rows
        assert np.array_equal(rowval[b0 + 2 * L:e0] - ndu, run[:npr[0]])  # ... and they are the same nodes in the same order
        P, Q = np.zeros((2, max(npr[0], 1))), np.zeros((2, L))
        isv = j < nnodes
        if isv:
            pb = cp0[ndu + j]
            assert np.array_equal(rowval[pb:pb + L], run) and np.array_equal(rowval[pb + L:pb + 2 * L], run + n)
        for cell, k in zip(*np.nonzero(cdu[:, :6] == j + 1)):
            x = coords[cn[cell]]
            a00, a01, a10, a11 = x[1, 0] - x[0, 0], x[2, 0] - x[0, 0], x[1, 1] - x[0, 1], x[2, 1] - x[0, 1]
            sc = vol[cell] / (a00 * a11 - a01 * a10)
            D = np.zeros((2, 3))
            D[0, 1], D[1, 1], D[0, 2], D[1, 2] = a11 * sc, -a01 * sc, -a10 * sc, a00 * sc
            D[:, 0] = -(D[:, 1] + D[:, 2])
            for role in range(6):
                il = perm[k][role]
                pos = int(np.searchsorted(run, cdu[cell, il]))  # the map's position byte of local dof il
                for c in range(2):
                    if il < 3:
                        P[c, pos] -= _th_value(k, il, D[c])
                    if isv:
                        Q[c, pos] -= _th_value(il, k, D[c])
        for c in range(2):
            ub = cp0[j + c * n]
            out[ub + 2 * L: ub + 2 * L + npr[c]] = P[c, :npr[c]]
        if isv:
            out[pb:pb + L], out[pb + L:pb + 2 * L] = Q[0], Q[1]
    assert np.abs(out - want).max() <= 1e-14 * np.abs(want).max()


def test_rt0_mixed_columns_written_in_one_piece():
    """HDIVRT0 + L2P0 on tetrahedra: a face column is [mass rows | rows of its one or two cells], the cell column holds the four (div v, q)
    entries; the transposed entries of a face column are copies of them in ascending cell order (femb_hdiv.cu k_hdiv_xsrc)."""
    X = np.linspace(0, 1, 4)
    grid = fb.jitter(fb.simplexgrid(X, X, X), 0.1)
    FESv, FESq = fb.FESpace(fb.HDIVRT0(3), grid), fb.FESpace(fb.L2P0(1), grid)
    g = grid.geometry
    el, elq = orc.Element("HDIVRT0", 3), orc.Element("L2P0", 1)
    xref, w = orc.quadrature(g, 2)
    args = (g, xref, grid["Coordinates"], grid["CellNodes"], grid["CellFaceSigns"], None, None)
    V, D, Q = orc.update_basis(el, "Identity", *args), orc.update_basis(el, "Divergence", *args), orc.update_basis(elq, "Identity", *args)
    vol = grid["CellVolumes"]
    M, B = orc.bilinear_local(V, V, w, vol), orc.bilinear_local(D, Q, w, vol)
    cdv, cdq = FESv["CellDofs"], FESq["CellDofs"]
    I1, J1, V1 = orc.block_triplets(M, cdv, cdv)
    I2, J2, V2 = orc.block_triplets(B, cdv, cdq, 0, FESv.ndofs)
    nall = FESv.ndofs + FESq.ndofs
    colptr, rowval, nzval = orc.flush_csc(nall, nall, np.concatenate([I1, I2, J2]), np.concatenate([J1, J2, I2]), np.concatenate([V1, V2, V2]))
    cp0 = colptr - 1
    nf = FESv.ndofs
    fc = grid["FaceCells"]
    for f in range(nf):
        rows = rowval[cp0[f]:cp0[f + 1]]
        nmass = int(np.searchsorted(rows, nf, side="right"))
        cells = rows[nmass:] - nf
        assert 1 <= cells.size <= 2 and np.array_equal(cells, np.sort(fc[f][fc[f] > 0]))  # the (1,0) run follows the mass run
        for r, c in enumerate(cells):
            col = rowval[cp0[nf + c - 1]:cp0[nf + c]]
            src = cp0[nf + c - 1] + int(np.searchsorted(col, f + 1))
            assert rowval[src] == f + 1 and nzval[src] == nzval[cp0[f] + nmass + r]  # the copy the gather makes
    for c in range(grid.ncells):
        assert cp0[nf + c + 1] - cp0[nf + c] == 4  # one 32-byte sector per cell column
