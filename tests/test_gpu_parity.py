"""GPU parity: libfemb (through the C ABI, via the host mirror of the reference API) against the CPU
oracle on the same seeded inputs.  Bars: CellDofs / CSC pattern bit-exact; values within 1e-12
relative, norm-wise (|delta| <= 1e-12 * max|local contribution|, SURVEY.md §7 hard part 2)."""
import numpy as np
import pytest

import femb_oracle as orc
import femb_b200 as fb
from femb_b200 import _lib

pytestmark = pytest.mark.gpu

RTOL = 1.0e-12


def rhs_affine():
    return fb.AffineFunction([0.0], [[1.0, -1.0]])  # Example200:38  x -> x[1] - x[2]


def _grid2d(n, jit=False):
    X = np.linspace(0, 1, n + 1)
    g = fb.simplexgrid(X, X)
    return fb.jitter(g) if jit else g


def _grid3d(n, jit=False):
    X = np.linspace(0, 1, n + 1)
    g = fb.simplexgrid(X, X, X)
    return fb.jitter(g, 0.15) if jit else g


def _oracle_el(ft):
    name = ft.name.split("{")[0]
    return orc.Element(name, ft.ncomponents if ft.family != "Hdiv" else ft.ncomponents, ft.order)


def _signs(grid, ft):
    h = ft.handler(grid.geometry)
    fs = grid["CellFaceSigns"] if h in (1, 3, 4, 5) else None
    es = grid["CellEdgeSigns"] if h in (2, 6) else None
    ori = grid["CellFaceOrientations"] if h == 5 else None
    return fs, es, ori


# ---------------------------------------------------------------------------------------------
# stages 1-2: batched update_basis!
# ---------------------------------------------------------------------------------------------
EVAL_CASES = [
    (lambda: fb.H1P1(1), 2, "Gradient", 0), (lambda: fb.H1Pk(1, 2, 2), 2, "Gradient", 2), (lambda: fb.H1Pk(2, 2, 2), 2, "Gradient", 5),
    (lambda: fb.H1Pk(2, 2, 2), 2, "Identity", 5), (lambda: fb.H1Pk(2, 2, 2), 2, "Divergence", 5), (lambda: fb.H1P3(1, 2), 2, "Gradient", 4),
    (lambda: fb.H1P3(1, 2), 2, "Identity", 4), (lambda: fb.H1Pk(1, 2, 4), 2, "Identity", 6), (lambda: fb.H1P1(1), 3, "Gradient", 0),
    (lambda: fb.H1P2(1, 3), 3, "Gradient", 2), (lambda: fb.H1P2(3, 3), 3, "Divergence", 2), (lambda: fb.H1P3(1, 3), 3, "Gradient", 2),
    (lambda: fb.HDIVRT0(2), 2, "Identity", 2), (lambda: fb.HDIVRT0(2), 2, "Divergence", 2), (lambda: fb.HDIVBDM1(2), 2, "Identity", 2),
    (lambda: fb.HDIVRT0(3), 3, "Identity", 2), (lambda: fb.HDIVRT0(3), 3, "Divergence", 2), (lambda: fb.HDIVBDM1(3), 3, "Identity", 2),
    (lambda: fb.HDIVBDM1(3), 3, "Divergence", 2), (lambda: fb.L2P0(1), 3, "Identity", 2),
    (lambda: fb.HCURLN0(2), 2, "Identity", 2), (lambda: fb.HCURLN0(3), 3, "Identity", 2),
]


@pytest.mark.parametrize("mk,dim,op,qorder", EVAL_CASES)
def test_update_basis(mk, dim, op, qorder):
    """femb_evaluator_eval == oracle update_basis! (feevaluator_h1.jl / feevaluator_hdiv.jl / feevaluator_hcurl.jl) for all cells."""
    ft = mk()
    grid = _grid2d(5, True) if dim == 2 else _grid3d(3, True)
    g = grid.geometry
    fes = fb.FESpace(ft, grid)
    qf = fb.QuadratureRule(g, qorder)
    from femb_b200.assembly import _Session

    S = _Session([fes], qf)
    feb = S.evaluator(0, op)
    got = feb.update_basis_all()  # [cell, qp, dof, r]
    fs, es, ori = _signs(grid, ft)
    want = orc.update_basis(_oracle_el(ft), op, g, qf.xref, grid["Coordinates"], grid["CellNodes"], fs, es, ori)  # [cell, r, dof, qp]
    want = np.transpose(want, (0, 3, 2, 1))
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= RTOL * np.abs(want).max()
    # single-cell update_basis!(FEB, cell) mirrors FEB.cvals[r, dof, qp]
    cv = feb.update_basis(3)
    assert np.abs(cv - np.transpose(want[2], (2, 1, 0))).max() <= RTOL * np.abs(want).max()
    xq = S.ctx.qp_coords(len(qf))
    A, b, _, _ = orc.l2g(grid["Coordinates"], grid["CellNodes"])
    assert np.abs(xq - orc.eval_trafo(A, b, qf.xref)).max() < 1e-15


# ---------------------------------------------------------------------------------------------
# Example200
# ---------------------------------------------------------------------------------------------
def _ex200_oracle(grid, ft, fes, mu, rhs, drop_tol=1e-15):
    fs, es, _ = _signs(grid, ft)
    el = _oracle_el(ft)
    return orc.example200_assemble(el, grid.geometry, grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], fes["CellDofs"], fes.ndofs,
                                   mu, rhs, fs, es, drop_tol), el


def _local_scale(grid, ft, mu):
    fs, es, _ = _signs(grid, ft)
    Aloc, _ = orc.example200_local(_oracle_el(ft), grid.geometry, grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], mu, None, fs, es)
    return np.abs(Aloc).max()


def _ambiguous_free_pattern_equal(A, colptr, rowval, nzval, scale):
    """Pattern equality; entries whose |value| sits in the ambiguous band around Example200's 1e-15
    filter (summation-order dependent, SURVEY.md §7 hard part 1) may differ."""
    import scipy.sparse as sp

    P = sp.csc_matrix((np.ones_like(A.nzval), A.rowval - 1, A.colptr - 1), shape=A.shape)
    Q = sp.csc_matrix((np.ones_like(nzval), rowval - 1, colptr - 1), shape=A.shape)
    D = (P - Q).tocoo()
    if D.nnz == 0:
        return True
    Av = sp.csc_matrix((A.nzval, A.rowval - 1, A.colptr - 1), shape=A.shape)
    Bv = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=A.shape)
    for i, j, v in zip(D.row, D.col, D.data):
        if v == 0:
            continue
        if max(abs(Av[i, j]), abs(Bv[i, j])) > 1e-13 * scale:
            return False
    return True


EX200_CASES = [
    ("p1-3", lambda: fb.H1Pk(1, 2, 1), 2, 3, False), ("p1-63", lambda: fb.H1Pk(1, 2, 1), 2, 63, False), ("p1-40-jit", lambda: fb.H1P1(1), 2, 40, True),
    ("p2-16", lambda: fb.H1Pk(1, 2, 2), 2, 16, False), ("p2-63", lambda: fb.H1Pk(1, 2, 2), 2, 63, False), ("p2-24-jit", lambda: fb.H1P2(1, 2), 2, 24, True),
    ("p3-12-jit", lambda: fb.H1Pk(1, 2, 3), 2, 12, True), ("p1-3d-8", lambda: fb.H1P1(1), 3, 8, False), ("p1-3d-6-jit", lambda: fb.H1P1(1), 3, 6, True),
    ("p2-3d-5", lambda: fb.H1P2(1, 3), 3, 5, False), ("p2-3d-4-jit", lambda: fb.H1P2(1, 3), 3, 4, True), ("p3-3d-3-jit", lambda: fb.H1P3(1, 3), 3, 3, True),
]


@pytest.mark.parametrize("tag,mk,dim,n,jit", EX200_CASES, ids=[c[0] for c in EX200_CASES])
@pytest.mark.parametrize("scatter", [_lib.SCATTER_GATHER, _lib.SCATTER_ATOMIC], ids=["gather", "atomic"])
def test_example200(tag, mk, dim, n, jit, scatter):
    """assemble!(A, b, fe_space, rhs, mu) (Example200:103-184): first assembly builds the filtered
    pattern, second assembly into the same matrix doubles the values (Example200:205-211)."""
    ft = mk()
    grid = _grid2d(n, jit) if dim == 2 else _grid3d(n, jit)
    fes = fb.FESpace(ft, grid)
    mu = 1.3
    rhs = rhs_affine() if dim == 2 else fb.AffineFunction([0.25], [[1.0, -1.0, 0.5]])
    (colptr, rowval, nzval, b), el = _ex200_oracle(grid, ft, fes, mu, rhs)
    scale = _local_scale(grid, ft, mu)

    A = fb.FEMatrix(fes, fes)
    bv = fb.FEVector(fes)
    fb.assemble_poisson(A.entries, bv.entries, fes, rhs, mu, scatter=scatter)
    E = A.entries
    exact_pattern = np.array_equal(E.colptr, colptr) and np.array_equal(E.rowval, rowval)
    if ft.order == 1 or (jit and ft.order == 2):
        assert exact_pattern, f"CSC pattern must be bit-exact (nnz {E.nnz} vs {rowval.shape[0]})"
    else:
        # structured meshes / P3: entries that are 0 in exact arithmetic come out as O(1e-17) rounding noise whose
        # survival of Example200's 1e-15 filter depends on summation order (SURVEY.md §7 hard part 1)
        assert exact_pattern or _ambiguous_free_pattern_equal(E, colptr, rowval, nzval, scale), f"nnz {E.nnz} vs {rowval.shape[0]}"
    if exact_pattern:
        assert np.abs(E.nzval - nzval).max() <= RTOL * scale
    else:
        import scipy.sparse as sp

        D = E.to_scipy() - sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=E.shape)
        assert abs(D).max() <= RTOL * scale
    bscale = np.abs(b).max()
    assert np.abs(bv.entries - b).max() <= RTOL * max(bscale, 1e-300)
    # second assembly into the existing pattern: values add up
    first = E.nzval.copy()
    fb.assemble_poisson(A.entries, bv.entries, fes, rhs, mu, scatter=scatter)
    assert np.abs(E.nzval - 2 * first).max() <= 4 * RTOL * scale
    assert np.abs(bv.entries - 2 * b).max() <= 4 * RTOL * max(bscale, 1e-300)


def test_example200_closure_rhs_and_zero_rhs():
    """rhs given as a host closure -> evaluated at the device-reported quadrature points (FEMB_RHS_QPVALUES)."""
    grid = _grid2d(20, True)
    ft = fb.H1Pk(1, 2, 2)
    fes = fb.FESpace(ft, grid)
    f = lambda x: np.sin(3 * x[..., 0]) * x[..., 1] ** 2  # noqa: E731
    (colptr, rowval, nzval, b), _ = _ex200_oracle(grid, ft, fes, 1.0, f)
    A, bv = fb.FEMatrix(fes, fes), fb.FEVector(fes)
    fb.assemble_poisson(A.entries, bv.entries, fes, f, 1.0)
    assert np.array_equal(A.entries.rowval, rowval)
    assert np.abs(bv.entries - b).max() <= RTOL * np.abs(b).max()


def test_example200_structured_stencil():
    """Appendix C.2: structured P1 -> 5-point stencil (4; -1 -1 -1 -1), hypotenuse entries absent."""
    grid = _grid2d(16)
    fes = fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    A, bv = fb.FEMatrix(fes, fes), fb.FEVector(fes)
    fb.assemble_poisson(A.entries, bv.entries, fes, rhs_affine(), 1.0)
    M = A.entries.to_scipy().tocsr()
    i = 17 * 8 + 8
    row = M.getrow(i)
    assert row.nnz == 5 and M[i, i] == 4.0 and sorted(row.data) == [-1.0, -1.0, -1.0, -1.0, 4.0]
    assert abs(M @ np.ones(M.shape[0])).max() < 1e-13 and abs(M - M.T).max() == 0.0


def test_bit_reproducible_gather():
    """FEMB_SCATTER_GATHER is write-once in a fixed order: two assemblies give identical bits."""
    grid = _grid2d(50, True)
    fes = fb.FESpace(fb.H1P1(1), grid)
    outs = []
    for _ in range(2):
        A, bv = fb.FEMatrix(fes, fes), fb.FEVector(fes)
        fb.assemble_poisson(A.entries, bv.entries, fes, rhs_affine(), 1.0, scatter=_lib.SCATTER_GATHER)
        outs.append((A.entries.nzval.copy(), bv.entries.copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


# ---------------------------------------------------------------------------------------------
# Example210
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,jit", [(4, False), (16, True), (32, False)])
def test_example210(n, jit):
    """prepare_assembly!/update_system!(linear, nonlinear) (Example210:218-394) incl. the Newton-linearised
    convection term around the interpolant of u! (Example210:51-57)."""
    grid = _grid2d(n, jit)
    FESu, FESp = fb.FESpace(fb.H1Pk(2, 2, 2), grid), fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    mu = 1.0e-2
    # current solution = nodal interpolant of the exact velocity at t = 0 on vertices and edge midpoints
    coords, fn = grid["Coordinates"], grid["FaceNodes"]
    pts = np.concatenate([coords, 0.5 * (coords[fn[:, 0] - 1] + coords[fn[:, 1] - 1])])
    u1 = np.sin(2 * np.pi * pts[:, 0]) * np.sin(2 * np.pi * pts[:, 1])
    u2 = np.cos(2 * np.pi * pts[:, 0]) * np.cos(2 * np.pi * pts[:, 1])
    sol = fb.FEVector([FESu, FESp])
    sol.entries[: FESu.ndofs] = np.concatenate([u1, u2])
    sol.entries[FESu.ndofs:] = 0.1

    for linear, nonlinear in [(True, False), (False, True), (True, True)]:
        A, b = fb.FEMatrix([FESu, FESp]), fb.FEVector([FESu, FESp])
        upd = fb.prepare_assembly(A, b, FESu, FESp, sol, mu=mu)
        upd(linear, nonlinear)
        colptr, rowval, nzval, bb = orc.example210_assemble(coords, grid["CellNodes"], grid["CellVolumes"], FESu["CellDofs"], FESp["CellDofs"],
                                                            FESu.ndofs, FESp.ndofs, sol.entries, mu, 0.0, True, nonlinear)
        # the device pattern is the structural one of the (u,u),(u,p),(p,u) blocks + the pinned pressure diagonal
        # (= what update_system!(true, .) + the penalty A[ndofs_u+1, ndofs_u+1] produce, Example210:172,197-199)
        import scipy.sparse as sp

        n_all = FESu.ndofs + FESp.ndofs
        P = sp.csc_matrix((np.ones_like(nzval), rowval - 1, colptr - 1), shape=(n_all,) * 2).tolil()
        P[FESu.ndofs, FESu.ndofs] = 1
        P = P.tocsc()
        P.sort_indices()
        E = A.entries
        assert np.array_equal(E.colptr - 1, P.indptr) and np.array_equal(E.rowval - 1, P.indices)
        Aloc, Bloc, _ = orc.example210_local(coords, grid["CellNodes"], grid["CellVolumes"], FESu["CellDofs"], sol.entries, mu, 0.0, linear, nonlinear)
        scale = max(np.abs(Aloc).max(), np.abs(Bloc).max() if Bloc is not None else 0.0)
        c2, r2, v2, b2 = orc.example210_assemble(coords, grid["CellNodes"], grid["CellVolumes"], FESu["CellDofs"], FESp["CellDofs"],
                                                 FESu.ndofs, FESp.ndofs, sol.entries, mu, 0.0, linear, nonlinear)
        want = sp.csc_matrix((v2, r2 - 1, c2 - 1), shape=(n_all,) * 2)
        assert abs(E.to_scipy() - want).max() <= RTOL * scale
        assert np.abs(b.entries - b2).max() <= RTOL * max(np.abs(b2).max(), 1e-300)


def test_example210_divpressure_owner_computes_and_reductions_agree():
    """Linear part on a freshly zeroed matrix: the (u,p) / (p,u) blocks are stored by the column owners (k_th_divp_gather); on a matrix
    that already holds something they are added by the per-cell kernel.  Assembling twice without zeroing must give twice the first."""
    grid = _grid2d(12, True)
    FESu, FESp = fb.FESpace(fb.H1Pk(2, 2, 2), grid), fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    sol = fb.FEVector([FESu, FESp])
    A, b = fb.FEMatrix([FESu, FESp]), fb.FEVector([FESu, FESp])
    upd = fb.prepare_assembly(A, b, FESu, FESp, sol, mu=0.3)
    upd(True, False)
    ctx = upd.session.ctx
    first = ctx.matrix_get().copy()
    S = upd.session
    egu, eiu, eip = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity"), S.evaluator(1, "Identity")
    ctx.assemble_navierstokes(egu._eid, eiu._eid, eip._eid, 0.3, True, False, _lib.RHS_NONE, None)  # no femb_matrix_zero in between
    twice = ctx.matrix_get()
    assert np.abs(twice - 2.0 * first).max() <= RTOL * np.abs(first).max()
    ctx.matrix_zero()
    ctx.assemble_navierstokes(egu._eid, eiu._eid, eip._eid, 0.3, True, False, _lib.RHS_NONE, None)
    assert np.array_equal(ctx.matrix_get(), first)  # write-once paths: bit-reproducible


# ---------------------------------------------------------------------------------------------
# Hdiv mass + divergence blocks (config C4)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,dim,n", [("HDIVRT0", 2, 9), ("HDIVBDM1", 2, 7), ("HDIVRT0", 3, 4), ("HDIVBDM1", 3, 3)])
def test_hdiv_mixed(name, dim, n):
    grid = _grid2d(n, True) if dim == 2 else _grid3d(n, True)
    _check_hdiv_mixed(grid, {"HDIVRT0": fb.HDIVRT0, "HDIVBDM1": fb.HDIVBDM1}[name](dim))


def _check_hdiv_mixed(grid, ft):
    g = grid.geometry
    FESv, FESq = fb.FESpace(ft, grid), fb.FESpace(fb.L2P0(1), grid)
    A = fb.FEMatrix([FESv, FESq])
    fb.assemble_hdiv_mixed(A, FESv, FESq)
    el, elq = _oracle_el(ft), orc.Element("L2P0", 1)
    xref, w = orc.quadrature(g, 2)
    fs, ori = grid["CellFaceSigns"], grid["CellFaceOrientations"]
    args = (g, xref, grid["Coordinates"], grid["CellNodes"], fs, None, ori)
    V = orc.update_basis(el, "Identity", *args)
    D = orc.update_basis(el, "Divergence", *args)
    Q = orc.update_basis(elq, "Identity", *args)
    vol = grid["CellVolumes"]
    M = orc.bilinear_local(V, V, w, vol)
    B = orc.bilinear_local(D, Q, w, vol)
    I1, J1, V1 = orc.block_triplets(M, FESv["CellDofs"], FESv["CellDofs"])
    I2, J2, V2 = orc.block_triplets(B, FESv["CellDofs"], FESq["CellDofs"], 0, FESv.ndofs)
    nall = FESv.ndofs + FESq.ndofs
    colptr, rowval, nzval = orc.flush_csc(nall, nall, np.concatenate([I1, I2, J2]), np.concatenate([J1, J2, I2]), np.concatenate([V1, V2, V2]))
    E = A.entries
    assert np.array_equal(E.colptr, colptr) and np.array_equal(E.rowval, rowval)
    scale = max(np.abs(M).max(), np.abs(B).max())
    assert np.abs(E.nzval - nzval).max() <= RTOL * scale
    # divergence theorem: B^T 1_v-coefficients of a constant field ... row sums of the (q, v) block = sum_f sign * 1 -> net flux 0 for interior
    Msp = E.to_scipy()
    assert abs(Msp - Msp.T).max() <= RTOL * scale


@pytest.mark.parametrize("dim,device_dofmap", [(2, False), (3, False), (3, True)])
def test_hcurl_n0_mass_system(dim, device_dofmap):
    """HCURLN0 mass matrix and load vector (Identity push-forward feevaluator_hcurl.jl:2-16, coefficients hcurl_n0.jl:144-166) against
    the oracle; solving the system with a constant load returns its tangential-flux interpolant (the L2 projection onto N0 is exact
    for constant fields: the reference's exactness order 0, test/test_interpolators.jl:15,44)."""
    import scipy.sparse.linalg as spla

    grid = _grid2d(6, True) if dim == 2 else _grid3d(4, True)
    if device_dofmap:
        grid.instantiate_on_device()
    g = grid.geometry
    ft = fb.HCURLN0(dim)
    fes = fb.FESpace(ft, grid)
    A = fb.ExtendableSparseMatrixCSC(fes.ndofs, fes.ndofs)
    b = np.zeros(fes.ndofs)
    cvec = np.array([0.7, -1.3, 0.4])[:dim]
    rhs = fb.AffineFunction(list(cvec), [[0.0] * dim] * dim)
    S = fb.assemble_mass(A, b, fes, rhs)
    if device_dofmap:  # signs and CellDofs generated on the device from the pattern "e1" (femb_space_set_from_pattern): same system
        from femb_b200.assembly import _Session

        S2 = _Session([fes], S.qf, device_dofmaps=True)
        ei = S2.evaluator(0, "Identity")
        S2.ctx.pattern_build([(0, 0)])
        S2.ctx.matrix_zero()
        S2.ctx.assemble_bilinear(ei._eid, ei._eid, 0, 0, 1.0, 0, 0.0)
        cp, rv = S2.ctx.pattern_get(fes.ndofs)
        assert np.array_equal(cp, A.colptr) and np.array_equal(rv, A.rowval)
        assert np.abs(S2.ctx.matrix_get() - A.nzval).max() <= RTOL * np.abs(A.nzval).max()  # (atomic scatter: summation order differs)
        S2.ctx.close()
    el = orc.Element("HCURLN0", dim)
    xref, w = orc.quadrature(g, 2)
    fs, es, _ = _signs(grid, ft)
    fs = grid["CellFaceSigns"]
    V = orc.update_basis(el, "Identity", g, xref, grid["Coordinates"], grid["CellNodes"], fs, es)
    vol = grid["CellVolumes"]
    M = orc.bilinear_local(V, V, w, vol)
    cd = fes["CellDofs"]
    I, J, Vv = orc.block_triplets(M, cd, cd)
    colptr, rowval, nzval = orc.flush_csc(fes.ndofs, fes.ndofs, I, J, Vv)
    assert np.array_equal(A.colptr, colptr) and np.array_equal(A.rowval, rowval)
    scale = np.abs(M).max()
    assert np.abs(A.nzval - nzval).max() <= RTOL * scale
    # load vector: b_j = sum_cells |T| sum_qp w_qp <phi_j(x_qp), c>
    bl = np.einsum("crdq,r,q,c->cd", V, cvec, w, vol)
    bo = np.zeros(fes.ndofs)
    np.add.at(bo, cd.astype(np.int64) - 1, bl)
    assert np.abs(b - bo).max() <= RTOL * max(np.abs(bl).max(), 1e-300) * 8
    # M x = b  =>  x = interpolant of the constant field
    x = spla.spsolve(A.to_scipy().tocsc(), b)
    enodes = grid["FaceNodes"] if dim == 2 else grid["EdgeNodes"]
    want = orc.interpolate_hcurl(grid["Coordinates"], enodes, lambda p: np.broadcast_to(cvec, p.shape))
    assert np.abs(x - want).max() < 1e-10 * np.abs(want).max()


def test_hcurl_unsupported_operators_are_statuses():
    grid = _grid2d(3)
    fes = fb.FESpace(fb.HCURLN0(2), grid)
    from femb_b200.assembly import _Session

    S = _Session([fes], fb.QuadratureRule(grid.geometry, 2))
    with pytest.raises(_lib.FembError):
        S.evaluator(0, "Gradient")


def test_hdiv_blockwise_lazy_zero():
    """After femb_matrix_zero the RT0 kernels claim the blocks they overwrite entry by entry and the memset is skipped; everything
    the reference's fill!(A.nzval, 0) + assembly sequence guarantees must still hold: poisoned storage is never visible, blocks no
    kernel wrote read as zeros, a second assembly without zeroing accumulates, and a pattern with entries beyond the (face, cell)
    couplings falls back to zero + add."""
    import scipy.sparse as sp
    from femb_b200.assembly import _Session

    grid = _grid3d(4, True)
    FESv, FESq = fb.FESpace(fb.HDIVRT0(3), grid), fb.FESpace(fb.L2P0(1), grid)
    S = _Session([FESv, FESq], fb.QuadratureRule(grid.geometry, 2))
    ctx = S.ctx
    eid, ediv, eq = S.evaluator(0, "Identity"), S.evaluator(0, "Divergence"), S.evaluator(1, "Identity")
    ctx.pattern_build([(0, 0), (0, 1), (1, 0)])
    nall = FESv.ndofs + FESq.ndofs

    def mass():
        ctx.assemble_bilinear(eid._eid, eid._eid, 0, 0, 1.0, 0, 0.0)

    def div():
        ctx.assemble_bilinear(ediv._eid, eq._eid, 0, 1, 1.0, _lib.BIL_TRANSPOSE_COPY, 0.0)

    def poison():  # fill the device array with garbage through an accumulating call sequence, then request the lazy zero
        ctx.matrix_zero(); mass(); div(); mass(); div(); mass()
        ctx.matrix_zero()

    ctx.matrix_zero(); mass(); div()
    ref = ctx.matrix_get().copy()
    colptr, rowval = ctx.pattern_get(nall)
    cols = np.repeat(np.arange(nall), np.diff(colptr))
    in_mass = (rowval <= FESv.ndofs) & (cols < FESv.ndofs)
    assert np.abs(ref[in_mass]).max() > 0 and np.abs(ref[~in_mass]).max() > 0
    # both orders, repeated
    for order in ((mass, div), (div, mass)):
        poison()
        for f in order:
            f()
        assert np.array_equal(ctx.matrix_get(), ref)
    # the same three blocks in one call (femb_assemble_hdiv_mixed): fresh matrix -> two kernels writing every column once;
    # matrix already holding something -> the two assemblies, added
    mixed = lambda: ctx.assemble_hdiv_mixed(eid._eid, ediv._eid, eq._eid, 1.0, 1.0)
    poison(); mixed()
    assert np.array_equal(ctx.matrix_get(), ref)
    mixed()
    assert np.abs(ctx.matrix_get() - 2.0 * ref).max() <= RTOL * np.abs(ref).max()
    poison(); mass(); mixed()
    got = ctx.matrix_get()
    assert np.abs(got[in_mass] - 2.0 * ref[in_mass]).max() <= RTOL * np.abs(ref).max() and np.array_equal(got[~in_mass], ref[~in_mass])
    poison()
    ctx.assemble_hdiv_mixed(eid._eid, ediv._eid, eq._eid, 0.5, -2.0)
    got = ctx.matrix_get()
    assert np.array_equal(got[in_mass], 0.5 * ref[in_mass]) and np.array_equal(got[~in_mass], -2.0 * ref[~in_mass])
    # only one of the two: the blocks nobody wrote are zeros
    poison(); mass()
    got = ctx.matrix_get()
    assert np.array_equal(got[in_mass], ref[in_mass]) and not got[~in_mass].any()
    poison(); div()
    got = ctx.matrix_get()
    assert np.array_equal(got[~in_mass], ref[~in_mass]) and not got[in_mass].any()
    # no zeroing in between: accumulation (mass block: gather adds; divergence blocks: plain adds)
    poison(); mass(); div(); div(); mass()
    assert np.abs(ctx.matrix_get() - 2.0 * ref).max() <= RTOL * np.abs(ref).max()
    # claim, then a consumer in between (femb_matrix_get settles the pending blocks), then the rest
    poison(); mass()
    ctx.matrix_get()
    div()
    assert np.array_equal(ctx.matrix_get(), ref)
    # a pattern with additional entries in the divergence blocks: they must read as zeros, the rest is unchanged
    A = sp.csc_matrix((ref, rowval - 1, colptr - 1), shape=(nall, nall))
    cd = FESv["CellDofs"]
    free = lambda c: int(np.setdiff1d(np.arange(1, 12), cd[c])[0]) - 1  # a face dof (0-based) that is not one of cell c's
    ex = [(free(1), FESv.ndofs + 1), (free(7), FESv.ndofs + 7), (FESv.ndofs + 2, free(2))]
    extra = sp.csc_matrix((np.ones(3), ([e[0] for e in ex], [e[1] for e in ex])), shape=(nall, nall))
    P = (abs(A) + extra).tocsc()
    P.sort_indices()
    ctx.pattern_adopt(nall, nall, P.indptr.astype(np.int64) + 1, P.indices.astype(np.int64) + 1, [(0, 0), (0, 1), (1, 0)])
    for run in (lambda: (mass(), div()), mixed):
        ctx.matrix_zero(); mass(); div(); mass(); div()
        ctx.matrix_zero(); run()
        got = sp.csc_matrix((ctx.matrix_get(), P.indices, P.indptr), shape=(nall, nall))
        assert abs(got - A).max() <= RTOL * np.abs(ref).max()
        assert P.nnz == A.nnz + 3 and all(got[i, j] == 0.0 for i, j in ex)
    ctx.close()


# ---------------------------------------------------------------------------------------------
# error behaviour
# ---------------------------------------------------------------------------------------------
def test_errors_are_statuses_not_crashes():
    with _lib.Context(0) as ctx:
        with pytest.raises(_lib.FembError):
            ctx.quadrature_set(np.zeros((1, 2)), np.ones(1))  # mesh not set
        with pytest.raises(_lib.FembError):
            ctx.mesh_set(np.zeros((4, 2)), np.ones((2, 4), dtype=np.int32))  # quads: unsupported
        ctx.mesh_set(np.array([[0.0, 0], [1, 0], [0, 1]]), np.array([[1, 2, 3]], dtype=np.int32))
        with pytest.raises(_lib.FembError):
            ctx.evaluator_set(0, 3, 0, np.zeros(3))  # unknown space
        assert ctx.launch_count() == 0


def test_adopted_pattern_missing_entry_is_reported():
    """A contribution without a slot in an adopted pattern -> FEMB_ERR_PATTERN (never silently dropped)."""
    grid = _grid2d(6, True)
    fes = fb.FESpace(fb.H1P1(1), grid)
    A, bv = fb.FEMatrix(fes, fes), fb.FEVector(fes)
    fb.assemble_poisson(A.entries, bv.entries, fes, None, 1.0)
    E = A.entries
    # remove one off-diagonal entry from the pattern
    col = 10
    lo, hi = E.colptr[col] - 1, E.colptr[col + 1] - 1
    kill = lo if E.rowval[lo] != col + 1 else lo + 1
    keep = np.ones(E.nnz, dtype=bool)
    keep[kill] = False
    colptr = E.colptr.copy()
    colptr[col + 1:] -= 1
    B = fb.FEMatrix(fes, fes)
    B.entries.set_csc(colptr, E.rowval[keep], np.zeros(keep.sum()))
    fes2 = fb.FESpace(fb.H1P1(1), grid)
    with pytest.raises(_lib.FembError) as ei:
        fb.assemble_poisson(B.entries, bv.entries, fes2, None, 1.0)
    assert ei.value.status == 4


@pytest.mark.parametrize("dim,n,nchunks,use_vol", [(2, 70, 5, True), (2, 33, 1, False), (2, 200, 16, True), (3, 9, 4, True)])
def test_streamed_host_assembly_is_bit_identical(dim, n, nchunks, use_vol):
    """femb_assemble_poisson_host (upload / kernel / download pipelined over column chunks) == mesh_set + assemble + get."""
    from femb_b200.assembly import _Session

    grid = _grid2d(n, True) if dim == 2 else _grid3d(n, True)
    fes = fb.FESpace(fb.H1P1(1), grid)
    qf = fb.QuadratureRule(grid.geometry, 0)
    S = _Session([fes], qf, cellvolumes=use_vol)
    ctx = S.ctx
    eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
    nnz = ctx.pattern_build([(0, 0)])
    p = np.array([0.3, 1.0, -1.0] if dim == 2 else [0.3, 1.0, -1.0, 0.5])
    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg._eid, ei._eid, 1.7, _lib.RHS_AFFINE, p, 1e-15)
    nz1, b1 = ctx.matrix_get(), ctx.vector_get(fes.ndofs)
    # new coordinates for the streamed call (same topology): results must follow the uploaded mesh
    coords2 = np.ascontiguousarray(grid["Coordinates"] * 1.25 + 0.1)
    vol2 = np.ascontiguousarray(grid["CellVolumes"] * 1.25 ** dim) if use_vol else None
    nz2, b2 = np.full(nnz, np.nan), np.full(fes.ndofs, np.nan)
    ctx.assemble_poisson_host(eg._eid, ei._eid, 1.7, _lib.RHS_AFFINE, p, 1e-15, coords2, grid["CellNodes"], vol2, nz2, b2, nchunks)
    ctx.mesh_set(coords2, grid["CellNodes"], vol2)
    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg._eid, ei._eid, 1.7, _lib.RHS_AFFINE, p, 1e-15)
    nz3, b3 = ctx.matrix_get(), ctx.vector_get(fes.ndofs)
    assert np.array_equal(nz2, nz3) and np.array_equal(b2, b3)
    assert not np.array_equal(nz1, nz3) or dim == 2  # 2-D stiffness is scale invariant, 3-D is not
    assert not np.array_equal(b1, b3)


@pytest.mark.parametrize("n", [1500, 5657])
def test_full_size_p1_against_c_oracle(n):
    """BASELINE.json configs[1] at full size (n = 5657: 64,003,298 triangles): the device result against the C
    restatement of the reference loop (all host threads) on the SAME pattern — every reference contribution finds
    a slot (pattern >= reference), every slot is non-zero (pattern <= reference), values within 1e-12 — plus the
    size-independent identities A 1 = 0, A = A^T, sum(b) = int f."""
    import oracle_c
    from femb_b200.assembly import _Session

    X = np.linspace(0, 1, n + 1)
    grid = fb.simplexgrid(X, X)
    fes = fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    qf = fb.QuadratureRule(grid.geometry, 0)
    S = _Session([fes], qf)
    ctx = S.ctx
    eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
    ctx.pattern_track(True)
    ctx.pattern_build([(0, 0)])
    p = np.array([0.0, 1.0, -1.0])
    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 1e-15)
    nnz = ctx.pattern_compress()
    ctx.pattern_track(False)
    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 1e-15)
    nz, b = ctx.matrix_get(), ctx.vector_get(fes.ndofs)
    colptr, rowval = ctx.pattern_get(fes.ndofs)
    assert nnz == rowval.shape[0] and np.all(np.diff(colptr) >= 3) and np.all(np.diff(colptr) <= 5)  # 5-point stencil after the 1e-15 filter
    el = orc.Element("H1Pk", 1, 1)
    xref, w = orc.quadrature("Triangle2D", 0)
    vals, der = el.tabulate("Triangle2D", xref)
    nz2, b2 = np.zeros(nnz), np.zeros(fes.ndofs)
    cn = grid["CellNodes"]
    miss = oracle_c.example200(grid["Coordinates"], cn, grid["CellVolumes"], cn, w, xref, vals[:, :, 0], der[:, :, 0, :], 1.0, p, 1e-15, colptr, rowval, nz2, b2,
                               oracle_c.max_threads())
    assert miss == 0 and np.all(nz2 != 0.0)
    assert np.abs(nz - nz2).max() <= RTOL * np.abs(nz2).max()
    assert np.abs(b - b2).max() <= RTOL * np.abs(b2).max()
    cols = np.repeat(np.arange(fes.ndofs, dtype=np.int64), np.diff(colptr))
    assert np.abs(np.bincount(rowval - 1, weights=nz, minlength=fes.ndofs)).max() < 1e-13  # A 1 = 0
    assert np.abs(np.bincount(cols, weights=nz, minlength=fes.ndofs)).max() < 1e-13  # 1^T A = 0
    # exact symmetry: value of (i, j) equals value of (j, i)
    key = cols * fes.ndofs + (rowval - 1)
    keyT = (rowval - 1) * fes.ndofs + cols
    assert np.array_equal(nz[np.searchsorted(key, keyT)], nz)
    assert abs(b.sum()) < 1e-12  # int (x - y) over the unit square


# ---------------------------------------------------------------------------------------------
# edge cases: tiny / ragged / mirrored inputs
# ---------------------------------------------------------------------------------------------
def _assemble_and_compare(grid, ft, rhs, mu=0.7):
    fes = fb.FESpace(ft, grid)
    (colptr, rowval, nzval, b), _ = _ex200_oracle(grid, ft, fes, mu, rhs)
    scale = _local_scale(grid, ft, mu)
    for scatter in (_lib.SCATTER_GATHER, _lib.SCATTER_ATOMIC):
        A, bv = fb.FEMatrix(fes, fes), fb.FEVector(fes)
        fes2 = fb.FESpace(ft, grid)  # fresh session per policy
        fb.assemble_poisson(A.entries, bv.entries, fes2, rhs, mu, scatter=scatter)
        E = A.entries
        import scipy.sparse as sp

        D = E.to_scipy() - sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=E.shape)
        assert abs(D).max() <= RTOL * scale
        assert np.abs(bv.entries - b).max() <= RTOL * max(np.abs(b).max(), 1e-300)
        if ft.order == 1:
            assert np.array_equal(E.colptr, colptr) and np.array_equal(E.rowval, rowval)


def test_single_cells():
    """one triangle (the non-reference triangle of test/test_operators.jl:83) and one stretched tetrahedron"""
    _assemble_and_compare(fb.grid_triangle([[-1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), fb.H1Pk(1, 2, 1), rhs_affine())
    _assemble_and_compare(fb.grid_triangle([[-1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), fb.H1Pk(1, 2, 2), rhs_affine())
    tet = fb.grid_tetrahedron([[0.0, 0, 0], [2.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]])
    _assemble_and_compare(tet, fb.H1P1(1), fb.AffineFunction([0.25], [[1.0, -1.0, 0.5]]))
    _assemble_and_compare(tet, fb.H1P2(1, 3), fb.AffineFunction([0.25], [[1.0, -1.0, 0.5]]))
    # Appendix C.1 golden: P1 stiffness of that triangle
    fes = fb.FESpace(fb.H1P1(1), fb.grid_triangle([[-1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]))
    A, bv = fb.FEMatrix(fes, fes), fb.FEVector(fes)
    fb.assemble_poisson(A.entries, bv.entries, fes, None, 1.0)
    assert np.allclose(A.entries.to_scipy().toarray(), [[0.5, 0, -0.5], [0, 0.5, -0.5], [-0.5, -0.5, 1.0]], atol=1e-15)


def test_mirrored_cells_and_ragged_sizes():
    """negatively oriented cells (two nodes swapped: det A < 0), node counts that are not multiples of the 32-dof
    slices, 2 cells only"""
    for n in (1, 2, 5, 11):
        grid = _grid2d(n, n > 2)
        cn = grid["CellNodes"].copy()
        cn[::2] = cn[::2][:, [0, 2, 1]]  # mirror every other cell
        g2 = fb.ExtendableGrid(grid["Coordinates"], cn, "Triangle2D")
        _assemble_and_compare(g2, fb.H1Pk(1, 2, 1), rhs_affine())
    grid = _grid3d(2, True)
    cn = grid["CellNodes"].copy()
    cn[1::3] = cn[1::3][:, [0, 1, 3, 2]]
    g3 = fb.ExtendableGrid(grid["Coordinates"], cn, "Tetrahedron3D")
    _assemble_and_compare(g3, fb.H1P1(1), fb.AffineFunction([0.25], [[1.0, -1.0, 0.5]]))


def test_empty_cell_range_and_unused_nodes():
    """nodes that belong to no cell (halo nodes of a partitioned mesh) get empty columns; evaluating an empty cell range is a no-op"""
    grid = _grid2d(4, True)
    coords = np.concatenate([grid["Coordinates"], [[2.0, 2.0], [3.0, 3.0], [4.0, 2.5]]])
    g2 = fb.ExtendableGrid(coords, grid["CellNodes"], "Triangle2D")
    fes = fb.FESpace(fb.H1P1(1), g2)
    A, bv = fb.FEMatrix(fes, fes), fb.FEVector(fes)
    fb.assemble_poisson(A.entries, bv.entries, fes, rhs_affine(), 1.0)
    E = A.entries
    assert np.all(np.diff(E.colptr)[-3:] == 0) and np.all(bv.entries[-3:] == 0.0)
    S = getattr(fes, "_femb_session_0")
    out = S.evaluator(0, "Gradient").update_basis_all(3, 3)
    assert out.shape[0] == 0


# ---------------------------------------------------------------------------------------------
# the examples end to end (device assembly + host solve), as SURVEY.md §8(c) asks
# ---------------------------------------------------------------------------------------------
def _load_example(name):
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", name + ".py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_example200_solution_matches_oracle_solution():
    """Example200 main loop: the Poisson solution from the device-assembled system equals the one from the oracle system"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    ex = _load_example("example200_lowlevel_poisson")
    X = np.linspace(0, 1, 17)
    grid = fb.simplexgrid(X, X)
    ft = fb.H1Pk(1, 2, 2)
    fes = fb.FESpace(ft, grid)
    rhs = rhs_affine()
    solution, _, _ = ex.solve_poisson_lowlevel(fes, 1.0, rhs)
    (colptr, rowval, nzval, b), _ = _ex200_oracle(grid, ft, fes, 1.0, rhs)
    M = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(fes.ndofs,) * 2).tolil()
    for d in fb.boundarydofs(fes):
        M[d - 1, d - 1] = 1.0e60
        b[d - 1] = 0.0
    u = spla.spsolve(M.tocsc(), b)
    assert np.abs(solution.entries - u).max() <= 1e-10 * max(np.abs(u).max(), 1e-300)
    assert np.abs(u).max() > 1e-4  # non-trivial solution


def test_example210_newton_converges():
    """Example210: Newton iteration on the device-assembled Taylor-Hood system reaches the reference's stopping criterion
    (nonlinear residual < max(1e-12, 20 linres), Example210:207-209) and the discrete solution approximates the exact one"""
    ex = _load_example("example210_lowlevel_navierstokes")
    errs = []
    for nref in (3, 4):
        sol, nlres, its, err = ex.main(nref=nref)
        assert its <= 12 and nlres < 1e-8
        errs.append(err)
    assert errs[1] < errs[0] / 4  # at least second-order decay of the nodal velocity error under refinement


def test_example210_device_resident_loop_matches_host_loop():
    """The Newton loop with penalisation, saved linear part and both residual norms on the device (femb_system.cu) takes the
    same iterations to the same solution as the host loop"""
    ex = _load_example("example210_lowlevel_navierstokes")
    sol_h, nl_h, its_h, err_h = ex.main(nref=4)
    sol_d, nl_d, its_d, err_d = ex.main(nref=4, device_loop=True)
    assert its_d == its_h and nl_d < 1e-8
    assert np.abs(sol_d.entries - sol_h.entries).max() <= 1e-9 * np.abs(sol_h.entries).max()


# ---------------------------------------------------------------------------------------------
# FEMB_SCATTER_COLORED: cell colouring, plain read-modify-write per colour, bit-reproducible
# ---------------------------------------------------------------------------------------------
def test_colored_scatter_is_reproducible_and_correct():
    import scipy.sparse as sp

    # generic forms: Hdiv mass + divergence blocks (two spaces in the layout)
    grid = _grid3d(4, True)
    g = grid.geometry
    runs = []
    for _ in range(2):
        FESv, FESq = fb.FESpace(fb.HDIVBDM1(3), grid), fb.FESpace(fb.L2P0(1), grid)
        A = fb.FEMatrix([FESv, FESq])
        fb.assemble_hdiv_mixed(A, FESv, FESq, scatter=_lib.SCATTER_COLORED)
        runs.append(A.entries.nzval.copy())
    assert np.array_equal(runs[0], runs[1])
    FESv, FESq = fb.FESpace(fb.HDIVBDM1(3), grid), fb.FESpace(fb.L2P0(1), grid)
    B = fb.FEMatrix([FESv, FESq])
    fb.assemble_hdiv_mixed(B, FESv, FESq, scatter=_lib.SCATTER_ATOMIC)
    assert np.abs(runs[0] - B.entries.nzval).max() <= RTOL * np.abs(runs[0]).max()
    # Example200 with a P3 element (per-cell basis subsets -> generic kernel) and with P1 / P2 forced onto the coloured path
    for ft, grid in ((fb.H1Pk(1, 2, 3), _grid2d(9, True)), (fb.H1Pk(1, 2, 1), _grid2d(14, True)), (fb.H1P2(1, 3), _grid3d(3, True))):
        rhs = rhs_affine() if grid.dim == 2 else fb.AffineFunction([0.25], [[1.0, -1.0, 0.5]])
        outs = []
        for _ in range(2):
            fes = fb.FESpace(ft, grid)
            A, bv = fb.FEMatrix(fes, fes), fb.FEVector(fes)
            fb.assemble_poisson(A.entries, bv.entries, fes, rhs, 0.9, scatter=_lib.SCATTER_COLORED)
            outs.append((A.entries.nzval.copy(), A.entries.rowval.copy()))
        assert np.array_equal(outs[0][0], outs[1][0])
        fes = fb.FESpace(ft, grid)
        (colptr, rowval, nzval, b), _ = _ex200_oracle(grid, ft, fes, 0.9, rhs)
        D = A.entries.to_scipy() - sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=A.entries.shape)
        assert abs(D).max() <= RTOL * _local_scale(grid, ft, 0.9)


# ---------------------------------------------------------------------------------------------
# face / edge enumeration and the pattern-driven CellDofs generator on the device, bit-exact against
# the host restatement (grids.py / fespace.py) — all integer work
# ---------------------------------------------------------------------------------------------
def _shuffled(grid, seed=3):
    """Same mesh with the cells in random order and the local node order rotated (keeps the orientation)."""
    rng = np.random.default_rng(seed)
    cn = grid["CellNodes"][rng.permutation(grid.ncells)].copy()
    if grid.dim == 2:
        rot = rng.integers(0, 3, cn.shape[0])
        cn = np.stack([np.roll(c, r) for c, r in zip(cn, rot)]).astype(np.int32)
    else:
        ev = np.array([[0, 1, 2, 3], [1, 2, 0, 3], [2, 0, 1, 3], [1, 0, 3, 2], [3, 2, 1, 0], [0, 3, 1, 2]])  # even permutations
        cn = np.stack([c[ev[r]] for c, r in zip(cn, rng.integers(0, len(ev), cn.shape[0]))]).astype(np.int32)
    return fb.ExtendableGrid(grid["Coordinates"], cn, grid.geometry)


@pytest.mark.parametrize("dim,n,shuffle", [(2, 1, False), (2, 7, False), (2, 12, True), (3, 1, False), (3, 4, False), (3, 5, True)])
def test_device_grid_entities_match_host(dim, n, shuffle):
    grid = _grid2d(n) if dim == 2 else _grid3d(n)
    if shuffle:
        grid = _shuffled(grid)
    dev = fb.ExtendableGrid(grid["Coordinates"], grid["CellNodes"], grid.geometry).instantiate_on_device()
    names = ["CellFaces", "FaceNodes", "CellFaceSigns", "CellFaceOrientations", "FaceCells", "CellEdges", "EdgeNodes", "CellEdgeSigns"]
    for name in names:
        assert name in dev._c, name  # filled by the device, not lazily by numpy
        host = grid[name]
        assert dev[name].dtype == np.int32 and dev[name].shape == host.shape, name
        assert np.array_equal(dev[name], host), name


DOFMAP_CASES = [
    (lambda: fb.H1P1(1), 2), (lambda: fb.H1P1(2), 2), (lambda: fb.H1P2(1, 2), 2), (lambda: fb.H1Pk(2, 2, 2), 2), (lambda: fb.H1P3(1, 2), 2),
    (lambda: fb.H1Pk(1, 2, 4), 2), (lambda: fb.HDIVRT0(2), 2), (lambda: fb.HDIVBDM1(2), 2), (lambda: fb.L2P0(1), 2),
    (lambda: fb.H1P1(1), 3), (lambda: fb.H1P2(1, 3), 3), (lambda: fb.H1P2(3, 3), 3), (lambda: fb.H1P3(1, 3), 3), (lambda: fb.HDIVRT0(3), 3),
    (lambda: fb.HDIVBDM1(3), 3), (lambda: fb.L2P0(3), 3), (lambda: fb.HCURLN0(2), 2), (lambda: fb.HCURLN0(3), 3),
]


@pytest.mark.parametrize("mk,dim", DOFMAP_CASES)
def test_device_celldofs_match_host(mk, dim):
    from femb_b200.assembly import _register_space

    ft = mk()
    grid = _shuffled(_grid2d(6) if dim == 2 else _grid3d(3), seed=11)
    fes = fb.FESpace(ft, grid)
    ctx = _lib.Context(0)
    ctx.mesh_set(grid["Coordinates"], grid["CellNodes"], None)
    _register_space(ctx, 0, fes, device_dofmap=True)  # raises if ndofs / coffset differ
    nd = ft.get_ndofs(grid.geometry)
    assert np.array_equal(ctx.space_get_celldofs(0, nd), fes["CellDofs"])
    ctx.close()


def test_device_dofmaps_assemble_like_host_dofmaps():
    """Example200 P2-3D and the Hdiv blocks with CellDofs / signs / orientations generated on the device: same matrix."""
    from femb_b200.assembly import _Session

    grid = _grid3d(3, True)
    res = []
    for dev in (False, True):
        fes = fb.FESpace(fb.H1P2(1, 3), grid)
        qf = fb.QuadratureRule(grid.geometry, 2)
        S = _Session([fes], qf, device_dofmaps=dev)
        eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
        S.ctx.pattern_build([(0, 0)])
        S.ctx.matrix_zero(); S.ctx.vector_zero()
        S.ctx.assemble_poisson(eg._eid, ei._eid, 0.7, _lib.RHS_AFFINE, fb.AffineFunction([0.25], [[1.0, -1.0, 0.5]]).params(), 0.0)
        res.append((S.ctx.pattern_get(fes.ndofs), S.ctx.matrix_get(), S.ctx.vector_get(fes.ndofs)))
        S.ctx.close()
    assert all(np.array_equal(a, b) for a, b in zip(res[0][0], res[1][0]))
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    res = []
    for dev in (False, True):
        FESv, FESq = fb.FESpace(fb.HDIVBDM1(3), grid), fb.FESpace(fb.L2P0(1), grid)
        S = _Session([FESv, FESq], fb.QuadratureRule(grid.geometry, 2), device_dofmaps=dev)
        eid, ediv, eq = S.evaluator(0, "Identity"), S.evaluator(0, "Divergence"), S.evaluator(1, "Identity")
        S.ctx.pattern_build([(0, 0), (0, 1), (1, 0)])
        S.ctx.set_scatter(_lib.SCATTER_COLORED)  # reproducible sums
        S.ctx.matrix_zero()
        S.ctx.assemble_bilinear(eid._eid, eid._eid, 0, 0, 1.0, 0, 0.0)
        S.ctx.assemble_bilinear(ediv._eid, eq._eid, 0, 1, 1.0, _lib.BIL_TRANSPOSE_COPY, 0.0)
        res.append(S.ctx.matrix_get())
        S.ctx.close()
    assert np.array_equal(res[0], res[1])


# ---------------------------------------------------------------------------------------------
# boundarydofs, Dirichlet penalisation and the residual A x - b on the device
# ---------------------------------------------------------------------------------------------
BDOF_CASES = [(lambda: fb.H1P1(1), 2), (lambda: fb.H1Pk(2, 2, 2), 2), (lambda: fb.H1P3(1, 2), 2), (lambda: fb.H1Pk(1, 2, 4), 2),
              (lambda: fb.H1P1(1), 3), (lambda: fb.H1P2(1, 3), 3), (lambda: fb.H1P2(3, 3), 3), (lambda: fb.H1P3(1, 3), 3)]


@pytest.mark.parametrize("mk,dim", BDOF_CASES)
def test_device_boundarydofs_match_host(mk, dim):
    from femb_b200.assembly import _register_space
    from femb_b200.fespace import boundarydofs, parse_pattern

    ft = mk()
    grid = _shuffled(_grid2d(5) if dim == 2 else _grid3d(3), seed=4)
    fes = fb.FESpace(ft, grid)
    ctx = _lib.Context(0)
    ctx.mesh_set(grid["Coordinates"], grid["CellNodes"], None)
    for dev in (False, True):  # host-provided and device-generated CellDofs
        _register_space(ctx, 0, fes, device_dofmap=dev)
        got = ctx.boundary_dofs(0, parse_pattern(ft.get_dofmap_pattern(grid.geometry)))
        assert np.array_equal(got, boundarydofs(fes))
    ctx.close()


def test_device_dirichlet_and_residual_match_host():
    """Example200 (P2, 2-D): penalised system and residual of a random vector, device vs scipy on the downloaded matrix;
    Example210 layout: extra pressure dof, boundary values."""
    import scipy.sparse as sp
    from femb_b200.assembly import _Session
    from femb_b200.fespace import boundarydofs, parse_pattern

    grid = _grid2d(9, True)
    FESu, FESp = fb.FESpace(fb.H1Pk(2, 2, 2), grid), fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    n = FESu.ndofs + FESp.ndofs
    S = _Session([FESu, FESp], fb.QuadratureRule(grid.geometry, 5))
    ctx = S.ctx
    egu, eiu, eip = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity"), S.evaluator(1, "Identity")
    ctx.pattern_build([(0, 0), (0, 1), (1, 0)], [FESu.ndofs + 1])
    rng = np.random.default_rng(2)
    sol = rng.uniform(-1, 1, n)
    ctx.vector_set(1, sol)
    ctx.matrix_zero(); ctx.vector_zero()
    ctx.assemble_navierstokes(egu._eid, eiu._eid, eip._eid, 1e-2, True, True, _lib.RHS_EX210, fb.Example210RHS(1e-2, 0.0).params())
    colptr, rowval = ctx.pattern_get(n)
    A0 = sp.csc_matrix((ctx.matrix_get(), rowval - 1, colptr - 1), shape=(n, n))
    b0 = ctx.vector_get(n)
    # residual without fixed rows
    nrm, res = ctx.residual(None, zero_fixed=False, want_vector=True, n=n)
    ref = A0 @ sol - b0
    scale = np.abs(A0).max() * np.abs(sol).max()
    assert np.abs(res - ref).max() <= 1e-12 * scale * 50
    assert abs(nrm - np.linalg.norm(ref)) <= 1e-12 * np.linalg.norm(ref)
    # Dirichlet rows: boundary dofs of the velocity space + one pressure dof, values from a vector
    bd = ctx.boundary_dofs(0, parse_pattern(FESu.fetype.get_dofmap_pattern(grid.geometry)))
    assert np.array_equal(bd, boundarydofs(FESu))
    uinit = rng.uniform(-1, 1, n)
    ctx.dirichlet_apply(True, 0, np.array([FESu.ndofs + 1]), uinit, 1.0e60)
    A1 = sp.csc_matrix((ctx.matrix_get(), rowval - 1, colptr - 1), shape=(n, n)).tolil()
    b1 = ctx.vector_get(n)
    Ah, bh = A0.tolil(), b0.copy()
    fixed = np.concatenate([bd, [FESu.ndofs + 1]]) - 1
    for d in fixed:
        Ah[d, d] = 1.0e60
        bh[d] = 1.0e60 * uinit[d]
    assert (abs(A1 - Ah)).max() == 0.0 and np.array_equal(b1, bh)
    x2 = rng.uniform(-1, 1, n)
    nrm2, res2 = ctx.residual(x2, zero_fixed=True, want_vector=True, n=n)
    ref2 = Ah.tocsc() @ x2 - bh
    ref2[fixed] = 0.0
    assert np.abs(res2 - ref2).max() <= 1e-12 * scale * 50
    assert abs(nrm2 - np.linalg.norm(ref2)) <= 1e-12 * np.linalg.norm(ref2)
    # twice the same: bit-identical (write-once row sums)
    nrm3, res3 = ctx.residual(x2, zero_fixed=True, want_vector=True, n=n)
    assert nrm3 == nrm2 and np.array_equal(res3, res2)
    ctx.close()


# ---------------------------------------------------------------------------------------------
# compute_error (Example210:110-154) on the device vs the oracle restatement
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mk,dim,op", [(lambda: fb.H1Pk(2, 2, 2), 2, "Identity"), (lambda: fb.H1Pk(1, 2, 1), 2, "Identity"), (lambda: fb.H1P3(1, 2), 2, "Gradient"),
                                       (lambda: fb.H1P2(1, 3), 3, "Identity"), (lambda: fb.HDIVBDM1(3), 3, "Identity")])
def test_compute_error_matches_oracle(mk, dim, op):
    ft = mk()
    grid = _grid2d(6, True) if dim == 2 else _grid3d(3, True)
    fes = fb.FESpace(ft, grid)
    rng = np.random.default_rng(8)
    uh = rng.uniform(-1, 1, fes.ndofs)
    el = _oracle_el(ft)
    fs, es, ori = _signs(grid, ft)
    R = {"Identity": ft.ncomponents if ft.family != "Hdiv" else dim, "Gradient": ft.ncomponents * dim}[op]

    def u(x):
        return np.stack([np.sin(3 * x[..., 0]) * (r + 1) + x[..., -1] ** 2 for r in range(R)], axis=-1)

    for exact in (u, None):
        got = fb.compute_error(fes, uh, exact, order=2, p=2, operator=op)
        ref = orc.compute_error(el, grid.geometry, grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], fes["CellDofs"], uh, exact, 2, 2, op, fs, es, ori)
        assert got.shape == ref.shape == (grid.ncells, R)
        assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max()
    # nodal interpolant of a quadratic in P2: zero error at the Lagrange points of order 2
    if ft.family == "H1" and ft.order == 2 and ft.ncomponents == 1 and dim == 3:
        pts = np.concatenate([grid["Coordinates"], 0.5 * (grid["Coordinates"][grid["EdgeNodes"][:, 0] - 1] + grid["Coordinates"][grid["EdgeNodes"][:, 1] - 1])])
        quad = lambda x: (1 + x[..., 0] - 2 * x[..., 1] * x[..., 2] + x[..., 0] ** 2)[..., None]  # noqa: E731
        err = fb.compute_error(fes, quad(pts)[:, 0], quad, order=2)
        assert np.sqrt(err.sum()) < 1e-14


def test_block_matmul_and_lrmatmul_match_scipy():
    """addblock_matmul! (fematrix.jl:486-563) per block / transposed / whole matrix, lrmatmul and ldrdmatmul (:607-637)"""
    import scipy.sparse as sp
    from femb_b200.assembly import _Session

    grid = _grid2d(7, True)
    FESu, FESp = fb.FESpace(fb.H1Pk(2, 2, 2), grid), fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    nu, npr = FESu.ndofs, FESp.ndofs
    n = nu + npr
    S = _Session([FESu, FESp], fb.QuadratureRule(grid.geometry, 5))
    ctx = S.ctx
    egu, eiu, eip = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity"), S.evaluator(1, "Identity")
    ctx.pattern_build([(0, 0), (0, 1), (1, 0)], [nu + 1])
    rng = np.random.default_rng(6)
    ctx.vector_set(1, rng.uniform(-1, 1, n))
    ctx.matrix_zero(); ctx.vector_zero()
    ctx.assemble_navierstokes(egu._eid, eiu._eid, eip._eid, 1e-2, True, True, _lib.RHS_NONE, None)
    colptr, rowval = ctx.pattern_get(n)
    A = sp.csc_matrix((ctx.matrix_get(), rowval - 1, colptr - 1), shape=(n, n)).tocsr()
    off = [0, nu, n]
    scale = abs(A).max()
    for br in (0, 1):
        for bc in (0, 1):
            B = A[off[br]:off[br + 1], off[bc]:off[bc + 1]]
            for transposed in (False, True):
                b = rng.uniform(-1, 1, B.shape[0] if transposed else B.shape[1])
                a0 = rng.uniform(-1, 1, B.shape[1] if transposed else B.shape[0])
                a = ctx.block_matmul(a0.copy(), b, br, bc, factor=-0.75, transposed=transposed)
                ref = a0 - 0.75 * ((B.T @ b) if transposed else (B @ b))
                assert np.abs(a - ref).max() <= 1e-12 * scale * 30
    b, a0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    a = ctx.block_matmul(a0.copy(), b, factor=2.0)
    assert np.abs(a - (a0 + 2.0 * (A @ b))).max() <= 1e-12 * scale * 60
    a1, a2, b1, b2 = (rng.uniform(-1, 1, n) for _ in range(4))
    assert abs(ctx.lrmatmul(a1, b1, factor=1.5) - 1.5 * (a1 @ (A @ b1))) <= 1e-12 * scale * n
    assert abs(ctx.lrmatmul(a1, b1, a2, b2) - (a1 - a2) @ (A @ (b1 - b2))) <= 1e-12 * scale * n
    assert ctx.lrmatmul(a1, b1, a2, b2) == ctx.lrmatmul(a1, b1, a2, b2)  # reproducible
    ctx.close()


def test_device_topology_hand_derived_answers():
    """The hand-derived tables of tests/test_grids.py (two triangles, two glued tetrahedra) straight from the device enumeration."""
    from test_grids import hand_grids

    for grid, hand in hand_grids():
        dev = fb.ExtendableGrid(grid["Coordinates"], grid["CellNodes"], grid.geometry).instantiate_on_device()
        for name, want in hand.items():
            if name != "cellnodes":
                assert np.array_equal(dev._c[name], want), name


def test_compute_error_known_answers():
    """uh = 1, u = nothing: error = |T| (VertexRule weights sum to 1, any p); P1 interpolant of a linear u: zero error;
    integral mean: sum_cells error(p = 1) = integral of uh (Example210:88)."""
    grid = _grid2d(5, True)
    fes = fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    one = np.ones(fes.ndofs)
    for p in (1, 2, 3):
        err = fb.compute_error(fes, one, None, order=2, p=p)
        assert np.abs(err[:, 0] - grid["CellVolumes"]).max() <= 1e-15
    x = grid["Coordinates"]
    lin = lambda y: (2.0 - 3.0 * y[..., 0] + 0.5 * y[..., 1])[..., None]  # noqa: E731
    assert fb.compute_error(fes, lin(x)[:, 0], lin, order=3).sum() <= 1e-28
    # integral of the P1 interpolant of x + y over the (jittered, but still unit) square = 1
    assert abs(fb.compute_error(fes, x[:, 0] + x[:, 1], None, order=1, p=1).sum() - 1.0) <= 1e-13


def test_device_topology_matches_the_oracle_loop():
    """device face / edge enumeration against the literal first-appearance loop of oracle/femb_oracle.py (bit-exact)"""
    for grid in (_shuffled(_grid2d(4), seed=21), _shuffled(_grid3d(3), seed=22)):
        dev = fb.ExtendableGrid(grid["Coordinates"], grid["CellNodes"], grid.geometry).instantiate_on_device()
        f = orc.enumerate_entities(grid.geometry, grid["CellNodes"], "faces")
        e = orc.enumerate_entities(grid.geometry, grid["CellNodes"], "edges")
        for name, want in (("CellFaces", f["cell_entities"]), ("FaceNodes", f["entity_nodes"]), ("CellFaceSigns", f["signs"]),
                           ("FaceCells", f["facecells"]), ("CellFaceOrientations", f["orientations"]), ("CellEdges", e["cell_entities"]),
                           ("EdgeNodes", e["entity_nodes"]), ("CellEdgeSigns", e["signs"])):
            assert np.array_equal(dev._c[name], want), name


# ---------------------------------------------------------------------------------------------
# sizes that reach the code paths only big meshes take (VERDICT r01 "what's weak" 1): several launch ranges, SPLIT-4
# vertex columns, first-touch accumulators, multi-CTA grids of the generic kernels — against the oracle, not identities
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,jit,device_dofmaps", [(64, False, True), (48, False, True), (24, True, False)])
def test_midsize_p2_3d_against_c_oracle(m, jit, device_dofmaps):
    """BASELINE.json configs[2] (H1P2{1,3} stiffness + rhs on the Kuhn mesh) at m = 48 (663,552 tets, ndofs 912,673): device
    result against the C restatement of the reference loop on the device's pattern; the pattern itself must take every
    reference contribution (miss == 0) and hold no slot the reference never writes."""
    import oracle_c
    from femb_b200.assembly import _Session

    grid = _grid3d(m, jit)
    if device_dofmaps:
        grid.instantiate_on_device(0, components=("edges",))
    fes = fb.FESpace(fb.H1P2(1, 3), grid)
    qf = fb.QuadratureRule(grid.geometry, 2)
    S = _Session([fes], qf, device_dofmaps=device_dofmaps)
    ctx = S.ctx
    eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
    nnz = ctx.pattern_build([(0, 0)])
    p = np.array([0.25, 1.0, -1.0, 0.5])
    mu = 1.3
    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg._eid, ei._eid, mu, _lib.RHS_AFFINE, p, 0.0)
    nz, b = ctx.matrix_get(), ctx.vector_get(fes.ndofs)
    colptr, rowval = ctx.pattern_get(fes.ndofs)
    cd = np.ascontiguousarray(ctx.space_get_celldofs(0, 10) if device_dofmaps else fes["CellDofs"])
    el = _oracle_el(fb.H1P2(1, 3))
    xref, w = orc.quadrature("Tetrahedron3D", 2)
    vals, der = el.tabulate("Tetrahedron3D", xref)
    nz2, b2 = np.zeros(nnz), np.zeros(fes.ndofs)
    miss = oracle_c.example200(grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], cd, w, xref, vals[:, :, 0], der[:, :, 0, :], mu, p, 0.0, colptr, rowval,
                               nz2, b2, oracle_c.max_threads())
    assert miss == 0
    scale = np.abs(nz2).max()
    assert np.abs(nz - nz2).max() <= RTOL * scale
    assert np.abs(b - b2).max() <= RTOL * np.abs(b2).max()
    # structural pattern = union of dofs x dofs: every slot gets at least one contribution; a second assembly adds up
    ctx.assemble_poisson(eg._eid, ei._eid, mu, _lib.RHS_AFFINE, p, 0.0)
    assert np.abs(ctx.matrix_get() - 2 * nz).max() <= 4 * RTOL * scale
    cols = np.repeat(np.arange(fes.ndofs, dtype=np.int64), np.diff(colptr))
    assert np.abs(np.bincount(cols, weights=nz, minlength=fes.ndofs)).max() <= 1e-12 * scale  # 1^T A = 0


@pytest.mark.parametrize("name,n,scatter", [("HDIVRT0", 32, None), ("HDIVBDM1", 24, None), ("HDIVRT0", 20, _lib.SCATTER_COLORED)])
def test_midsize_hdiv_3d_against_oracle(name, n, scatter):
    """BASELINE.json configs[3] at m = 32 / 24 (196,608 / 82,944 tets): mass + (div v, q) blocks against the numpy oracle."""
    from femb_b200.assembly import _Session

    grid = _grid3d(n, True)
    g = grid.geometry
    ft = {"HDIVRT0": fb.HDIVRT0, "HDIVBDM1": fb.HDIVBDM1}[name](3)
    FESv, FESq = fb.FESpace(ft, grid), fb.FESpace(fb.L2P0(1), grid)
    qf = fb.QuadratureRule(g, 2)
    S = _Session([FESv, FESq], qf)
    ctx = S.ctx
    eid, ediv, eq = S.evaluator(0, "Identity"), S.evaluator(0, "Divergence"), S.evaluator(1, "Identity")
    nnz = ctx.pattern_build([(0, 0), (0, 1), (1, 0)])
    if scatter is not None:
        ctx.set_scatter(scatter)
    ctx.matrix_zero()
    ctx.assemble_bilinear(eid._eid, eid._eid, 0, 0, 1.0, 0, 0.0)
    ctx.assemble_bilinear(ediv._eid, eq._eid, 0, 1, 1.0, _lib.BIL_TRANSPOSE_COPY, 0.0)
    ctx.sync()
    nz = ctx.matrix_get()
    nall = FESv.ndofs + FESq.ndofs
    cp, rv = ctx.pattern_get(nall)
    el, elq = _oracle_el(ft), orc.Element("L2P0", 1)
    xref, w = orc.quadrature(g, 2)
    args = (g, xref, grid["Coordinates"], grid["CellNodes"], grid["CellFaceSigns"], None, grid["CellFaceOrientations"])
    V = orc.update_basis(el, "Identity", *args)
    D = orc.update_basis(el, "Divergence", *args)
    Q = orc.update_basis(elq, "Identity", *args)
    vol = grid["CellVolumes"]
    M = orc.bilinear_local(V, V, w, vol)
    B = orc.bilinear_local(D, Q, w, vol)
    I1, J1, V1 = orc.block_triplets(M, FESv["CellDofs"], FESv["CellDofs"])
    I2, J2, V2 = orc.block_triplets(B, FESv["CellDofs"], FESq["CellDofs"], 0, FESv.ndofs)
    colptr, rowval, nzval = orc.flush_csc(nall, nall, np.concatenate([I1, I2, J2]), np.concatenate([J1, J2, I2]), np.concatenate([V1, V2, V2]))
    assert nnz == rowval.shape[0] and np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    assert np.abs(nz - nzval).max() <= RTOL * max(np.abs(M).max(), np.abs(B).max())


def test_midsize_example210_against_oracle():
    """BASELINE.json configs[4] at n = 256 (131,072 triangles, 592k dofs): update_system!(true, true) against the numpy oracle."""
    n = 256
    grid = _grid2d(n, True)
    FESu, FESp = fb.FESpace(fb.H1Pk(2, 2, 2), grid), fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    mu = 1.0e-2
    sol = fb.FEVector([FESu, FESp])
    sol.entries[:] = np.random.default_rng(5).uniform(-1, 1, sol.entries.shape[0])
    A, b = fb.FEMatrix([FESu, FESp]), fb.FEVector([FESu, FESp])
    upd = fb.prepare_assembly(A, b, FESu, FESp, sol, mu=mu)
    upd(True, True)
    coords = grid["Coordinates"]
    c2, r2, v2, b2 = orc.example210_assemble(coords, grid["CellNodes"], grid["CellVolumes"], FESu["CellDofs"], FESp["CellDofs"], FESu.ndofs, FESp.ndofs,
                                             sol.entries, mu, 0.0, True, True)
    import scipy.sparse as sp

    n_all = FESu.ndofs + FESp.ndofs
    want = sp.csc_matrix((v2, r2 - 1, c2 - 1), shape=(n_all,) * 2)
    E = A.entries
    D = (E.to_scipy() - want).tocoo()
    assert np.abs(D.data).max() <= RTOL * np.abs(v2).max()
    # pattern: the oracle's entries plus the pinned pressure diagonal, nothing else
    P = sp.csc_matrix((np.ones_like(v2), r2 - 1, c2 - 1), shape=(n_all,) * 2).tolil()
    P[FESu.ndofs, FESu.ndofs] = 1
    P = P.tocsc()
    P.sort_indices()
    assert np.array_equal(E.colptr - 1, P.indptr) and np.array_equal(E.rowval - 1, P.indices)
    assert np.abs(b.entries - b2).max() <= RTOL * np.abs(b2).max()


# ---------------------------------------------------------------------------------------------
# connectivity guard: equal counts are not equal topology
# ---------------------------------------------------------------------------------------------
def test_changed_connectivity_is_detected():
    from femb_b200.assembly import _Session

    grid = _grid2d(24, True)
    fes = fb.FESpace(fb.H1P1(1), grid)
    S = _Session([fes], fb.QuadratureRule(grid.geometry, 0))
    ctx = S.ctx
    eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
    nnz = ctx.pattern_build([(0, 0)])
    p = np.array([0.3, 1.0, -1.0])
    coords, cn, vol = grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"]
    nz, b = np.empty(nnz), np.empty(fes.ndofs)
    # unchanged connectivity, passed or declared (NULL): fine
    ctx.assemble_poisson_host(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 0.0, coords, cn, vol, nz, b, 4)
    nz0 = nz.copy()
    ctx.assemble_poisson_host(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 0.0, coords, None, vol, nz, b, 4)
    assert np.array_equal(nz, nz0)
    # an edge flip between the two triangles of one quad: same counts, different topology
    cn2 = cn.copy()
    quad = np.unique(cn2[:2])
    assert quad.shape[0] == 4
    a, c = np.intersect1d(cn2[0], cn2[1])  # the shared edge
    o0, o1 = np.setdiff1d(cn2[0], [a, c])[0], np.setdiff1d(cn2[1], [a, c])[0]
    cn2[0], cn2[1] = [a, o0, o1], [c, o1, o0]
    with pytest.raises(_lib.FembError) as ei_:
        ctx.assemble_poisson_host(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 0.0, coords, cn2, vol, nz, b, 4)
    assert ei_.value.status == 1 and "connectivity" in str(ei_.value)
    # femb_mesh_set with the same counts: identical connectivity keeps the symbolic phase ...
    ctx.mesh_set(coords, cn, vol)
    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 0.0)
    assert np.array_equal(ctx.matrix_get(), nz0)
    # ... a changed one drops spaces and pattern: assembling without registering them again is an error, not a stale matrix
    ctx.mesh_set(coords, cn2, vol)
    with pytest.raises(_lib.FembError):
        ctx.assemble_poisson(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 0.0)


# ---------------------------------------------------------------------------------------------
# interpolate!(u_init[1], u!) on the device (node values + edge-mean preserving edge dofs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,ncomp,kind", [(2, 2, "ex210"), (2, 1, "affine"), (3, 1, "affine"), (3, 3, "affine")])
def test_device_interpolation_matches_oracle(dim, ncomp, kind):
    from femb_b200.assembly import _Session
    from femb_b200.fespace import parse_pattern

    grid = _grid2d(9, True) if dim == 2 else _grid3d(4, True)
    ft = fb.H1Pk(ncomp, 2, 2) if dim == 2 else fb.H1P2(ncomp, 3)
    fes = fb.FESpace(ft, grid)
    S = _Session([fes], fb.QuadratureRule(grid.geometry, 2))
    ctx = S.ctx
    segs = parse_pattern(ft.get_dofmap_pattern(grid.geometry))
    edgenodes = grid["FaceNodes"] if dim == 2 else grid["EdgeNodes"]
    mu, t = 1.0e-2, 0.3
    if kind == "ex210":
        e = np.exp(-8 * np.pi**2 * mu * t)
        u = lambda x: np.array([e * np.sin(2 * np.pi * x[0]) * np.sin(2 * np.pi * x[1]), e * np.cos(2 * np.pi * x[0]) * np.cos(2 * np.pi * x[1])])  # noqa: E731
        fk, params = _lib.FUNC_EX210_U, np.array([mu, t])
    else:
        rng = np.random.default_rng(4)
        c0, g = rng.uniform(-1, 1, ncomp), rng.uniform(-1, 1, (ncomp, dim))
        u = lambda x: c0 + g @ x  # noqa: E731
        fk, params = _lib.FUNC_AFFINE, np.concatenate([np.concatenate([[c0[c]], g[c]]) for c in range(ncomp)])
    want = orc.interpolate_h1_nodal_edge(grid["Coordinates"], edgenodes, u, ncomp)
    qx, qw = orc.gauss_rule_edge(4)
    got = ctx.interpolate(0, segs, fes.ndofs, fk, params, qx, qw)
    assert np.abs(got - want).max() <= RTOL * np.abs(want).max()
    if kind == "affine":  # the P2 interpolant of an affine function is the function: edge dofs = midpoint values
        mid = ctx.interpolate(0, segs, fes.ndofs, fk, params)
        assert np.abs(mid - want).max() <= 1e-14
    # boundary dofs only: everything else keeps its value (zero), and dirichlet_apply(values=None) uses the device vector
    bd = ctx.boundary_dofs(0, segs)
    ctx.pattern_build([(0, 0)])
    S2 = _Session([fes], fb.QuadratureRule(grid.geometry, 2))
    c2 = S2.ctx
    c2.pattern_build([(0, 0)])
    c2.boundary_dofs(0, segs, download=False)
    only = c2.interpolate(0, segs, fes.ndofs, fk, params, qx, qw, only_boundary=True)
    ref = np.zeros_like(want)
    ref[bd - 1] = want[bd - 1]
    assert np.abs(only - ref).max() <= RTOL * np.abs(want).max()
    c2.matrix_zero()
    c2.vector_zero()
    c2.dirichlet_apply(True, 0, None, None, 1.0e60)
    b = c2.vector_get(fes.ndofs)
    assert np.abs(b * 1.0e-60 - ref).max() <= RTOL * max(np.abs(want).max(), 1.0)


@pytest.mark.parametrize("dim", [3, 2])
def test_p2_closed_form_and_quadrature_kernels_agree(dim):
    """H1P2{1,3} / H1P2{1,2}: the closed-form gather (femb_p2.cu; gather map in role order) and the general quadrature gather
    (femb_h1.cu; local dof order — taken when the 1e-15 value filter of Example200:150 is on) run on the SAME context one
    after the other; the map is re-ordered in place between them and both match the oracle."""
    from femb_b200.assembly import _Session

    grid = _grid3d(7, True) if dim == 3 else _grid2d(37, True)
    fes = fb.FESpace(fb.H1P2(1, dim), grid)
    qf = fb.QuadratureRule(grid.geometry, 2)
    S = _Session([fes], qf)
    ctx = S.ctx
    eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
    ctx.pattern_build([(0, 0)])
    p = np.array([0.25, 1.0, -1.0, 0.5][:dim + 1])
    rhs = fb.AffineFunction([0.25], [[1.0, -1.0, 0.5][:dim]])
    colptr, rowval, nzref, bref = orc.example200_assemble(_oracle_el(fb.H1P2(1, dim)), grid.geometry, grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"],
                                                           fes["CellDofs"], fes.ndofs, 0.7, rhs)
    cp, rv = ctx.pattern_get(fes.ndofs)
    ref = {(int(r), c): v for c in range(fes.ndofs) for r, v in zip(rowval[colptr[c] - 1:colptr[c + 1] - 1], nzref[colptr[c] - 1:colptr[c + 1] - 1])}
    scale = np.abs(nzref).max()
    results = []
    for drop_tol in (0.0, 1.0e-15, 0.0, -1.0):
        ctx.matrix_zero()
        ctx.vector_zero()
        ctx.assemble_poisson(eg._eid, ei._eid, 0.7, _lib.RHS_AFFINE, p, drop_tol)
        nz, b = ctx.matrix_get(), ctx.vector_get(fes.ndofs)
        results.append(nz)
        assert np.abs(b - bref).max() <= RTOL * np.abs(bref).max()
        for c in range(0, fes.ndofs, 3):
            for r, v in zip(rv[cp[c] - 1:cp[c + 1] - 1], nz[cp[c] - 1:cp[c + 1] - 1]):
                assert abs(v - ref.get((int(r), c), 0.0)) <= RTOL * scale
    assert np.array_equal(results[0], results[2])  # bit-reproducible across the re-ordering of the map
    assert np.abs(results[0] - results[1]).max() <= RTOL * scale


@pytest.mark.parametrize("m,nchunks,use_vol", [(9, 5, True), (14, 12, False)])
def test_streamed_host_assembly_p2_tets_is_bit_identical(m, nchunks, use_vol):
    """femb_assemble_poisson_host for scalar P2 on tetrahedra (closed-form gather; geometry uploaded, finished nzval / b ranges
    downloaded behind the kernels) == mesh_set + assemble + get, bit for bit, and it follows the uploaded geometry."""
    from femb_b200.assembly import _Session

    grid = _grid3d(m, True)
    fes = fb.FESpace(fb.H1P2(1, 3), grid)
    qf = fb.QuadratureRule(grid.geometry, 2)
    S = _Session([fes], qf, cellvolumes=use_vol)
    ctx = S.ctx
    eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
    nnz = ctx.pattern_build([(0, 0)])
    p = np.array([0.3, 1.0, -1.0, 0.5])
    coords2 = np.ascontiguousarray(grid["Coordinates"] * 1.25 + 0.1)
    vol2 = np.ascontiguousarray(grid["CellVolumes"] * 1.25 ** 3) if use_vol else None
    nz2, b2 = np.full(nnz, np.nan), np.full(fes.ndofs, np.nan)
    ctx.assemble_poisson_host(eg._eid, ei._eid, 1.7, _lib.RHS_AFFINE, p, 0.0, coords2, grid["CellNodes"], vol2, nz2, b2, nchunks)
    assert np.isfinite(nz2).all() and np.isfinite(b2).all()  # every range came back
    ctx.mesh_set(coords2, grid["CellNodes"], vol2)
    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg._eid, ei._eid, 1.7, _lib.RHS_AFFINE, p, 0.0)
    nz3, b3 = ctx.matrix_get(), ctx.vector_get(fes.ndofs)
    assert np.array_equal(nz2, nz3) and np.array_equal(b2, b3)
    # and against the oracle on the new geometry
    rhs = fb.AffineFunction([0.3], [[1.0, -1.0, 0.5]])
    colptr, rowval, nzref, bref = orc.example200_assemble(_oracle_el(fb.H1P2(1, 3)), grid.geometry, coords2, grid["CellNodes"],
                                                           vol2 if use_vol else np.ascontiguousarray(grid["CellVolumes"] * 1.25 ** 3), fes["CellDofs"], fes.ndofs, 1.7, rhs)
    assert np.abs(b2 - bref).max() <= RTOL * np.abs(bref).max()


def _mirrored3d(n, every=3):
    """jittered Kuhn mesh with every `every`-th cell negatively oriented (two nodes swapped: det A < 0)"""
    grid = _grid3d(n, True)
    cn = grid["CellNodes"].copy()
    cn[1::every] = cn[1::every][:, [0, 1, 3, 2]]
    return fb.ExtendableGrid(grid["Coordinates"], cn, "Tetrahedron3D")


def test_closed_form_kernels_on_mirrored_cells():
    """negatively oriented cells and ragged sizes (cell / dof counts that are no multiples of 128 / 32) through the closed-form
    paths: P2 on tetrahedra (G_ab is orientation-independent, |T| comes from CellVolumes); tiny RT0 meshes"""
    from femb_b200.assembly import _Session

    for n in (1, 2, 5):
        grid = _mirrored3d(n)
        fes = fb.FESpace(fb.H1P2(1, 3), grid)
        qf = fb.QuadratureRule(grid.geometry, 2)
        S = _Session([fes], qf)
        ctx = S.ctx
        eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
        ctx.pattern_build([(0, 0)])
        rhs = fb.AffineFunction([0.25], [[1.0, -1.0, 0.5]])
        colptr, rowval, nzref, bref = orc.example200_assemble(_oracle_el(fb.H1P2(1, 3)), grid.geometry, grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"],
                                                               fes["CellDofs"], fes.ndofs, 0.7, rhs)
        ctx.matrix_zero()
        ctx.vector_zero()
        ctx.assemble_poisson(eg._eid, ei._eid, 0.7, _lib.RHS_AFFINE, np.array([0.25, 1.0, -1.0, 0.5]), 0.0)
        nz, b = ctx.matrix_get(), ctx.vector_get(fes.ndofs)
        cp, rv = ctx.pattern_get(fes.ndofs)
        import scipy.sparse as sp

        got = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(fes.ndofs, fes.ndofs))
        want = sp.csc_matrix((nzref, rowval - 1, colptr - 1), shape=(fes.ndofs, fes.ndofs))
        assert abs(got - want).max() <= RTOL * np.abs(nzref).max()
        assert np.abs(b - bref).max() <= RTOL * np.abs(bref).max()
    # (Hdiv spaces on mirrored cells are not a case: the face orientations of ExtendableGrids assume consistently oriented cells)
    for n in (1, 2):  # ragged sizes through the RT0 closed form: 6 / 48 cells
        _check_hdiv_mixed(_grid3d(n, True), fb.HDIVRT0(3))
