"""Pins of the CPU oracle (and of the host-side tables the CUDA kernels consume) against the
reference's own known-answer tests for the assembly path:

  (1) nodal Kronecker property of every basis        /root/reference/test/test_febasis.jl:24-46
  (2) quadrature exactness per order                 /root/reference/test/test_quadrature.jl:16-54
  (3) P2 derivative constants on a non-trivial cell  /root/reference/test/test_operators.jl:67-143,192-271
  (4) interpolation -> evaluation round trip         /root/reference/test/test_interpolators.jl:43-115
  (5) FEMatrix / FEVector block offsets              /root/reference/test/test_fematrix_and_vector.jl

No GPU needed.
"""
from fractions import Fraction as Fr
from math import factorial

import numpy as np
import pytest

import femb_oracle as orc
import femb_b200 as fb
from femb_b200 import fedefs
from femb_b200.fespace import init_celldofs_from_pattern

TOL = 6.0e-12  # test/runtests.jl:66


# ---------------------------------------------------------------------------------------------
def vertex_rule(geometry, order):
    """VertexRule(EG, order; T = Rational) — /root/reference/src/quadrature.jl:91-204"""
    if geometry == "Triangle2D":
        x = [(Fr(0), Fr(0)), (Fr(1), Fr(0)), (Fr(0), Fr(1))]
        edges = [(0, 1), (1, 2), (2, 0)]
        for a, b in edges:
            for j in range(1, order):
                t = Fr(j, order)
                x.append(tuple(t * x[b][i] + (1 - t) * x[a][i] for i in range(2)))
        if order == 3:
            x.append((Fr(1, 3), Fr(1, 3)))
        if order == 4:
            x += [(Fr(1, 4), Fr(1, 4)), (Fr(1, 2), Fr(1, 4)), (Fr(1, 4), Fr(1, 2))]
        if order == 5:
            x += [(Fr(1, 5), Fr(1, 5)), (Fr(3, 5), Fr(1, 5)), (Fr(1, 5), Fr(3, 5)), (Fr(2, 5), Fr(1, 5)), (Fr(2, 5), Fr(2, 5)), (Fr(1, 5), Fr(2, 5))]
        return x
    x = [(Fr(0),) * 3, (Fr(1), Fr(0), Fr(0)), (Fr(0), Fr(1), Fr(0)), (Fr(0), Fr(0), Fr(1))]
    edges = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    for a, b in edges:
        for j in range(1, order):
            t = Fr(j, order)
            x.append(tuple(t * x[b][i] + (1 - t) * x[a][i] for i in range(3)))
    if order > 2:
        faces = [(0, 2, 1), (0, 1, 3), (1, 2, 3), (0, 3, 2)]
        for j in range(1, order - 1):
            for k in range(1, order - 1):
                for f in faces:
                    tj, tk = Fr(j, order), Fr(k, order)
                    x.append(tuple(tj * x[f[2]][i] + tk * x[f[1]][i] + (1 - tj - tk) * x[f[0]][i] for i in range(3)))
    return x


KRONECKER_CASES = [("Triangle2D", "H1Pk", o) for o in range(1, 6)] + [
    ("Triangle2D", "H1P1", 1), ("Triangle2D", "H1P2", 2), ("Triangle2D", "H1P3", 3),
    ("Tetrahedron3D", "H1P1", 1), ("Tetrahedron3D", "H1P2", 2), ("Tetrahedron3D", "H1P3", 3),
]


def _product_fetype(name, order, ncomp=1, dim=2):
    if name == "H1Pk":
        return fb.H1Pk(ncomp, dim, order)
    if name == "H1P1":
        return fb.H1P1(ncomp)
    return {"H1P2": fb.H1P2, "H1P3": fb.H1P3}[name](ncomp, dim)


@pytest.mark.parametrize("geometry,name,order", KRONECKER_CASES)
def test_vertex_values(geometry, name, order):
    """(1): basis_j(vertex-rule point_i) = delta_ij — exactly (rationals) for the product's polynomial
    tables, to rounding for the oracle's float closures."""
    pts = vertex_rule(geometry, order)
    d = len(pts[0])
    ft = _product_fetype(name, order, 1, d)
    sb = ft.scalar_basis(geometry)
    assert len(sb) == len(pts)
    for i, p in enumerate(pts):
        for j, poly in enumerate(sb):
            assert poly(p) == (1 if i == j else 0), (name, i, j)
    el = orc.Element(name, 1, order)
    vals = el.basis(geometry, np.array([[float(c) for c in p] for p in pts]))[:, :, 0]
    assert np.abs(vals - np.eye(len(pts))).max() < 1e-13


@pytest.mark.parametrize("geometry,name,order,ncomp", [
    ("Triangle2D", "H1P1", 1, 1), ("Triangle2D", "H1P2", 2, 2), ("Triangle2D", "H1P3", 3, 1), ("Triangle2D", "H1Pk", 2, 2),
    ("Triangle2D", "H1Pk", 3, 1), ("Triangle2D", "H1Pk", 4, 1), ("Tetrahedron3D", "H1P1", 1, 1), ("Tetrahedron3D", "H1P2", 2, 1),
    ("Tetrahedron3D", "H1P3", 3, 1), ("Triangle2D", "HDIVRT0", 1, 2), ("Tetrahedron3D", "HDIVRT0", 1, 3),
    ("Triangle2D", "HDIVBDM1", 1, 2), ("Tetrahedron3D", "HDIVBDM1", 1, 3), ("Triangle2D", "L2P0", 0, 1), ("Tetrahedron3D", "L2P0", 0, 1),
])
def test_tables_product_vs_oracle(geometry, name, order, ncomp):
    """The tables handed to the CUDA library (closed-form polynomial derivatives) equal the oracle's
    (complex-step, the ForwardDiff stand-in) to rounding."""
    d = orc.DIM[geometry]
    rng = np.random.default_rng(7)
    x = rng.dirichlet(np.ones(d + 1), size=11)[:, :d]
    if name.startswith("HDIV"):
        ft = {"HDIVRT0": fb.HDIVRT0, "HDIVBDM1": fb.HDIVBDM1}[name](d)
    elif name == "L2P0":
        ft = fb.L2P0(ncomp)
    else:
        ft = _product_fetype(name, order, ncomp, d)
    el = orc.Element(name, ncomp, order)
    v1, d1 = ft.tabulate(geometry, x)
    v2, d2 = el.tabulate(geometry, x)
    assert v1.shape == v2.shape and d1.shape == d2.shape
    assert ft.get_ndofs(geometry) == el.nd(geometry) and ft.get_ndofs_all(geometry) == el.nd_all(geometry)
    assert ft.get_dofmap_pattern(geometry) == el.pattern(geometry)
    scale = max(1.0, np.abs(d2).max())
    assert np.abs(v1 - v2).max() < 1e-13 * scale and np.abs(d1 - d2).max() < 1e-12 * scale


# ---------------------------------------------------------------------------------------------
def _monomial_integral(exps):
    """int over the reference simplex of prod x_i^e_i, divided by the simplex volume (weights sum to 1)."""
    d = len(exps)
    num = 1
    for e in exps:
        num *= factorial(e)
    return Fr(num * factorial(d), factorial(sum(exps) + d))


@pytest.mark.parametrize("order", [0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 15, 20])
def test_quadrature_triangle(order):
    """(2) test/test_quadrature.jl:24-35: the order-p rule integrates all monomials of degree <= p."""
    for mod in (orc.quadrature, lambda g, o: (fb.QuadratureRule(g, o).xref, fb.QuadratureRule(g, o).w)):
        x, w = mod("Triangle2D", order)
        assert abs(w.sum() - 1) < 1e-14
        for a in range(order + 1):
            for b in range(order + 1 - a):
                assert abs(np.sum(w * x[:, 0] ** a * x[:, 1] ** b) - float(_monomial_integral((a, b)))) < 2e-14, (order, a, b)
    x1, w1 = orc.quadrature("Triangle2D", order)
    q = fb.QuadratureRule("Triangle2D", order)
    # same point ORDER too (the kernels see the product's table, the oracle its own)
    assert np.abs(x1 - q.xref).max() < 1e-14 and np.abs(w1 - q.w).max() < 1e-14


@pytest.mark.parametrize("order", [0, 1, 2, 3, 4])
def test_quadrature_tet(order):
    """(2) test/test_quadrature.jl:46-54 for the tet rules on the path."""
    x, w = orc.quadrature("Tetrahedron3D", order)
    q = fb.QuadratureRule("Tetrahedron3D", order)
    assert np.abs(x - q.xref).max() == 0 and np.abs(w - q.w).max() == 0
    for a in range(order + 1):
        for b in range(order + 1 - a):
            for c in range(order + 1 - a - b):
                assert abs(np.sum(w * x[:, 0] ** a * x[:, 1] ** b * x[:, 2] ** c) - float(_monomial_integral((a, b, c)))) < 2e-14


def test_stroud_order5_is_example210_rule():
    x, w = orc.quadrature("Triangle2D", 5)  # Example210:237 -> 9 points
    assert x.shape == (9, 2)
    # point order (s_j outer, r_i inner): quadrature.jl:656-662
    assert np.all(np.diff(x[:3, 1]) > 0) and abs(x[0, 0] - x[1, 0]) < 1e-15 and x[3, 0] > x[0, 0]


# ---------------------------------------------------------------------------------------------
def _nodal_points_2d(order):
    return np.array([[float(c) for c in p] for p in vertex_rule("Triangle2D", order)])


def test_p2_gradient_golden_triangle():
    """(3) test/test_operators.jl:67-143: H1P2{2,2} on the triangle (-1,0),(1,0),(0,1), interpolant of
    u = (x^2, 3y^2 + xy).  The reference asserts Curl2D = 1/3 at the midpoint to 1e-14; Curl2D is
    grad[3] - grad[2] of the Gradient evaluator (rows k + edim*(c-1)), which pins A^{-T} and the row order."""
    coords = np.array([[-1.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    cn = np.array([[1, 2, 3]], dtype=np.int32)
    el = orc.Element("H1P2", 2)
    xref, _ = orc.quadrature("Triangle2D", 0)
    G = orc.update_basis(el, "Gradient", "Triangle2D", xref, coords, cn)[0, :, :, 0]  # [4, 12]
    A, b, det, M = orc.l2g(coords, cn)
    assert np.allclose(A[0], [[2, 1], [0, 1]]) and det[0] == 2.0 and np.allclose(M[0], [[0.5, 0], [-0.5, 1]])
    pts = orc.eval_trafo(A, b, _nodal_points_2d(2))[0]  # physical P2 nodes
    u = np.concatenate([pts[:, 0] ** 2, 3 * pts[:, 1] ** 2 + pts[:, 0] * pts[:, 1]])
    g = G @ u
    xm = np.array([0.0, 1.0 / 3.0])
    expected = np.array([2 * xm[0], 0.0, xm[1], 6 * xm[1] + xm[0]])
    assert np.abs(g - expected).max() < 1e-14
    assert abs((g[2] - g[1]) - 1.0 / 3.0) < 1e-14  # expected_curl2


def test_p2_gradient_golden_tet():
    """(3) test/test_operators.jl:192-271: H1P2{3,3} on the reference tet stretched to x2 = (2,0,0);
    expected_curl3 = [1/2 - 6/4, 0, 1/4 - 1/4]."""
    coords = np.array([[0.0, 0, 0], [2.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]])
    cn = np.array([[1, 2, 3, 4]], dtype=np.int32)
    el = orc.Element("H1P2", 3)
    xref, _ = orc.quadrature("Tetrahedron3D", 0)
    G = orc.update_basis(el, "Gradient", "Tetrahedron3D", xref, coords, cn)[0, :, :, 0]  # [9, 30]
    A, b, _, _ = orc.l2g(coords, cn)
    nodal = np.array([[float(c) for c in p] for p in vertex_rule("Tetrahedron3D", 2)])
    p = orc.eval_trafo(A, b, nodal)[0]
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    u = np.concatenate([x**2 + z * y, 3 * z**2 + x * y, x * y])
    g = G @ u  # (d1u1,d2u1,d3u1, d1u2,d2u2,d3u2, d1u3,d2u3,d3u3)
    curl = np.array([g[7] - g[5], g[2] - g[6], g[3] - g[1]])
    assert np.abs(curl - np.array([0.5 - 1.5, 0.0, 0.0])).max() < 1e-14


# ---------------------------------------------------------------------------------------------
def _grids():
    X = np.linspace(0, 1, 4)
    g2 = fb.jitter(fb.simplexgrid(X, X), 0.2)
    g3 = fb.jitter(fb.simplexgrid(np.linspace(0, 1, 3), np.linspace(0, 1, 3), np.linspace(0, 1, 3)), 0.15)
    return g2, g3


def _poly(dim, order, ncomp):
    """degree-`order` polynomial with `ncomp` components (the role of exact_function, test/runtests.jl:68-144)"""
    rng = np.random.default_rng(100 * dim + order)
    exps = [e for e in np.ndindex(*([order + 1] * dim)) if sum(e) <= order]
    coef = rng.uniform(-1, 1, size=(ncomp, len(exps)))

    def u(x):
        out = np.zeros(x.shape[:-1] + (ncomp,))
        for c in range(ncomp):
            for co, e in zip(coef[c], exps):
                t = co
                for i in range(dim):
                    t = t * x[..., i] ** e[i]
                out[..., c] += t
        return out

    return u


def _interpolate_h1(grid, el, celldofs, ndofs, u):
    """nodal interpolation (src/interpolators.jl:15-128): vertices, then points j/order along the GLOBAL
    face/edge direction, face centroids (P3 tets), cell-interior Lagrange points."""
    g, d, o = grid.geometry, grid.dim, el.order
    coords = grid["Coordinates"]
    vals = np.zeros(ndofs)
    nds = el.nd(g) // el.ncomp
    nn = grid.nnodes
    coffset = ndofs // el.ncomp
    uv = u(coords)
    for c in range(el.ncomp):
        vals[c * coffset: c * coffset + nn] = uv[:, c]
    off = nn
    if o >= 2:
        en = grid["FaceNodes"] if d == 2 else grid["EdgeNodes"]
        ne = en.shape[0]
        x1, x2 = coords[en[:, 0] - 1], coords[en[:, 1] - 1]
        for m in range(1, o):
            uv = u(x1 + (m / o) * (x2 - x1))
            for c in range(el.ncomp):
                vals[c * coffset + off + (m - 1) * ne: c * coffset + off + m * ne] = uv[:, c]
        off += ne * (o - 1)
    if o == 3 and d == 3:
        fn = grid["FaceNodes"]
        uv = u(coords[fn - 1].mean(axis=1))
        for c in range(el.ncomp):
            vals[c * coffset + off: c * coffset + off + fn.shape[0]] = uv[:, c]
    if o >= 3 and d == 2:
        ninter = (o - 2) * (o - 1) // 2
        pts = _nodal_points_2d(o)[3 * o:]
        A, b, _, _ = orc.l2g(coords, grid["CellNodes"])
        xp = orc.eval_trafo(A, b, pts)  # [cell, ninter, 2]
        uv = u(xp)
        nc = grid.ncells
        for c in range(el.ncomp):
            for m in range(ninter):
                vals[c * coffset + off + m * nc: c * coffset + off + (m + 1) * nc] = uv[:, m, c]
    return vals


H1_CASES = [("H1P1", 1, 2, 2), ("H1P2", 2, 2, 2), ("H1P3", 3, 2, 2), ("H1Pk", 2, 2, 2), ("H1Pk", 3, 2, 2), ("H1Pk", 4, 1, 2), ("H1Pk", 5, 1, 2),
            ("H1P1", 1, 3, 3), ("H1P2", 2, 3, 3), ("H1P3", 3, 1, 3)]


@pytest.mark.parametrize("name,order,ncomp,dim", H1_CASES)
def test_interpolation_roundtrip_h1(name, order, ncomp, dim):
    """(4) test/test_interpolators.jl:60-115 for the H1 elements on the path: interpolate a degree-p
    polynomial, evaluate through Identity cvals + CellDofs at the vertex rule of order p+1."""
    grid = _grids()[dim - 2]
    g = grid.geometry
    el = orc.Element(name, ncomp, order)
    ft = _product_fetype(name, order, ncomp, dim)
    fes = fb.FESpace(ft, grid)
    celldofs = fes["CellDofs"]
    # CellDofs: vectorised product generator == plain-loop restatement of dofmaps.jl:419-627
    cd2, ndofs2 = orc.celldofs_from_pattern(el, g, grid.nnodes, grid["CellNodes"], grid["CellFaces"], grid.nfaces,
                                            grid["CellEdges"] if dim == 3 else None, grid.nedges if dim == 3 else 0)
    assert ndofs2 == fes.ndofs and np.array_equal(cd2, celldofs)
    u = _poly(dim, order, ncomp)
    sol = _interpolate_h1(grid, el, celldofs, fes.ndofs, u)
    xq = np.array([[float(c) for c in p] for p in vertex_rule(g, min(order + 1, 3 if dim == 3 else 5))])
    cv = orc.update_basis(el, "Identity", g, xq, grid["Coordinates"], grid["CellNodes"], grid["CellFaceSigns"],
                          grid["CellEdgeSigns"] if dim == 3 else None)  # [cell, comp, dof, qp]
    uh = np.einsum("crdq,cd->cqr", cv, sol[celldofs.astype(np.int64) - 1])
    A, b, _, _ = orc.l2g(grid["Coordinates"], grid["CellNodes"])
    assert np.abs(uh - u(orc.eval_trafo(A, b, xq))).max() < TOL


@pytest.mark.parametrize("name,order,dim", [("HDIVRT0", 0, 2), ("HDIVBDM1", 1, 2), ("HDIVRT0", 0, 3), ("HDIVBDM1", 1, 3)])
def test_interpolation_roundtrip_hdiv(name, order, dim):
    """(4) test/test_interpolators.jl:43-57: face-moment interpolation of a degree-p field, then evaluation
    through the contravariant Piola map with CellFaceSigns / CellFaceOrientations handlers.  For
    HDIVBDM1{3} this is the only check that the synthesised orientation codes (grids.py) are consistent
    with the subset table hdiv_bdm1.jl:234-249."""
    grid = _grids()[dim - 2]
    g = grid.geometry
    el = orc.Element(name, dim)
    ft = {"HDIVRT0": fb.HDIVRT0, "HDIVBDM1": fb.HDIVBDM1}[name](dim)
    fes = fb.FESpace(ft, grid)
    celldofs = fes["CellDofs"]
    cd2, ndofs2 = orc.celldofs_from_pattern(el, g, grid.nnodes, grid["CellNodes"], grid["CellFaces"], grid.nfaces)
    assert ndofs2 == fes.ndofs and np.array_equal(cd2, celldofs)
    u = _poly(dim, order, dim)
    sol = orc.interpolate_hdiv(el, g, grid["Coordinates"], grid["FaceNodes"], u)
    xq, _ = orc.quadrature(g, 2)
    cv = orc.update_basis(el, "Identity", g, xq, grid["Coordinates"], grid["CellNodes"], grid["CellFaceSigns"], None,
                          grid["CellFaceOrientations"])
    uh = np.einsum("crdq,cd->cqr", cv, sol[celldofs.astype(np.int64) - 1])
    A, b, _, _ = orc.l2g(grid["Coordinates"], grid["CellNodes"])
    assert np.abs(uh - u(orc.eval_trafo(A, b, xq))).max() < TOL
    # divergence of the interpolant = divergence of u (constant per cell for p <= 1)
    dv = orc.update_basis(el, "Divergence", g, xq, grid["Coordinates"], grid["CellNodes"], grid["CellFaceSigns"], None,
                          grid["CellFaceOrientations"])
    divh = np.einsum("crdq,cd->cq", dv, sol[celldofs.astype(np.int64) - 1])
    h = 1e-6
    x0 = orc.eval_trafo(A, b, xq)
    div = sum((u(x0 + h * np.eye(dim)[i])[..., i] - u(x0 - h * np.eye(dim)[i])[..., i]) / (2 * h) for i in range(dim))
    assert np.abs(divh - div).max() < 1e-8


@pytest.mark.parametrize("dim", [2, 3])
def test_interpolation_roundtrip_hcurl(dim):
    """(4) test/test_interpolators.jl:15,44 (HCURLN0{2} / HCURLN0{3} => order 0): tangential-flux interpolation of a constant field, then
    evaluation through the covariant Piola map (feevaluator_hcurl.jl:2-16) with the CellFaceSigns / CellEdgeSigns coefficients
    (hcurl_n0.jl:144-166) reproduces the field; so does a rigid rotation a + w x (x), the rest of the first-kind Nedelec space."""
    grid = _grids()[dim - 2]
    g = grid.geometry
    el = orc.Element("HCURLN0", dim)
    ft = fb.HCURLN0(dim)
    fes = fb.FESpace(ft, grid)
    celldofs = fes["CellDofs"]
    cd2, ndofs2 = orc.celldofs_from_pattern(el, g, grid.nnodes, grid["CellNodes"], grid["CellFaces"], grid.nfaces,
                                            grid["CellEdges"] if dim == 3 else None, grid.nedges if dim == 3 else 0)
    assert ndofs2 == fes.ndofs == (grid.nfaces if dim == 2 else grid.nedges) and np.array_equal(cd2, celldofs)
    # product tables == oracle tables
    xq, _ = orc.quadrature(g, 2)
    v1, d1 = ft.tabulate(g, xq)
    v2, d2 = el.tabulate(g, xq)
    assert np.abs(v1 - v2).max() < 1e-15 and np.abs(d1 - d2).max() < 1e-15
    enodes = grid["FaceNodes"] if dim == 2 else grid["EdgeNodes"]
    A, b, _, _ = orc.l2g(grid["Coordinates"], grid["CellNodes"])
    x = orc.eval_trafo(A, b, xq)
    cv = orc.update_basis(el, "Identity", g, xq, grid["Coordinates"], grid["CellNodes"], grid["CellFaceSigns"],
                          grid["CellEdgeSigns"] if dim == 3 else None)
    const = _poly(dim, 0, dim)

    def rot(x):
        if dim == 2:
            return np.stack([0.3 - 0.7 * x[..., 1], -0.2 + 0.7 * x[..., 0]], axis=-1)
        w = np.array([0.4, -0.9, 0.6])
        return np.array([0.1, 0.2, -0.3]) + np.cross(np.broadcast_to(w, x.shape), x)

    for u in (const, rot):
        sol = orc.interpolate_hcurl(grid["Coordinates"], enodes, u)
        uh = np.einsum("crdq,cd->cqr", cv, sol[celldofs.astype(np.int64) - 1])
        assert np.abs(uh - u(x)).max() < TOL


# ---------------------------------------------------------------------------------------------
def test_block_offsets():
    """(5) test/test_fematrix_and_vector.jl: block (i,j) starts at row sum(ndofs X[<i]) / col sum(ndofs Y[<j])."""
    grid = _grids()[0]
    f1, f2 = fb.FESpace(fb.H1Pk(2, 2, 2), grid), fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    A = fb.FEMatrix([f1, f2])
    assert A.shape == (f1.ndofs + f2.ndofs,) * 2
    assert (A[1, 2].offset, A[1, 2].offsetY) == (0, f1.ndofs) and (A[2, 1].offset, A[2, 1].offsetY) == (f1.ndofs, 0)
    assert A[2, 2].last_index == f1.ndofs + f2.ndofs
    b = fb.FEVector([f1, f2])
    assert b[2].offset == f1.ndofs and len(b.entries) == f1.ndofs + f2.ndofs


# ---------------------------------------------------------------------------------------------
def test_analytic_identities_example200():
    """Analytic substitutes for the unpinned assembled values."""
    X = np.linspace(0, 1, 9)
    grid = fb.simplexgrid(X, X)
    el = orc.Element("H1Pk", 1, 1)
    cd = grid["CellNodes"]
    rhs = lambda x: x[..., 0] - x[..., 1]  # noqa: E731  Example200:38
    colptr, rowval, nzval, b = orc.example200_assemble(el, "Triangle2D", grid["Coordinates"], cd, grid["CellVolumes"], cd, grid.nnodes, 1.0, rhs)
    import scipy.sparse as sp

    A = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(grid.nnodes,) * 2)
    interior = 9 * 4 + 4  # node (ix=5, iy=5), 1-based 41
    row = A[interior].toarray().ravel()
    assert row[interior] == 4.0 and sorted(row[row != 0]) == [-1.0, -1.0, -1.0, -1.0, 4.0]  # 5-point stencil, Appendix C.2
    assert A.getrow(interior).nnz == 5  # hypotenuse couplings are exact zeros -> dropped by the 1e-15 filter
    assert abs(A @ np.ones(grid.nnodes)).max() < 1e-13 and abs(A - A.T).max() == 0
    assert abs(b.sum() - 0.0) < 1e-14  # int (x - y) over the unit square
    # u^T A u = int |grad u|^2 for u = 2x + 3y
    u = 2 * grid["Coordinates"][:, 0] + 3 * grid["Coordinates"][:, 1]
    assert abs(u @ (A @ u) - 13.0) < 1e-12


def test_analytic_identities_p2_local():
    """Appendix C.3 local P2 matrix on the right-isosceles cell."""
    h = 0.125
    coords = np.array([[0, 0], [h, 0], [h, h]], dtype=np.float64)
    cn = np.array([[1, 2, 3]], dtype=np.int32)
    Aloc, _ = orc.example200_local(orc.Element("H1Pk", 1, 2), "Triangle2D", coords, cn, orc.cell_volumes(coords, cn), 1.0, None)
    U = Aloc[0] + np.triu(Aloc[0], 1).T
    ref = np.array([[1 / 2, 1 / 6, 0, -2 / 3, 0, 0], [1 / 6, 1, 1 / 6, -2 / 3, -2 / 3, 0], [0, 1 / 6, 1 / 2, 0, -2 / 3, 0],
                    [-2 / 3, -2 / 3, 0, 8 / 3, 0, -4 / 3], [0, -2 / 3, -2 / 3, 0, 8 / 3, -4 / 3], [0, 0, 0, -4 / 3, -4 / 3, 8 / 3]])
    assert np.abs(U - ref).max() < 1e-14


def test_bdm1_orientation_codes():
    """The CellFaceOrientations codes 1..4 that grids.py synthesises are the unique ones for which the
    HDIVBDM1{3} subset (hdiv_bdm1.jl:234-249) + coefficient (:215-228) tables are dual to the face
    functionals int_F v.n, int_F v.n (xref_1 - 1/3), int_F v.n (xref_2 - 1/3) (hdiv_bdm1.jl:47-54) of the
    GLOBAL face — checked on the reference tet for every local face and every admissible node order."""
    from femb_b200.grids import LOCAL_CELLFACENODES, _face_orientation

    el = orc.Element("HDIVBDM1", 3)
    verts = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    fx, fw = orc._stroud(4)
    for j, lf in enumerate(LOCAL_CELLFACENODES["Tetrahedron3D"] - 1):
        for gn in [(lf[0], lf[1], lf[2]), (lf[0], lf[2], lf[1]), (lf[2], lf[1], lf[0]), (lf[1], lf[0], lf[2])]:
            code = int(_face_orientation(np.array([lf]), np.array([gn]))[0])
            sign = 1.0 if code == 1 else -1.0
            orient = np.ones((1, 4), dtype=np.int32)
            orient[0, j] = code
            sub = orc.basis_subset(el, "Tetrahedron3D", 1, orient=orient)[0, 3 * j:3 * j + 3]
            x = verts[list(gn)]
            n = np.cross(x[1] - x[0], x[2] - x[0])
            area = 0.5 * np.linalg.norm(n)
            n = n / np.linalg.norm(n)
            pts = x[0] + fx[:, [0]] * (x[1] - x[0]) + fx[:, [1]] * (x[2] - x[0])
            flux = el.basis("Tetrahedron3D", pts)[:, sub, :] @ n  # [q, 3]
            N = area * np.stack([fw @ flux, (fw * (fx[:, 0] - 1 / 3)) @ flux, (fw * (fx[:, 1] - 1 / 3)) @ flux])
            assert np.allclose(N @ np.diag([sign, -1.0, 1.0]), np.eye(3), atol=1e-12), (j, gn, code)


def test_vertex_rule_points_and_host_mirror_agree():
    """VertexRule (quadrature.jl:103-186): point counts = Lagrange nodes of the order, weights sum to 1, the vertices come first;
    the host mirror builds the same table in exact rationals."""
    import femb_b200 as fb

    for g, d, counts in (("Triangle2D", 2, {0: 1, 1: 3, 2: 6, 3: 10, 4: 15, 5: 21}), ("Tetrahedron3D", 3, {0: 1, 1: 4, 2: 10, 3: 20})):
        for order, n in counts.items():
            xref, w = orc.vertex_rule(g, order)
            assert xref.shape == (n, d) and abs(w.sum() - 1.0) < 1e-15
            assert len({tuple(np.round(p, 12)) for p in xref}) == n  # distinct points
            assert (xref >= -1e-15).all() and (xref.sum(axis=1) <= 1 + 1e-15).all()
            qf = fb.VertexRule(g, order)
            assert np.abs(qf.xref - xref).max() < 1e-15 and np.array_equal(qf.w, w)
    # order 2 on the triangle: vertices then the midpoints of the edges (1,2), (2,3), (3,1)
    assert np.allclose(orc.vertex_rule("Triangle2D", 2)[0][3:], [[0.5, 0.0], [0.5, 0.5], [0.0, 0.5]])
