"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the cell-wise assembly hot path of
WIAS-PDELib/ExtendableFEMBase.jl v1.6.0.  NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this file; the
product path (``extendablefembase.jl_b200/`` + ``libfemb.so``) never does.

What is restated (reference file:line, all relative to /root/reference):

* quadrature tables           src/quadrature.jl:257-279 (triangle), :352-409 (tet), :631-665 (Stroud)
* reference bases             src/fedefs/h1_p1.jl:64-76, h1_p2.jl:130-163, h1_p3.jl:142-197,
                              h1_pk.jl:170-276, hdiv_rt0.jl:60-100, hdiv_bdm1.jl:73-90,123-197, l2_p0.jl:65-72
* per-cell handlers           h1_pk.jl:280-304, h1_p3.jl:200-241, hdiv_rt0.jl:114-124, hdiv_bdm1.jl:200-249
* FEEvaluator tables          src/feevaluator.jl:83-215, :233-291 (ForwardDiff -> complex-step here)
* update_basis!               src/feevaluator_h1.jl:2-14,61-74,119-130; feevaluator_hdiv.jl:2-19,66-83
* CellDofs                    src/dofmaps.jl:419-627 (plain loops, small meshes only)
* assembly loops              examples/Example200_LowLevelPoisson.jl:103-184,
                              examples/Example210_LowLevelNavierStokes.jl:218-394
* L2GTransformer, CellVolumes [EXT E1] ExtendableGrids (compat "1.13.0", no Manifest -> version unpinned);
                              semantics constrained by the call sites src/feevaluator.jl:344-363
* rawupdateindex!/flush!      [EXT E2] ExtendableSparse (compat "1.5.1, 2", unpinned): A[i,j] += v with
                              structural insertion, flush! -> CSC with ascending rows per column

Parity pinning (tests/test_oracle_pins.py): the reference's own known-answer tests that touch this
path — nodal Kronecker property of every basis in exact arithmetic (test/test_febasis.jl:24-46),
quadrature exactness per order (test/test_quadrature.jl:16-54), P2 derivative constants on the
triangle [-1 0; 1 0; 0 1] (test/test_operators.jl:76-83), interpolation -> evaluation round trip
for H1 and Hdiv elements incl. BDM1-3D orientations (test/test_interpolators.jl:43-115), block
offsets (test/test_fematrix_and_vector.jl).  The ASSEMBLED matrix/vector values are NOT pinned by
any reference test (Example200:192-212 checks allocations only) and Julia cannot run in this
image: for those the header of DESIGN.md says "parity unpinned (values)"; analytic identities
(Appendix C of SURVEY.md) substitute.

Everything here is vectorised over cells but keeps the reference's per-entry operation order
(qp-outer, component-inner dot products; cells added in ascending order via np.add.at).
"""
from __future__ import annotations

import numpy as np

TRI, TET = "Triangle2D", "Tetrahedron3D"
DIM = {TRI: 2, TET: 3}

# ---------------------------------------------------------------------------------------------
# quadrature
# ---------------------------------------------------------------------------------------------


def quadrature(geometry: str, order: int):
    """QuadratureRule{Float64,EG}(order) -> (xref[nqp,dim], w[nqp])."""
    if geometry == TRI:
        if order <= 1:  # quadrature.jl:258-262
            return np.full((1, 2), 1.0 / 3.0), np.array([1.0])
        if order == 2:  # :263-269
            return np.array([[0.5, 0.5], [0.0, 0.5], [0.5, 0.0]]), np.array([1 / 3, 1 / 3, 1 / 3])
        if order == 8 or 12 <= order <= 14:
            raise NotImplementedError("symmetric rules are not on the assembly path")
        return _stroud(order)  # :274-276
    if geometry == TET:
        if order <= 1:  # :353-357
            return np.full((1, 3), 0.25), np.array([1.0])
        if order == 2:  # :358-368
            a, b = 0.1381966011250105, 0.5854101966249685
            return np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]]), np.full(4, 0.25)
        if order <= 3:  # :369-381 (Keast)
            s = 1.0 / 6.0
            return (np.array([[0.25, 0.25, 0.25], [0.5, s, s], [s, s, s], [s, s, 0.5], [s, 0.5, s]]),
                    np.array([-4.0 / 5.0, 9.0 / 20.0, 9.0 / 20.0, 9.0 / 20.0, 9.0 / 20.0]))
        if order <= 4:  # :382-401 (Keast, 11 points)
            p, q = 0.7857142857142857, 0.0714285714285714
            u, v = 0.1005964238332008, 0.3994035761667992
            xr = np.array([[0.25, 0.25, 0.25], [p, q, q], [q, q, q], [q, q, p], [q, p, q],
                           [u, v, v], [v, u, v], [v, v, u], [v, u, u], [u, v, u], [u, u, v]])
            return xr, np.array([-0.0789333333333333] + [0.0457333333333333] * 4 + [0.1493333333333333] * 6)
        raise NotImplementedError("symmetric tet rules of order > 4 are not used by the assembly configs")
    raise ValueError(geometry)


def vertex_rule(geometry: str, order: int = 1):
    """VertexRule(EG, order) (quadrature.jl:103-138 triangle, :157-186 tetrahedron): Lagrange points, equal weights."""
    d = DIM[geometry]
    pts = [[1.0 / (d + 1)] * d] if order == 0 else [[0.0] * d] + [[float(i == j) for i in range(d)] for j in range(d)]
    edges = [(0, 1), (1, 2), (2, 0)] if d == 2 else [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    for a, b in edges:  # quadrature.jl:110-115 / :165-170
        for j in range(1, order):
            pts.append([j / order * pts[b][i] + (1 - j / order) * pts[a][i] for i in range(d)])
    if d == 2:  # :117-132
        extra = {3: [(1, 1)], 4: [(1, 1), (2, 1), (1, 2)], 5: [(1, 1), (3, 1), (1, 3), (2, 1), (2, 2), (1, 2)]}.get(order, [])
        pts += [[x / order, y / order] for x, y in extra]
    elif order > 2:  # :172-179
        faces = [(0, 2, 1), (0, 1, 3), (1, 2, 3), (0, 3, 2)]
        for j in range(1, order - 1):
            for k in range(1, order - 1):
                for f in faces:
                    pts.append([j / order * pts[f[2]][i] + k / order * pts[f[1]][i] + (1 - j / order - k / order) * pts[f[0]][i] for i in range(d)])
    return np.array(pts), np.full(len(pts), 1.0 / len(pts))


def compute_error(el, g, coords, cellnodes, cellvolumes, celldofs, uh, u, order, p=2, op="Identity", facesigns=None, edgesigns=None, orient=None):
    """Example210_LowLevelNavierStokes.jl:110-154: error[cell, d] = sum_qp (uh_d(x_qp) - u_d(x_qp))^p |T| w_qp over the VertexRule."""
    xref, w = vertex_rule(g, order)
    cv = update_basis(el, op, g, xref, coords, cellnodes, facesigns, edgesigns, orient)  # [cell, r, dof, qp]
    coef = np.asarray(uh)[celldofs - 1]  # [cell, dof]
    uhq = np.einsum("crdq,cd->cqr", cv, coef)  # eval_febe! (:135)
    if u is not None:
        A, b, det, M = l2g(coords, cellnodes)
        x = np.einsum("ckj,qj->cqk", A, xref) + b[:, None, :]  # eval_trafo! (:140)
        uhq = uhq - np.asarray(u(x)).reshape(uhq.shape)
    return np.einsum("cqr,q,c->cr", uhq ** p, w, cellvolumes)  # :146


def _stroud(order: int):
    """quadrature.jl:631-665.  The reference gets the 1-D nodes from Golub-Welsch eigen-decompositions;
    here they come from scipy's Gauss-Legendre / Gauss-Jacobi(1,0) routines (same nodes, ascending)."""
    from scipy.special import roots_jacobi, roots_legendre

    n = order // 2 + 1
    r, a = roots_legendre(n)  # weights sum to 2
    s, b = roots_jacobi(n, 1.0, 0.0)  # weight (1-x), weights sum to 2
    r, s, a, b = 0.5 * r + 0.5, 0.5 * s + 0.5, 0.5 * a, 0.5 * b
    xref = np.empty((n * n, 2))
    w = np.empty(n * n)
    for j in range(n):  # s_j outer
        for i in range(n):  # r_i inner
            q = i + n * j
            xref[q, 0] = s[j]
            xref[q, 1] = -r[i] * (s[j] - 1.0)
            w[q] = a[i] * b[j]
    return xref, w


# ---------------------------------------------------------------------------------------------
# reference bases.  f(x) with x[..., dim] (float or complex) -> [..., nd_all, ncomp]
# ---------------------------------------------------------------------------------------------
class Element:
    """name in {H1P1,H1P2,H1P3,H1Pk,HDIVRT0,HDIVBDM1,HCURLN0,L2P0}; ncomp; order (H1Pk)."""

    def __init__(self, name, ncomp=1, order=None):
        self.name, self.ncomp = name, int(ncomp)
        self.order = {"H1P1": 1, "H1P2": 2, "H1P3": 3, "L2P0": 0, "HDIVRT0": 1, "HDIVBDM1": 1, "HCURLN0": 1}.get(name, order)
        self.family = "Hdiv" if name.startswith("HDIV") else ("Hcurl" if name.startswith("HCURL") else ("L2" if name.startswith("L2") else "H1"))
        self.vector_valued = self.family in ("Hdiv", "Hcurl")  # one block of nd vector functions instead of ncomp scalar blocks

    def polyorder(self):
        return self.order

    def pattern(self, g):
        d, o = DIM[g], self.order
        return {
            "H1P1": "N1", "H1P2": "N1F1" if d == 2 else "N1E1", "H1P3": "N1F2I1" if d == 2 else "N1E2F1",
            "H1Pk": "N1" if o == 1 else ("N1F1" if o == 2 else f"N1F{o - 1}I{(o - 2) * (o - 1) // 2}"),
            "HDIVRT0": "f1", "HDIVBDM1": "f2" if d == 2 else "f3", "L2P0": "I1",
            "HCURLN0": "f1" if d == 2 else "e1",  # hcurl_n0.jl:31-34
        }[self.name]

    def scalar_basis(self, g, x):
        """one-component H1/L2 basis -> list of arrays"""
        d = DIM[g]
        one = np.ones(x.shape[:-1], dtype=x.dtype)
        xs = [x[..., i] for i in range(d)]
        l0 = one - sum(xs)
        n, o = self.name, self.order
        if n == "L2P0":
            return [one]
        if n == "H1P1" or (n == "H1Pk" and o == 1):
            return [l0] + xs
        if n == "H1P2":
            if d == 2:
                x1, x2 = xs
                return [2 * l0 * (l0 - 0.5), 2 * x1 * (x1 - 0.5), 2 * x2 * (x2 - 0.5), 4 * l0 * x1, 4 * x1 * x2, 4 * x2 * l0]
            x1, x2, x3 = xs
            return [2 * l0 * (l0 - 0.5), 2 * x1 * (x1 - 0.5), 2 * x2 * (x2 - 0.5), 2 * x3 * (x3 - 0.5),
                    4 * l0 * x1, 4 * l0 * x2, 4 * l0 * x3, 4 * x1 * x2, 4 * x1 * x3, 4 * x2 * x3]
        if n == "H1P3":
            t, tt = 1.0 / 3.0, 2.0 / 3.0
            if d == 2:
                x1, x2 = xs
                return [4.5 * l0 * (l0 - t) * (l0 - tt), 4.5 * x1 * (x1 - t) * (x1 - tt), 4.5 * x2 * (x2 - t) * (x2 - tt),
                        13.5 * x1 * l0 * (l0 - t), 13.5 * x1 * l0 * (x1 - t),
                        13.5 * x2 * x1 * (x1 - t), 13.5 * x2 * x1 * (x2 - t),
                        13.5 * l0 * x2 * (x2 - t), 13.5 * l0 * x2 * (l0 - t),
                        27 * x1 * x2 * l0]
            x1, x2, x3 = xs
            return [4.5 * l0 * (l0 - t) * (l0 - tt), 4.5 * x1 * (x1 - t) * (x1 - tt), 4.5 * x2 * (x2 - t) * (x2 - tt), 4.5 * x3 * (x3 - t) * (x3 - tt),
                    13.5 * x1 * l0 * (l0 - t), 13.5 * x1 * l0 * (x1 - t),
                    13.5 * x2 * l0 * (l0 - t), 13.5 * x2 * l0 * (x2 - t),
                    13.5 * x3 * l0 * (l0 - t), 13.5 * x3 * l0 * (x3 - t),
                    13.5 * x1 * x2 * (x1 - t), 13.5 * x1 * x2 * (x2 - t),
                    13.5 * x1 * x3 * (x1 - t), 13.5 * x1 * x3 * (x3 - t),
                    13.5 * x2 * x3 * (x2 - t), 13.5 * x2 * x3 * (x3 - t),
                    27 * x1 * x2 * l0, 27 * x1 * x3 * l0, 27 * x1 * x2 * x3, 27 * x2 * x3 * l0]
        if n == "H1Pk":
            if d != 2:
                raise NotImplementedError("H1Pk: Triangle2D only on this path")
            return _h1pk_tri(o, l0, xs[0], xs[1], one)
        raise ValueError(n)

    def basis(self, g, x):
        x = np.asarray(x)
        d = DIM[g]
        if not self.vector_valued:
            sb = self.scalar_basis(g, x)
            nds = len(sb)
            out = np.zeros(x.shape[:-1] + (nds * self.ncomp, self.ncomp), dtype=x.dtype)
            for c in range(self.ncomp):  # component-blocked (h1_p2.jl:133-141)
                for j, p in enumerate(sb):
                    out[..., c * nds + j, c] = p
            return out
        xs = [x[..., i] for i in range(d)]
        if self.name == "HDIVRT0":
            if d == 2:
                x1, x2 = xs
                rows = [[x1, x2 - 1], [x1, x2], [x1 - 1, x2]]
            else:
                x1, x2, x3 = xs
                rows = [[2 * x1, 2 * x2, 2 * (x3 - 1)], [2 * x1, 2 * (x2 - 1), 2 * x3], [2 * x1, 2 * x2, 2 * x3], [2 * (x1 - 1), 2 * x2, 2 * x3]]
        elif self.name == "HDIVBDM1":
            if d == 2:
                x1, x2 = xs
                rows = [[x1, x2 - 1], [6 * x1, 6 - 12 * x1 - 6 * x2], [x1, x2], [-6 * x1, 6 * x2], [x1 - 1, x2], [6 * (x1 - 1) + 12 * x2, -6 * x2]]
            else:
                x1, x2, x3 = xs
                l = 1 - x1 - x2 - x3
                z = 0 * x1
                rows = [
                    [2 * x1, 2 * x2, 2 * (x3 - 1)],
                    [24 * x1, z, 24 * (l - x1)], [z, -24 * x2, -24 * (l - x2)], [-24 * x1, 24 * x2, -24 * (x2 - x1)],
                    [2 * x1, 2 * (x2 - 1), 2 * x3],
                    [z, 24 * (l - x3), 24 * x3], [-24 * x1, -24 * (l - x1), z], [24 * x1, -24 * (x1 - x3), -24 * x3],
                    [2 * x1, 2 * x2, 2 * x3],
                    [-24 * x1, z, 24 * x3], [24 * x1, -24 * x2, z], [z, 24 * x2, -24 * x3],
                    [2 * (x1 - 1), 2 * x2, 2 * x3],
                    [24 * (l - x2), 24 * x2, z], [-24 * (l - x3), z, -24 * x3], [-24 * (x3 - x2), -24 * x2, 24 * x3],
                ]
        elif self.name == "HCURLN0":
            one = np.ones(x.shape[:-1], dtype=x.dtype)
            if d == 2:  # hcurl_n0.jl:94-104
                x1, x2 = xs
                rows = [[one - x2, x1], [-x2, x1], [-x2, x1 - 1]]
            else:  # hcurl_n0.jl:120-142
                x1, x2, x3 = xs
                z = 0 * x1
                rows = [[one - x2 - x3, x1, x1], [x2, one - x3 - x1, x2], [x3, x3, one - x1 - x2],
                        [-x2, x1, z], [-x3, z, x1], [z, -x3, x2]]
        else:
            raise ValueError(self.name)
        out = np.zeros(x.shape[:-1] + (len(rows), d), dtype=x.dtype)
        for i, r in enumerate(rows):
            for k in range(d):
                out[..., i, k] = r[k]
        return out

    def nd_all(self, g):
        return self.basis(g, np.zeros((1, DIM[g]))).shape[-2]

    def nd(self, g):
        if self.name == "HDIVBDM1" and DIM[g] == 3:
            return 12
        return self.nd_all(g)

    def tabulate(self, g, xref):
        """refbasisvals [qp, dof_all, comp] and d/dxref_j [qp, dof_all, comp, j] by complex step
        (exact to rounding for polynomials, like ForwardDiff's dual numbers, feevaluator.jl:233-291)."""
        xref = np.asarray(xref, dtype=np.float64)
        vals = self.basis(g, xref)
        d = DIM[g]
        h = 1e-40
        der = np.zeros(vals.shape + (d,))
        for j in range(d):
            xc = xref.astype(np.complex128)
            xc[:, j] += 1j * h
            der[..., j] = self.basis(g, xc).imag / h
        return vals, der


def _h1pk_tri(order, l0, x1, x2, one):
    """h1_pk.jl:170-276, orders 2..5"""
    co = [k / order for k in range(order + 1)]
    from fractions import Fraction as Fr

    cf = [Fr(k, order) for k in range(order + 1)]
    fnode = Fr(1)
    for c in cf[1:]:
        fnode *= c
    ff = [Fr(1)] * (order - 1)
    for j in range(2, order + 1):
        for k in range(1, order + 2):
            if k < j:
                ff[j - 2] *= cf[j - 1] - cf[k - 1]
            elif k > j:
                ff[j - 2] *= cf[k - 1] - cf[j - 1]
    fnode = float(fnode)
    ff = [float(f) for f in ff]
    out = []
    for v in (l0, x1, x2):
        p = one
        for k in range(order):
            p = p * (v - co[k])
        out.append(p / fnode)
    for fa, fb in ((l0, x1), (x1, x2), (x2, l0)):
        for k in range(1, order):
            p = fa * fb / ff[k - 1]
            if order > 2:
                for m in range(1, order):
                    if m > k:
                        p = p * (fa - (1 - co[m]))
                    elif m < k:
                        p = p * (fb - co[m])
            out.append(p)
    bub = l0 * x1 * x2
    if order == 3:
        out.append(bub * 27)
    elif order == 4:
        out += [bub * (l0 - 0.25) * 128, bub * (x1 - 0.25) * 128, bub * (x2 - 0.25) * 128]
    elif order == 5:
        out += [bub * (l0 - 0.2) * (l0 - 0.4) * (3125 / 6), bub * (x1 - 0.2) * (x1 - 0.4) * (3125 / 6), bub * (x2 - 0.2) * (x2 - 0.4) * (3125 / 6),
                bub * (l0 - 0.2) * (x1 - 0.2) * (3125 / 4), bub * (x1 - 0.2) * (x2 - 0.2) * (3125 / 4), bub * (x2 - 0.2) * (l0 - 0.2) * (3125 / 4)]
    elif order > 5:
        raise NotImplementedError
    return out


# ---------------------------------------------------------------------------------------------
# per-cell handlers
# ---------------------------------------------------------------------------------------------
def basis_subset(el: Element, g, ncells, facesigns=None, edgesigns=None, orient=None):
    """subset_ids[cell, dof] (0-based) — get_basissubset"""
    nd = el.nd(g)
    sub = np.tile(np.arange(nd), (ncells, 1))
    d = DIM[g]
    if el.family == "H1" and el.order is not None and el.order >= 3:
        nds = nd // el.ncomp
        if d == 2:  # h1_pk.jl:280-304 / h1_p3.jl:200-217
            nf = el.order - 1
            for j in range(3):
                neg = facesigns[:, j] != 1
                for c in range(el.ncomp):
                    for k in range(nf):
                        sub[neg, c * nds + 3 + j * nf + k] = c * nds + 3 + j * nf + (nf - 1 - k)
        else:  # h1_p3.jl:220-241
            for j in range(6):
                neg = edgesigns[:, j] != 1
                for c in range(el.ncomp):
                    sub[neg, c * 20 + 4 + 2 * j] = c * 20 + 4 + 2 * j + 1
                    sub[neg, c * 20 + 4 + 2 * j + 1] = c * 20 + 4 + 2 * j
    elif el.name == "HDIVBDM1" and d == 3:  # hdiv_bdm1.jl:234-249
        s1 = np.array([1, 0, 1, 2])
        s2 = np.array([2, 2, 0, 1])
        for j in range(1, 5):
            o = orient[:, j - 1] - 1
            sub[:, 3 * j - 3] = 4 * j - 3 - 1
            sub[:, 3 * j - 2] = 4 * j - s1[o] - 1
            sub[:, 3 * j - 1] = 4 * j - s2[o] - 1
    return sub


def coefficients(el: Element, g, ncells, facesigns=None, edgesigns=None):
    """coefficients[cell, dof] (identical for all components in every on-path element) — get_coefficients"""
    nd = el.nd(g)
    co = np.ones((ncells, nd))
    d = DIM[g]
    if el.name == "HDIVRT0":  # hdiv_rt0.jl:114-124
        co[:] = facesigns
    elif el.name == "HCURLN0":  # hcurl_n0.jl:144-166: face signs on triangles, edge signs on tetrahedra
        co[:] = facesigns if d == 2 else edgesigns
    elif el.name == "HDIVBDM1":
        if d == 2:  # hdiv_bdm1.jl:200-212
            co[:, 0::2] = facesigns
        else:  # :215-228
            co[:, 0::3] = facesigns
            co[:, 1::3] = -1.0
            co[:, 2::3] = 1.0
    return co


# ---------------------------------------------------------------------------------------------
# stage 1: L2GTransformer   [EXT E1]
# ---------------------------------------------------------------------------------------------
def l2g(coords, cellnodes):
    """A[cell,k,i] = x_{i+1,k} - x_{1,k}; b = x_1; det; M = A^{-T} (mapderiv!)."""
    x = coords[cellnodes.astype(np.int64) - 1]  # (nc, d+1, d)
    b = x[:, 0, :]
    A = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))  # A[:, k, i]
    d = coords.shape[1]
    if d == 2:
        det = A[:, 0, 0] * A[:, 1, 1] - A[:, 0, 1] * A[:, 1, 0]
        M = np.empty_like(A)
        M[:, 0, 0] = A[:, 1, 1] / det
        M[:, 0, 1] = -A[:, 1, 0] / det
        M[:, 1, 0] = -A[:, 0, 1] / det
        M[:, 1, 1] = A[:, 0, 0] / det
    else:
        C = np.empty_like(A)
        for j in range(3):
            for k in range(3):
                j1, j2 = (j + 1) % 3, (j + 2) % 3
                k1, k2 = (k + 1) % 3, (k + 2) % 3
                C[:, j, k] = A[:, j1, k1] * A[:, j2, k2] - A[:, j1, k2] * A[:, j2, k1]
        det = A[:, 0, 0] * C[:, 0, 0] + A[:, 0, 1] * C[:, 0, 1] + A[:, 0, 2] * C[:, 0, 2]
        M = C / det[:, None, None]
    return A, b, det, M


def cell_volumes(coords, cellnodes):
    _, _, det, _ = l2g(coords, cellnodes)
    fact = 2.0 if coords.shape[1] == 2 else 6.0
    return np.abs(det) / fact


def eval_trafo(A, b, xref):
    """x[cell, qp, :] = A xref[qp] + b   (eval_trafo!)"""
    return np.einsum("cki,qi->cqk", A, xref) + b[:, None, :]


# ---------------------------------------------------------------------------------------------
# stage 2: update_basis!  -> cvals[cell, r, dof, qp]  (Julia FEB.cvals[r, dof, qp] per cell)
# ---------------------------------------------------------------------------------------------
def update_basis(el: Element, op: str, g, xref, coords, cellnodes, facesigns=None, edgesigns=None, orient=None):
    d = DIM[g]
    nc = cellnodes.shape[0]
    vals, der = el.tabulate(g, xref)  # [qp, da, comp], [qp, da, comp, j]
    A, b, det, M = l2g(coords, cellnodes)
    sub = basis_subset(el, g, nc, facesigns, edgesigns, orient)  # [cell, dof]
    ncomp = el.ncomp if not el.vector_valued else d
    nd, nq = el.nd(g), xref.shape[0]
    if not el.vector_valued:
        if op == "Identity":  # feevaluator_h1.jl:2-14
            v = vals[:, sub, :]  # [qp, cell, dof, comp]
            return np.ascontiguousarray(np.transpose(v, (1, 3, 2, 0)))
        dsub = der[:, sub, :, :]  # [qp, cell, dof, comp, j]
        if op == "Gradient":  # feevaluator_h1.jl:61-74: cvals[k + d*(c-1)] = sum_j M[k,j] dref[.., c, j]
            out = np.zeros((nc, d * ncomp, nd, nq))
            for c in range(ncomp):
                for k in range(d):
                    acc = np.zeros((nq, nc, nd))
                    for j in range(d):
                        acc = acc + M[None, :, k, j, None] * dsub[:, :, :, c, j]
                    out[:, k + d * c, :, :] = np.transpose(acc, (1, 2, 0))
            return out
        if op == "Divergence":  # feevaluator_h1.jl:119-130
            acc = np.zeros((nq, nc, nd))
            for k in range(d):
                for j in range(d):
                    acc = acc + M[None, :, k, j, None] * dsub[:, :, :, k, j]
            return np.transpose(acc, (1, 2, 0))[:, None, :, :]
        raise ValueError(op)
    co = coefficients(el, g, nc, facesigns, edgesigns)  # [cell, dof]
    if el.family == "Hcurl":
        if op != "Identity":
            raise ValueError(op)
        v = vals[:, sub, :]  # feevaluator_hcurl.jl:2-16: (sum_l L2GAinv[k,l] phi[sub, l]) * coeff
        out = np.zeros((nc, d, nd, nq))
        for k in range(d):
            acc = np.zeros((nq, nc, nd))
            for l in range(d):
                acc = acc + M[None, :, k, l, None] * v[:, :, :, l]
            out[:, k, :, :] = np.transpose(acc * co[None], (1, 2, 0))
        return out
    if op == "Identity":  # feevaluator_hdiv.jl:2-19: (sum_l A[k,l] phi[sub, l]) * coeff / det
        v = vals[:, sub, :]  # [qp, cell, dof, l]
        out = np.zeros((nc, d, nd, nq))
        for k in range(d):
            acc = np.zeros((nq, nc, nd))
            for l in range(d):
                acc = acc + A[None, :, k, l, None] * v[:, :, :, l]
            out[:, k, :, :] = np.transpose(acc * co[None] / det[None, :, None], (1, 2, 0))
        return out
    if op == "Divergence":  # feevaluator_hdiv.jl:66-83
        dsub = der[:, sub, :, :]
        acc = np.zeros((nq, nc, nd))
        for j in range(d):
            acc = acc + dsub[:, :, :, j, j]
        return np.transpose(acc * co[None] / det[None, :, None], (1, 2, 0))[:, None, :, :]
    raise ValueError(op)


# ---------------------------------------------------------------------------------------------
# CellDofs   (dofmaps.jl:419-627, plain loops; 1-based int32 out, shape (ncells, nd))
# ---------------------------------------------------------------------------------------------
def parse_pattern(p):
    import re

    return [(m.group(1), int(m.group(2))) for m in re.finditer(r"([A-Za-z])(\d+)", p)]


def celldofs_from_pattern(el: Element, g, nnodes, cellnodes, cellfaces=None, nfaces=0, celledges=None, nedges=0):
    segs = parse_pattern(el.pattern(g))
    ncells = cellnodes.shape[0]
    ncomp = el.ncomp if not el.vector_valued else 1
    # dofs per component = coffset (finiteelements.jl:395-449)
    cnt = {"N": nnodes, "F": nfaces, "E": nedges, "I": ncells}
    coffset = sum(cnt[t] * q for t, q in segs if t.isupper())
    ndofs = coffset * (el.ncomp if not el.vector_valued else 1) + sum(cnt[t.upper()] * q for t, q in segs if t.islower())
    if el.vector_valued:
        coffset, ncomp = 0, 0
    out = []
    for cell in range(ncells):
        row = []
        for c in range(max(ncomp, 0)):
            off = c * coffset
            for t, q in segs:
                if not t.isupper():
                    continue
                if t == "N":
                    for n in range(cellnodes.shape[1]):
                        for m in range(q):
                            row.append(int(cellnodes[cell, n]) + off + m * nnodes)
                    off += nnodes * q
                elif t == "F":
                    for n in range(cellfaces.shape[1]):
                        for m in range(q):
                            row.append(int(cellfaces[cell, n]) + off + m * nfaces)
                    off += nfaces * q
                elif t == "E":
                    for n in range(celledges.shape[1]):
                        for m in range(q):
                            row.append(int(celledges[cell, n]) + off + m * nedges)
                    off += nedges * q
                elif t == "I":
                    for m in range(q):
                        row.append(cell + 1 + off)
                        off += ncells
        off = (el.ncomp if not el.vector_valued else 0) * coffset
        for t, q in segs:
            if t.isupper():
                continue
            if t == "n":
                for n in range(cellnodes.shape[1]):
                    for m in range(q):
                        row.append(int(cellnodes[cell, n]) + off + m * nnodes)
                off += nnodes * q
            elif t == "f":
                for n in range(cellfaces.shape[1]):
                    for m in range(q):
                        row.append(int(cellfaces[cell, n]) + off + m * nfaces)
                off += nfaces * q
            elif t == "e":
                for n in range(celledges.shape[1]):
                    for m in range(q):
                        row.append(int(celledges[cell, n]) + off + m * nedges)
                off += nedges * q
            elif t == "i":
                for m in range(q):
                    row.append(cell + 1 + off)
                    off += ncells
        out.append(row)
    return np.asarray(out, dtype=np.int32), ndofs


# ---------------------------------------------------------------------------------------------
# sparse container  [EXT E2]: accumulate (i, j, v) triplets in insertion order, flush to Julia CSC
# ---------------------------------------------------------------------------------------------
def flush_csc(n_rows, n_cols, I, J, V):
    """Triplets (1-based I, J) -> (colptr, rowval, nzval) 1-based int64, rows ascending per column,
    duplicates summed in insertion order (the serial reference loop's order)."""
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    V = np.asarray(V, dtype=np.float64)
    key = (J - 1) * n_rows + (I - 1)
    ukey, inv = np.unique(key, return_inverse=True)
    nzval = np.zeros(ukey.shape[0])
    np.add.at(nzval, inv, V)
    rowval = ukey % n_rows + 1
    cols = ukey // n_rows
    colptr = np.ones(n_cols + 1, dtype=np.int64)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr) - np.arange(n_cols + 1)
    return colptr.astype(np.int64), rowval.astype(np.int64), nzval


def add_into_csc(colptr, rowval, nzval, I, J, V):
    """rawupdateindex!(A, +, v, i, j) into an existing (flushed) pattern; returns count of misses."""
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    n_rows = int(rowval.max()) if rowval.size else 1
    cols = np.repeat(np.arange(colptr.shape[0] - 1), np.diff(colptr))
    key_all = cols * (n_rows + 1) + rowval
    key = (J - 1) * (n_rows + 1) + I
    pos = np.searchsorted(key_all, key)
    pos_c = np.minimum(pos, key_all.shape[0] - 1)
    hit = key_all[pos_c] == key
    np.add.at(nzval, pos_c[hit], np.asarray(V)[hit])
    return int((~hit).sum())


# ---------------------------------------------------------------------------------------------
# Example200.assemble!   (examples/Example200_LowLevelPoisson.jl:103-184)
# ---------------------------------------------------------------------------------------------
def example200_local(el, g, coords, cellnodes, cellvolumes, mu, rhs, facesigns=None, edgesigns=None, order=None):
    """-> Aloc[cell, j, k] (upper triangle filled, j <= k; already * mu * |T|) and bloc[cell, j]."""
    xref, w = quadrature(g, 2 * (el.polyorder() - 1) if order is None else order)  # :110
    grad = update_basis(el, "Gradient", g, xref, coords, cellnodes, facesigns, edgesigns)  # [cell, r, dof, qp]
    idv = update_basis(el, "Identity", g, xref, coords, cellnodes, facesigns, edgesigns)
    nc, R, nd, nq = grad.shape
    Aloc = np.zeros((nc, nd, nd))
    for j in range(nd):
        for k in range(j, nd):
            temp = np.zeros(nc)
            for qp in range(nq):  # :139-145
                dot = np.zeros(nc)
                for r in range(R):
                    dot = dot + grad[:, r, j, qp] * grad[:, r, k, qp]
                temp = temp + w[qp] * dot
            Aloc[:, j, k] = temp
    Aloc *= (mu * cellvolumes)[:, None, None]  # :146
    bloc = None
    if rhs is not None:
        A, b, _, _ = l2g(coords, cellnodes)
        x = eval_trafo(A, b, xref)  # [cell, qp, dim]
        f = rhs(x)  # [cell, qp]
        bloc = np.zeros((nc, nd))
        for j in range(nd):
            temp = np.zeros(nc)
            for qp in range(nq):  # :170-175
                temp = temp + w[qp] * idv[:, 0, j, qp] * f[:, qp]
            bloc[:, j] = temp * cellvolumes  # :177
    return Aloc, bloc


def example200_assemble(el, g, coords, cellnodes, cellvolumes, celldofs, ndofs, mu=1.0, rhs=None, facesigns=None, edgesigns=None,
                        drop_tol=1.0e-15, order=None):
    """First assembly into an empty matrix + flush!: returns (colptr, rowval, nzval, b)."""
    Aloc, bloc = example200_local(el, g, coords, cellnodes, cellvolumes, mu, rhs, facesigns, edgesigns, order)
    I, J, V = example200_triplets(Aloc, celldofs, drop_tol)
    colptr, rowval, nzval = flush_csc(ndofs, ndofs, I, J, V)
    b = np.zeros(ndofs)
    if bloc is not None:
        np.add.at(b, celldofs.astype(np.int64).reshape(-1) - 1, bloc.reshape(-1))
    return colptr, rowval, nzval, b


def example200_triplets(Aloc, celldofs, drop_tol=1.0e-15):
    """Insertion sequence of Example200:149-161 (cell-major, j, k >= j, mirrored entry right after)."""
    nc, nd, _ = Aloc.shape
    cd = celldofs.astype(np.int64)
    I, J, V = [], [], []
    for j in range(nd):
        for k in range(j, nd):
            v = Aloc[:, j, k]
            keep = np.abs(v) > drop_tol
            seq = np.nonzero(keep)[0] * (nd * nd * 2) + (j * nd + k) * 2
            I.append((cd[keep, j], seq))
            J.append(cd[keep, k])
            V.append(v[keep])
            if k > j:
                I.append((cd[keep, k], seq + 1))
                J.append(cd[keep, j])
                V.append(v[keep])
    seq = np.concatenate([s for _, s in I])
    order = np.argsort(seq, kind="stable")
    return np.concatenate([i for i, _ in I])[order], np.concatenate(J)[order], np.concatenate(V)[order]


# ---------------------------------------------------------------------------------------------
# Example210.update_system!(linear, nonlinear)   (examples/Example210_LowLevelNavierStokes.jl:218-394)
# ---------------------------------------------------------------------------------------------
def example210_f(x, mu, t):
    e = 8.0 * np.pi * np.pi * mu * np.exp(-8.0 * np.pi * np.pi * mu * t)
    return np.stack([e * np.sin(2.0 * np.pi * x[..., 0]) * np.sin(2.0 * np.pi * x[..., 1]),
                     e * np.cos(2.0 * np.pi * x[..., 0]) * np.cos(2.0 * np.pi * x[..., 1])], axis=-1)


def example210_local(coords, cellnodes, cellvolumes, celldofs_u, sol, mu, teval, linear, nonlinear, order_u=2, f=None):
    """-> Aloc[cell, row, col] (already * |T|), Bloc[cell, j, k] (* |T|) or None, bloc_u[cell, j]."""
    g = TRI
    elu, elp = Element("H1Pk", 2, order_u), Element("H1Pk", 1, order_u - 1)
    xref, w = quadrature(g, 3 * order_u - 1)  # :237
    G = update_basis(elu, "Gradient", g, xref, coords, cellnodes)  # [cell, 4, ndu, qp]
    Iu = update_basis(elu, "Identity", g, xref, coords, cellnodes)  # [cell, 2, ndu, qp]
    Ip = update_basis(elp, "Identity", g, xref, coords, cellnodes)  # [cell, 1, ndp, qp]
    nc, _, ndu, nq = G.shape
    ndp = Ip.shape[2]
    Aloc = np.zeros((nc, ndu, ndu))
    Bloc = None
    bloc = np.zeros((nc, ndu))
    if linear:
        for j in range(ndu):
            for k in range(ndu):
                temp = np.zeros(nc)
                for qp in range(nq):  # :284-290
                    dot = np.zeros(nc)
                    for r in range(4):
                        dot = dot + G[:, r, j, qp] * G[:, r, k, qp]
                    temp = temp + w[qp] * dot
                Aloc[:, k, j] = mu * temp
        Bloc = np.zeros((nc, ndu, ndp))
        for j in range(ndu):
            for k in range(ndp):
                temp = np.zeros(nc)
                for qp in range(nq):  # :293-300
                    temp = temp - w[qp] * (G[:, 0, j, qp] + G[:, 3, j, qp]) * Ip[:, 0, k, qp]
                Bloc[:, j, k] = temp
        Bloc *= cellvolumes[:, None, None]  # :301
        A, b0, _, _ = l2g(coords, cellnodes)
        x = eval_trafo(A, b0, xref)
        fv = example210_f(x, mu, teval) if f is None else f(x)  # [cell, qp, 2]
        for j in range(ndu):
            temp = np.zeros(nc)
            for qp in range(nq):  # :306-315
                temp = temp + w[qp] * (Iu[:, 0, j, qp] * fv[:, qp, 0] + Iu[:, 1, j, qp] * fv[:, qp, 1])
            bloc[:, j] += temp * cellvolumes
    if nonlinear:
        sl = sol[celldofs_u.astype(np.int64) - 1]  # [cell, ndu]
        for qp in range(nq):
            inp = np.zeros((nc, 6))
            for j in range(ndu):  # :324-333
                for d in range(2):
                    inp[:, d] += sl[:, j] * Iu[:, d, j, qp]
                for d in range(4):
                    inp[:, 2 + d] += sl[:, j] * G[:, d, j, qp]
            # operator! (:251-255) and its Jacobian (ForwardDiff :336): rows = result, cols = input
            value = np.stack([inp[:, 0] * inp[:, 2] + inp[:, 1] * inp[:, 3], inp[:, 0] * inp[:, 4] + inp[:, 1] * inp[:, 5]], axis=1)
            jac = np.zeros((nc, 2, 6))
            jac[:, 0, 0], jac[:, 0, 1], jac[:, 0, 2], jac[:, 0, 3] = inp[:, 2], inp[:, 3], inp[:, 0], inp[:, 1]
            jac[:, 1, 0], jac[:, 1, 1], jac[:, 1, 4], jac[:, 1, 5] = inp[:, 4], inp[:, 5], inp[:, 0], inp[:, 1]
            for j in range(ndu):  # :339-355
                tv = np.zeros((nc, 2))
                for d in range(2):
                    tv[:, 0] += jac[:, 0, d] * Iu[:, d, j, qp]
                    tv[:, 1] += jac[:, 1, d] * Iu[:, d, j, qp]
                for d in range(4):
                    tv[:, 0] += jac[:, 0, 2 + d] * G[:, d, j, qp]
                    tv[:, 1] += jac[:, 1, 2 + d] * G[:, d, j, qp]
                for k in range(ndu):
                    Aloc[:, k, j] += (tv[:, 0] * Iu[:, 0, k, qp] + tv[:, 1] * Iu[:, 1, k, qp]) * w[qp]
            tv = np.einsum("crk,ck->cr", jac, inp) - value  # :358-359
            for j in range(ndu):
                bloc[:, j] += (tv[:, 0] * Iu[:, 0, j, qp] + tv[:, 1] * Iu[:, 1, j, qp]) * w[qp] * cellvolumes
    Aloc *= cellvolumes[:, None, None]  # :368
    return Aloc, Bloc, bloc


def example210_assemble(coords, cellnodes, cellvolumes, celldofs_u, celldofs_p, ndofs_u, ndofs_p, sol, mu=1.0e-2, teval=0.0,
                        linear=True, nonlinear=False, order_u=2, f=None):
    """First assembly + flush! -> (colptr, rowval, nzval, b).  All entries are inserted (no filter, :369-381)."""
    Aloc, Bloc, bloc = example210_local(coords, cellnodes, cellvolumes, celldofs_u, sol, mu, teval, linear, nonlinear, order_u, f)
    I, J, V = example210_triplets(Aloc, Bloc, celldofs_u, celldofs_p, ndofs_u)
    n = ndofs_u + ndofs_p
    colptr, rowval, nzval = flush_csc(n, n, I, J, V)
    b = np.zeros(n)
    np.add.at(b, celldofs_u.astype(np.int64).reshape(-1) - 1, bloc.reshape(-1))
    return colptr, rowval, nzval, b


def example210_triplets(Aloc, Bloc, celldofs_u, celldofs_p, ndofs_u):
    nc, ndu, _ = Aloc.shape
    cu = celldofs_u.astype(np.int64)
    I = [np.repeat(cu[:, :, None], ndu, axis=2).reshape(-1)]
    J = [np.repeat(cu[:, None, :], ndu, axis=1).reshape(-1)]
    V = [Aloc.reshape(-1)]
    if Bloc is not None:
        cp = celldofs_p.astype(np.int64) + ndofs_u
        ndp = cp.shape[1]
        iu = np.repeat(cu[:, :, None], ndp, axis=2).reshape(-1)
        ip = np.repeat(cp[:, None, :], ndu, axis=1).reshape(-1)
        I += [iu, ip]
        J += [ip, iu]
        V += [Bloc.reshape(-1), Bloc.reshape(-1)]
    # insertion order differs from the serial loop only between blocks of one cell; per-entry sums
    # are still in ascending cell order, which is what fixes the floating-point result
    return np.concatenate(I), np.concatenate(J), np.concatenate(V)


# ---------------------------------------------------------------------------------------------
# generic bilinear form  Aloc[j,k] = factor * (sum_qp w <opr_j, opc_k>) * |T|   (config C4: Hdiv mass / div)
# ---------------------------------------------------------------------------------------------
def bilinear_local(cv_row, cv_col, w, cellvolumes, factor=1.0):
    nc, R, ndr, nq = cv_row.shape
    ndc = cv_col.shape[2]
    out = np.zeros((nc, ndr, ndc))
    for j in range(ndr):
        for k in range(ndc):
            temp = np.zeros(nc)
            for qp in range(nq):
                dot = np.zeros(nc)
                for r in range(R):
                    dot = dot + cv_row[:, r, j, qp] * cv_col[:, r, k, qp]
                temp = temp + w[qp] * dot
            out[:, j, k] = (factor * temp) * cellvolumes
    return out


def block_triplets(Aloc, rowdofs, coldofs, roff=0, coff=0):
    nc, ndr, ndc = Aloc.shape
    I = np.repeat(rowdofs.astype(np.int64)[:, :, None] + roff, ndc, axis=2).reshape(-1)
    J = np.repeat(coldofs.astype(np.int64)[:, None, :] + coff, ndr, axis=1).reshape(-1)
    return I, J, Aloc.reshape(-1)


# ---------------------------------------------------------------------------------------------
# interpolation functionals (only to pin conventions; src/interpolators.jl, hdiv_rt0.jl:36-52, hdiv_bdm1.jl:47-67)
# ---------------------------------------------------------------------------------------------
def face_geometry(coords, facenodes):
    """-> (measure, unit normal) per face, normal orientation given by the face's own node order:
    2-D: rotate the tangent (x2-x1) clockwise; 3-D: (x2-x1) x (x3-x1)."""
    x = coords[facenodes.astype(np.int64) - 1]
    if coords.shape[1] == 2:
        t = x[:, 1] - x[:, 0]
        n = np.stack([t[:, 1], -t[:, 0]], axis=1)
        meas = np.linalg.norm(t, axis=1)
        return meas, n / meas[:, None]
    n = np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0])
    nn = np.linalg.norm(n, axis=1)
    return 0.5 * nn, n / nn[:, None]


def interpolate_hdiv(el: Element, g, coords, facenodes, u):
    """Face-moment interpolation: RT0 dof_f = int_F u.n; BDM1 additionally int_F u.n (xref_m - 1/dim)
    with xref the face's own reference coordinates (BDM1_normalflux_eval!, hdiv_bdm1.jl:47-54)."""
    d = DIM[g]
    nf = facenodes.shape[0]
    meas, nrm = face_geometry(coords, facenodes)
    x = coords[facenodes.astype(np.int64) - 1]  # [face, d, dim]
    if d == 2:
        gp, gw = np.polynomial.legendre.leggauss(3)
        fx = np.stack([0.5 * gp + 0.5], axis=1)  # xref on the edge
        fw = 0.5 * gw
        pts = x[:, None, 0, :] + fx[None, :, 0, None] * (x[:, None, 1, :] - x[:, None, 0, :])
    else:
        fx, fw = _stroud(4)
        pts = x[:, None, 0, :] + fx[None, :, 0, None] * (x[:, None, 1, :] - x[:, None, 0, :]) + fx[None, :, 1, None] * (x[:, None, 2, :] - x[:, None, 0, :])
    flux = np.einsum("fqk,fk->fq", u(pts), nrm)  # [face, q]
    nmom = 1 if el.name == "HDIVRT0" else d
    out = np.zeros(nf * nmom)
    out[0::nmom] = meas * (flux @ fw)
    for m in range(1, nmom):
        out[m::nmom] = meas * ((flux * (fx[None, :, m - 1] - 1.0 / d)) @ fw)
    # dof numbering "f{nmom}": face + (m-1)*nfaces
    res = np.zeros(nf * nmom)
    for m in range(nmom):
        res[m * nf:(m + 1) * nf] = out[m::nmom]
    return res


def interpolate_hcurl(coords, entitynodes, u):
    """HCURLN0 interpolation (hcurl_n0.jl:44-56): dof_e = int_e u . t ds with t pointing from the first to the second node of the
    entity (2-D: the faces, t = the rotated FaceNormal (-n_2, n_1); 3-D: the edges, t = EdgeTangents); 2-point Gauss rule."""
    x = coords[entitynodes.astype(np.int64) - 1]  # [entity, 2, dim]
    gp, gw = np.polynomial.legendre.leggauss(2)
    t = x[:, 1, :] - x[:, 0, :]  # |e| t
    pts = x[:, None, 0, :] + (0.5 * gp + 0.5)[None, :, None] * t[:, None, :]
    return np.einsum("eqk,ek->eq", u(pts), t) @ (0.5 * gw)


# ---------------------------------------------------------------------------------------------
# grid topology ([EXT E1] recollection of ExtendableGrids): the literal loop over cells and
# local entities, numbering faces / edges in order of first appearance.  Small meshes only (pure Python).
# ---------------------------------------------------------------------------------------------
LOCAL_FACES = {TRI: [(0, 1), (1, 2), (2, 0)], TET: [(0, 2, 1), (0, 1, 3), (1, 2, 3), (0, 3, 2)]}
LOCAL_EDGES = {TRI: [(0, 1), (1, 2), (2, 0)], TET: [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]}


def enumerate_entities(g, cellnodes, kind="faces"):
    """-> dict(cell_entities [ncells, k] 1-based, entity_nodes [n, m] in the creating cell's local order, signs, and for
    faces: facecells [n, 2] (creator, neighbour or 0), orientations (1 creator sequence; 2: (l3,l2,l1), 3: (l2,l1,l3),
    4: (l1,l3,l2) of the global face, hdiv_bdm1.jl:238-246; triangles: 1 creator / 2 neighbour))."""
    table = LOCAL_FACES[g] if kind == "faces" else LOCAL_EDGES[g]
    ids, nodes, first_cell = {}, [], []
    nc = len(cellnodes)
    ce = np.zeros((nc, len(table)), dtype=np.int64)
    sg = np.zeros_like(ce)
    ori = np.zeros_like(ce)
    fc = []
    for c in range(nc):
        for l, loc in enumerate(table):
            tup = tuple(int(cellnodes[c][i]) for i in loc)
            key = tuple(sorted(tup))
            created = key not in ids
            if created:
                ids[key] = len(nodes)
                nodes.append(tup)
                fc.append([c + 1, 0])
            e = ids[key]
            ce[c, l] = e + 1
            glob = nodes[e]
            if kind == "faces":
                sg[c, l] = 1 if created else -1
                if not created:
                    fc[e][1] = c + 1
                if len(tup) == 2:
                    ori[c, l] = 1 if created else 2
                else:
                    l1, l2, l3 = tup
                    ori[c, l] = {(l1, l2, l3): 1, (l3, l2, l1): 2, (l2, l1, l3): 3, (l1, l3, l2): 4}.get(glob, 0)
            else:
                sg[c, l] = 1 if tup[0] == glob[0] else -1
    out = {"cell_entities": ce, "entity_nodes": np.array(nodes, dtype=np.int64), "signs": sg}
    if kind == "faces":
        out["facecells"] = np.array(fc, dtype=np.int64)
        out["orientations"] = ori
    return out



# ---------------------------------------------------------------------------------------------
# interpolate!(target, FES, u) for H1P1 / H1P2 / H1Pk{.,.,<=2}   (Example210:168-173)
#   nodes:  NodalInterpolator, /root/reference/src/interpolators.jl:93-127  (point evaluation, components coffset apart)
#   edges:  MomentInterpolator(FES, ON_EDGES | ON_FACES) with order 0, /root/reference/src/fedefs/h1_p2.jl:47-76,
#           /root/reference/src/interpolators.jl:137-460: the interior dof of the edge is set so that the edge mean of u is
#           preserved; MOMxBASIS = means of the P2 edge basis (1/6, 1/6, 2/3), moments integrated with the Edge1D rule of
#           order 2 * order_FE (quadrature.jl:214-233 -> generic Gauss rule, :609-628), vertex contributions subtracted.
# ---------------------------------------------------------------------------------------------
def gauss_rule_edge(order: int):
    """QuadratureRule{T, Edge1D}(order) for order > 2: get_generic_quadrature_Gauss (quadrature.jl:609-628)."""
    ngpts = order // 2 + 1
    k = np.arange(1, ngpts, dtype=np.float64)
    gamma = k / np.sqrt(4.0 * k * k - 1.0)
    vals, vecs = np.linalg.eigh(np.diag(gamma, 1) + np.diag(gamma, -1))
    return 0.5 * vals + 0.5, 0.5 * (2.0 * vecs[0, :] ** 2)


def interpolate_h1_nodal_edge(coords, edgenodes, u, ncomp, order_fe=2, bonus_quadorder=0):
    """-> target of length ncomp * (nnodes + nedges) (component-blocked, nodes then edges per component), literal loops."""
    nn = coords.shape[0]
    ne = 0 if edgenodes is None else edgenodes.shape[0]
    coffset = nn + ne
    target = np.zeros(ncomp * coffset)
    for j in range(nn):  # interpolators.jl:111-124
        r = np.atleast_1d(u(coords[j]))
        for k in range(ncomp):
            target[j + k * coffset] = r[k]
    if ne:
        qx, qw = gauss_rule_edge(2 * order_fe + bonus_quadorder)
        mom_vertex, mom_interior = 1.0 / 6.0, 2.0 / 3.0  # MOMxBASIS of the P2 basis on the reference edge against the constant
        for e in range(ne):
            a, b = edgenodes[e, 0] - 1, edgenodes[e, 1] - 1
            f_mom = np.zeros(ncomp)
            for q in range(qx.shape[0]):  # eval_trafo! on the edge: x = x_a + xref (x_b - x_a)
                x = coords[a] + qx[q] * (coords[b] - coords[a])
                f_mom += qw[q] * np.atleast_1d(u(x))
            for k in range(ncomp):  # subtract the exterior (vertex) dofs, solve the 1 x 1 interior system
                f = f_mom[k] - mom_vertex * (target[a + k * coffset] + target[b + k * coffset])
                target[nn + e + k * coffset] = f / mom_interior
    return target
