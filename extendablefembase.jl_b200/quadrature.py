"""QuadratureRule tables for the simplex geometries on the assembly path.

Mirrors ``QuadratureRule{T,EG}(order)`` of the reference
(/root/reference/src/quadrature.jl:257-279 triangle, :352-409 tetrahedron,
:609-665 generic Gauss / Stroud conical product).  Only the *tables*
(``xref``, ``w``) are on the hot path: they are staged in shared memory by the
CUDA kernels.  Weights sum to 1 on the reference cell, so every integral is
``|T| * sum_qp w_qp (...)`` (quadrature.jl:262,269,368).
"""
from __future__ import annotations

import numpy as np


class QuadratureRule:
    """``qf.xref`` is ``(nqp, dim)`` float64, ``qf.w`` is ``(nqp,)`` float64."""

    def __init__(self, geometry: str, order: int | None = None, *, xref=None, w=None, name="N.N."):
        self.geometry = geometry
        if xref is not None:
            self.xref = np.ascontiguousarray(xref, dtype=np.float64)
            self.w = np.ascontiguousarray(w, dtype=np.float64)
            self.name = name
        elif geometry == "Triangle2D":
            self.xref, self.w, self.name = _triangle(int(order))
        elif geometry == "Tetrahedron3D":
            self.xref, self.w, self.name = _tetrahedron(int(order))
        else:
            raise ValueError(f"no quadrature rule for {geometry}")
        self.dim = self.xref.shape[1]

    def __len__(self):
        return self.w.shape[0]

    def __repr__(self):
        return f"QuadratureRule({self.geometry}, {self.name!r}, npoints={len(self)})"


_TRI_EDGES = ((0, 1), (1, 2), (2, 0))
_TET_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
_TET_FACES = ((0, 2, 1), (0, 1, 3), (1, 2, 3), (0, 3, 2))


def VertexRule(geometry: str, order: int = 1) -> QuadratureRule:
    """``VertexRule(EG, order)`` (/root/reference/src/quadrature.jl:103-138 triangle, :157-186 tetrahedron): the Lagrange
    points of the given order with equal weights; the rule of ``compute_error`` (Example210:122)."""
    from fractions import Fraction as Fr

    d = 2 if geometry == "Triangle2D" else 3
    if order == 0:
        pts = [[Fr(1, d + 1)] * d]
    else:
        pts = [[Fr(0)] * d] + [[Fr(int(i == j)) for i in range(d)] for j in range(d)]
    for a, b in (_TRI_EDGES if d == 2 else _TET_EDGES):
        for j in range(1, order):
            pts.append([Fr(j, order) * pts[b][i] + (1 - Fr(j, order)) * pts[a][i] for i in range(d)])
    if d == 2:
        interior = {3: [(1, 1)], 4: [(1, 1), (2, 1), (1, 2)], 5: [(1, 1), (3, 1), (1, 3), (2, 1), (2, 2), (1, 2)]}.get(order, [])
        pts += [[Fr(x, order), Fr(y, order)] for x, y in interior]
        if order > 5:
            raise NotImplementedError("VertexRule for order > 5 on Triangle2D (quadrature.jl:133-135)")
    else:
        if order > 2:
            for j in range(1, order - 1):
                for k in range(1, order - 1):
                    for f in _TET_FACES:
                        pts.append([Fr(j, order) * pts[f[2]][i] + Fr(k, order) * pts[f[1]][i] + (1 - Fr(j, order) - Fr(k, order)) * pts[f[0]][i] for i in range(d)])
        if order > 3:
            raise NotImplementedError("VertexRule for order > 3 on Tetrahedron3D (quadrature.jl:181-183)")
    xref = np.array([[float(c) for c in p] for p in pts])
    w = np.full(len(pts), 1.0 / len(pts))
    return QuadratureRule(geometry, xref=xref, w=w, name="vertex rule")


def _triangle(order: int):
    # quadrature.jl:257-279
    if order <= 1:
        return np.array([[1 / 3, 1 / 3]]), np.array([1.0]), "midpoint rule"
    if order == 2:
        x = np.array([[0.5, 0.5], [0.0, 0.5], [0.5, 0.0]])
        return x, np.full(3, 1 / 3), "face midpoints rule"
    if order == 8 or 12 <= order <= 14:
        raise NotImplementedError("symmetric Zhang/Cui/Liu rules (orders 8, 12-14) are not on the assembly path")
    x, w = stroud_conical(order)
    return x, w, f"generic Stroud rule of order {order}"


def _tetrahedron(order: int):
    # quadrature.jl:352-409
    if order <= 1:
        return np.array([[0.25, 0.25, 0.25]]), np.array([1.0]), "midpoint rule"
    if order == 2:
        a, b = 0.1381966011250105, 0.5854101966249685
        x = np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]])
        return x, np.full(4, 0.25), "order 2 rule"
    if order <= 3:
        x = np.array([[1 / 4, 1 / 4, 1 / 4], [1 / 2, 1 / 6, 1 / 6], [1 / 6, 1 / 6, 1 / 6], [1 / 6, 1 / 6, 1 / 2], [1 / 6, 1 / 2, 1 / 6]])
        w = np.array([-4 / 5, 9 / 20, 9 / 20, 9 / 20, 9 / 20])
        return x, w, "order 3 rule"
    if order <= 4:
        a, b = 0.7857142857142857, 0.0714285714285714
        c, d = 0.1005964238332008, 0.3994035761667992
        x = np.array(
            [[0.25, 0.25, 0.25], [a, b, b], [b, b, b], [b, b, a], [b, a, b], [c, d, d], [d, c, d], [d, d, c], [d, c, c], [c, d, c], [c, c, d]]
        )
        w = np.array([-0.0789333333333333] + [0.0457333333333333] * 4 + [0.1493333333333333] * 6)
        return x, w, "order 4 rule"
    raise NotImplementedError("symmetric tetrahedron rules of order > 4 are not on the assembly path")


def gauss_jacobi_01(ngpts: int):
    """1-D Gauss points/weights on [0,1] via Golub-Welsch (quadrature.jl:609-628)."""
    k = np.arange(1, ngpts, dtype=np.float64)
    gamma = k / np.sqrt(4 * k**2 - 1)
    vals, vecs = np.linalg.eigh(np.diag(gamma, 1) + np.diag(gamma, -1))
    return 0.5 * vals + 0.5, 0.5 * 2 * vecs[0, :] ** 2


def stroud_conical(order: int):
    """Stroud conical product rule on the reference triangle (quadrature.jl:631-665):
    ``xref = (s_j, r_i (1 - s_j))``, ``w = a_i b_j``, point index ``i + ngpts*(j-1)``."""
    n = order // 2 + 1
    r, a = gauss_jacobi_01(n)
    k = np.arange(1, n + 1, dtype=np.float64)
    delta = -1.0 / (4 * k**2 - 1)
    kk = np.arange(1, n, dtype=np.float64)
    gamma = np.sqrt((kk + 1) * kk) / (2 * (kk + 1) - 1)
    vals, vecs = np.linalg.eigh(np.diag(delta) + np.diag(gamma, 1) + np.diag(gamma, -1))
    s = 0.5 * vals + 0.5
    b = 0.5 * 2 * vecs[0, :] ** 2
    # s = repeat(s', ngpts, 1)[:]  -> s varies slowest; r = repeat(r, ngpts) -> r varies fastest
    S = np.repeat(s, n)
    R = np.tile(r, n)
    xref = np.stack([S, -R * (S - 1.0)], axis=1)
    w = np.outer(a, b).reshape(-1, order="F")  # (a' * b)[:] column-major: i fastest
    return xref, w
