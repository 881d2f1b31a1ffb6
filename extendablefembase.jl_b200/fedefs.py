"""Finite element definitions on the assembly path.

Per FEType (mirroring /root/reference/src/fedefs/*.jl): ``ncomponents``, ``ndofs``
/ ``ndofs_all`` per cell, polynomial order, the CellDofs pattern string, the
reference basis, and which per-cell *handlers* apply (coefficient handler = face
signs, subset handler = orientation-dependent basis selection).

The reference evaluates basis closures in floating point and differentiates them
with ForwardDiff (src/feevaluator.jl:233-291).  Here every reference basis
function is held as an exact multivariate polynomial with rational coefficients
(``Poly``), so values *and* derivatives are tabulated from closed forms; the
tables agree with ForwardDiff to rounding.  The tables, not the closures, are what
the CUDA kernels consume (staged in shared memory).
"""
from __future__ import annotations

from fractions import Fraction as Fr
from functools import lru_cache

import numpy as np


# ----------------------------------------------------------------------
# exact multivariate polynomials
# ----------------------------------------------------------------------
class Poly:
    __slots__ = ("d", "c")

    def __init__(self, d: int, c: dict | None = None):
        self.d = d
        self.c = {k: Fr(v) for k, v in (c or {}).items() if v != 0}

    @staticmethod
    def const(d, v):
        return Poly(d, {(0,) * d: Fr(v)})

    @staticmethod
    def var(d, i):
        e = [0] * d
        e[i] = 1
        return Poly(d, {tuple(e): Fr(1)})

    def _coerce(self, o):
        return o if isinstance(o, Poly) else Poly.const(self.d, o)

    def __add__(self, o):
        o = self._coerce(o)
        c = dict(self.c)
        for k, v in o.c.items():
            c[k] = c.get(k, 0) + v
        return Poly(self.d, c)

    __radd__ = __add__

    def __neg__(self):
        return Poly(self.d, {k: -v for k, v in self.c.items()})

    def __sub__(self, o):
        return self + (-self._coerce(o))

    def __rsub__(self, o):
        return self._coerce(o) - self

    def __mul__(self, o):
        o = self._coerce(o)
        c: dict = {}
        for k1, v1 in self.c.items():
            for k2, v2 in o.c.items():
                k = tuple(a + b for a, b in zip(k1, k2))
                c[k] = c.get(k, 0) + v1 * v2
        return Poly(self.d, c)

    __rmul__ = __mul__

    def __truediv__(self, s):
        return Poly(self.d, {k: v / Fr(s) for k, v in self.c.items()})

    def deriv(self, i: int) -> "Poly":
        c = {}
        for k, v in self.c.items():
            if k[i] > 0:
                e = list(k)
                e[i] -= 1
                c[tuple(e)] = c.get(tuple(e), 0) + v * k[i]
        return Poly(self.d, c)

    def __call__(self, x):
        """Evaluate at points ``x`` of shape ``(..., d)`` in float64, or exactly
        at a tuple of Fractions."""
        if isinstance(x, (tuple, list)) and isinstance(x[0], (Fr, int)):
            tot = Fr(0)
            for k, v in self.c.items():
                t = v
                for xi, e in zip(x, k):
                    t *= Fr(xi) ** e
                tot += t
            return tot
        x = np.asarray(x, dtype=np.float64)
        out = np.zeros(x.shape[:-1])
        for k, v in self.c.items():
            t = np.full(x.shape[:-1], float(v))
            for i, e in enumerate(k):
                if e:
                    t = t * x[..., i] ** e
            out = out + t
        return out

    def is_zero(self):
        return not self.c


def _bary(d):
    xs = [Poly.var(d, i) for i in range(d)]
    lam0 = Poly.const(d, 1)
    for x in xs:
        lam0 = lam0 - x
    return lam0, xs


# ----------------------------------------------------------------------
# scalar reference bases (one component); vector-valued H1 spaces replicate them
# component-blocked: local dof (c-1)*nd + j lives in component c only
# (h1_p2.jl:133-141, h1_pk.jl:271-274).
# ----------------------------------------------------------------------
def _h1_pk_triangle(order: int):
    """h1_pk.jl:170-276 (orders 1..5)."""
    lam0, (x, y) = _bary(2)
    coeffs = [Fr(k, order) for k in range(order + 1)]
    factor_node = Fr(1)
    for c in coeffs[1:]:
        factor_node *= c
    factors_face = [Fr(1)] * (order - 1)
    for j in range(2, order + 1):
        for k in range(1, order + 2):
            if k < j:
                factors_face[j - 2] *= coeffs[j - 1] - coeffs[k - 1]
            elif k > j:
                factors_face[j - 2] *= coeffs[k - 1] - coeffs[j - 1]
    basis = []
    for v in (lam0, x, y):
        p = Poly.const(2, 1)
        for k in range(order):
            p = p * (v - coeffs[k])
        basis.append(p / factor_node)
    faces = [(lam0, x), (x, y), (y, lam0)]  # (first factor, second factor) per face [1,2],[2,3],[3,1]
    for fa, fb in faces:
        for k in range(1, order):
            p = fa * fb / factors_face[k - 1]
            if order > 2:
                for m in range(1, order):
                    if m > k:
                        p = p * (fa - (1 - coeffs[m]))
                    elif m < k:
                        p = p * (fb - coeffs[m])
            basis.append(p)
    bub = lam0 * x * y
    if order == 3:
        basis.append(bub * 27)
    elif order == 4:
        basis += [bub * (lam0 - Fr(1, 4)) * 128, bub * (x - Fr(1, 4)) * 128, bub * (y - Fr(1, 4)) * 128]
    elif order == 5:
        a, b = Fr(1, 5), Fr(2, 5)
        basis += [
            bub * (lam0 - a) * (lam0 - b) * Fr(3125, 6),
            bub * (x - a) * (x - b) * Fr(3125, 6),
            bub * (y - a) * (y - b) * Fr(3125, 6),
            bub * (lam0 - a) * (x - a) * Fr(3125, 4),
            bub * (x - a) * (y - a) * Fr(3125, 4),
            bub * (y - a) * (lam0 - a) * Fr(3125, 4),
        ]
    elif order > 5:
        raise NotImplementedError("H1Pk order > 5")
    return basis


def _h1_p1(dim):
    lam0, xs = _bary(dim)  # h1_p1.jl:64-76
    return [lam0] + xs


def _h1_p2(dim):
    lam0, xs = _bary(dim)
    h = Fr(1, 2)
    if dim == 2:  # h1_p2.jl:130-144
        x, y = xs
        return [2 * lam0 * (lam0 - h), 2 * x * (x - h), 2 * y * (y - h), 4 * lam0 * x, 4 * x * y, 4 * y * lam0]
    x, y, z = xs  # h1_p2.jl:146-163
    return [
        2 * lam0 * (lam0 - h), 2 * x * (x - h), 2 * y * (y - h), 2 * z * (z - h),
        4 * lam0 * x, 4 * lam0 * y, 4 * lam0 * z, 4 * x * y, 4 * x * z, 4 * y * z,
    ]


def _h1_p3(dim):
    lam0, xs = _bary(dim)
    t, tt = Fr(1, 3), Fr(2, 3)
    n = Fr(9, 2)
    e = Fr(27, 2)
    if dim == 2:  # h1_p3.jl:142-160
        x, y = xs
        return [
            n * lam0 * (lam0 - t) * (lam0 - tt), n * x * (x - t) * (x - tt), n * y * (y - t) * (y - tt),
            e * x * lam0 * (lam0 - t), e * x * lam0 * (x - t),
            e * y * x * (x - t), e * y * x * (y - t),
            e * lam0 * y * (y - t), e * lam0 * y * (lam0 - t),
            27 * x * y * lam0,
        ]
    x, y, z = xs  # h1_p3.jl:163-197
    return [
        n * lam0 * (lam0 - t) * (lam0 - tt), n * x * (x - t) * (x - tt), n * y * (y - t) * (y - tt), n * z * (z - t) * (z - tt),
        e * x * lam0 * (lam0 - t), e * x * lam0 * (x - t),
        e * y * lam0 * (lam0 - t), e * y * lam0 * (y - t),
        e * z * lam0 * (lam0 - t), e * z * lam0 * (z - t),
        e * x * y * (x - t), e * x * y * (y - t),
        e * x * z * (x - t), e * x * z * (z - t),
        e * y * z * (y - t), e * y * z * (z - t),
        27 * x * y * lam0, 27 * x * z * lam0, 27 * x * y * z, 27 * y * z * lam0,
    ]


def _vec(d, comps):
    return [c if isinstance(c, Poly) else Poly.const(d, c) for c in comps]


def _hdiv_rt0(dim):
    if dim == 2:  # hdiv_rt0.jl:60-69
        _, (x, y) = _bary(2)
        return [_vec(2, [x, y - 1]), _vec(2, [x, y]), _vec(2, [x - 1, y])]
    _, (x, y, z) = _bary(3)  # hdiv_rt0.jl:84-100
    return [_vec(3, [2 * x, 2 * y, 2 * (z - 1)]), _vec(3, [2 * x, 2 * (y - 1), 2 * z]), _vec(3, [2 * x, 2 * y, 2 * z]), _vec(3, [2 * (x - 1), 2 * y, 2 * z])]


def _hcurl_n0(dim):
    if dim == 2:  # hcurl_n0.jl:94-104
        _, (x, y) = _bary(2)
        return [_vec(2, [1 - y, x]), _vec(2, [-y, x]), _vec(2, [-y, x - 1])]
    _, (x, y, z) = _bary(3)  # hcurl_n0.jl:120-142 (local edges [1 2; 1 3; 1 4; 2 3; 2 4; 3 4])
    return [_vec(3, [1 - y - z, x, x]), _vec(3, [y, 1 - z - x, y]), _vec(3, [z, z, 1 - x - y]),
            _vec(3, [-y, x, 0]), _vec(3, [-z, 0, x]), _vec(3, [0, -z, y])]


def _hdiv_bdm1(dim):
    if dim == 2:  # hdiv_bdm1.jl:73-90
        _, (x, y) = _bary(2)
        return [
            _vec(2, [x, y - 1]), _vec(2, [6 * x, 6 - 12 * x - 6 * y]),
            _vec(2, [x, y]), _vec(2, [-6 * x, 6 * y]),
            _vec(2, [x - 1, y]), _vec(2, [6 * (x - 1) + 12 * y, -6 * y]),
        ]
    l, (x, y, z) = _bary(3)  # hdiv_bdm1.jl:123-197 (16 functions, 12 selected per cell)
    return [
        _vec(3, [2 * x, 2 * y, 2 * (z - 1)]),
        _vec(3, [24 * x, 0, 24 * (l - x)]), _vec(3, [0, -24 * y, -24 * (l - y)]), _vec(3, [-24 * x, 24 * y, -24 * (y - x)]),
        _vec(3, [2 * x, 2 * (y - 1), 2 * z]),
        _vec(3, [0, 24 * (l - z), 24 * z]), _vec(3, [-24 * x, -24 * (l - x), 0]), _vec(3, [24 * x, -24 * (x - z), -24 * z]),
        _vec(3, [2 * x, 2 * y, 2 * z]),
        _vec(3, [-24 * x, 0, 24 * z]), _vec(3, [24 * x, -24 * y, 0]), _vec(3, [0, 24 * y, -24 * z]),
        _vec(3, [2 * (x - 1), 2 * y, 2 * z]),
        _vec(3, [24 * (l - y), 24 * y, 0]), _vec(3, [-24 * (l - z), 0, -24 * z]), _vec(3, [-24 * (z - y), -24 * y, 24 * z]),
    ]


# ----------------------------------------------------------------------
# FEType descriptors
# ----------------------------------------------------------------------
HANDLER_NONE = 0
HANDLER_H1_FACESWAP2D = 1  # h1_p3.jl:200-217 / h1_pk.jl:280-304: reverse the face dofs where CellFaceSigns != 1
HANDLER_H1_EDGESWAP3D = 2  # h1_p3.jl:220-241: same with CellEdgeSigns on tets
HANDLER_RT0_SIGNS = 3      # hdiv_rt0.jl:114-124
HANDLER_BDM1_2D = 4        # hdiv_bdm1.jl:200-212
HANDLER_BDM1_3D = 5        # hdiv_bdm1.jl:215-249 (signs + orientation subset)
HANDLER_N0_EDGESIGNS3D = 6  # hcurl_n0.jl:156-166 (CellEdgeSigns)

_DIM = {"Triangle2D": 2, "Tetrahedron3D": 3}
_NNODES = {"Triangle2D": 3, "Tetrahedron3D": 4}
_NFACES = {"Triangle2D": 3, "Tetrahedron3D": 4}
_NEDGES = {"Triangle2D": 3, "Tetrahedron3D": 6}


class FEType:
    """Value-level stand-in for the reference's FE *types* (``H1P1{1}``, ``H1Pk{2,2,2}`` ...)."""

    family = "H1"  # "H1" | "Hdiv" | "L2"
    handler2d = HANDLER_NONE
    handler3d = HANDLER_NONE

    def __init__(self, name, ncomponents, edim=None, order=1):
        self.name, self.ncomponents, self.edim, self.order = name, int(ncomponents), edim, int(order)

    def __repr__(self):
        return self.name

    def __eq__(self, o):
        return isinstance(o, FEType) and self.name == o.name

    def __hash__(self):
        return hash(self.name)

    # -- interface mirrored from fedefs/*.jl ------------------------------
    def get_ncomponents(self):
        return self.ncomponents

    def get_polynomialorder(self, geometry):
        return self.order

    def get_dofmap_pattern(self, geometry) -> str:
        raise NotImplementedError

    def scalar_basis(self, geometry):
        raise NotImplementedError

    def get_ndofs(self, geometry) -> int:
        return len(self.reference_basis(geometry)[0]) if self.family != "H1" else len(self.scalar_basis(geometry)) * self.ncomponents

    def get_ndofs_all(self, geometry) -> int:
        return self.get_ndofs(geometry)

    def handler(self, geometry) -> int:
        return self.handler2d if _DIM[geometry] == 2 else self.handler3d

    def reference_basis(self, geometry):
        """-> (list over dof_all of list over component of Poly)."""
        sb = self.scalar_basis(geometry)
        d = _DIM[geometry]
        zero = Poly(d)
        out = []
        for c in range(self.ncomponents):
            for p in sb:
                out.append([p if k == c else zero for k in range(self.ncomponents)])
        return (out,)

    # -- tables -----------------------------------------------------------
    def tabulate(self, geometry, xref: np.ndarray):
        """``vals[qp, dof_all, comp]`` and ``derivs[qp, dof_all, comp, j]`` =
        d(phi_dof,comp)/d(xref_j): the contents of ``refbasisvals`` /
        ``refbasisderivvals`` (src/feevaluator.jl:123-126, :170-177)."""
        (basis,) = self.reference_basis(geometry)
        d = _DIM[geometry]
        nq, nda, nc = xref.shape[0], len(basis), self.ncomponents
        vals = np.zeros((nq, nda, nc))
        derivs = np.zeros((nq, nda, nc, d))
        for i, comps in enumerate(basis):
            for c, p in enumerate(comps):
                if p.is_zero():
                    continue
                vals[:, i, c] = p(xref)
                for j in range(d):
                    dp = p.deriv(j)
                    if not dp.is_zero():
                        derivs[:, i, c, j] = dp(xref)
        return vals, derivs


class _H1P1(FEType):
    def get_dofmap_pattern(self, geometry):
        return "N1"

    def scalar_basis(self, geometry):
        return _h1_p1(_DIM[geometry])


class _H1P2(FEType):
    def get_dofmap_pattern(self, geometry):
        return "N1F1" if _DIM[geometry] == 2 else "N1E1"

    def scalar_basis(self, geometry):
        return _h1_p2(_DIM[geometry])


class _H1P3(FEType):
    handler2d = HANDLER_H1_FACESWAP2D
    handler3d = HANDLER_H1_EDGESWAP3D

    def get_dofmap_pattern(self, geometry):
        return "N1F2I1" if _DIM[geometry] == 2 else "N1E2F1"

    def scalar_basis(self, geometry):
        return _h1_p3(_DIM[geometry])


class _H1Pk(FEType):
    def __init__(self, name, ncomponents, edim, order):
        super().__init__(name, ncomponents, edim, order)
        self.handler2d = HANDLER_H1_FACESWAP2D if order >= 3 else HANDLER_NONE

    def get_dofmap_pattern(self, geometry):
        o = self.order
        if _DIM[geometry] != 2:
            raise NotImplementedError("H1Pk is defined on Edge1D / Triangle2D only (h1_pk.jl:40-42)")
        if o == 1:
            return "N1"
        if o == 2:
            return f"N1F{o - 1}"
        return f"N1F{o - 1}I{(o - 2) * (o - 1) // 2}"

    def scalar_basis(self, geometry):
        return _h1_pk_triangle(self.order)


class _L2P0(FEType):
    family = "L2"

    def get_dofmap_pattern(self, geometry):
        return "I1"

    def get_polynomialorder(self, geometry):
        return 0

    def scalar_basis(self, geometry):
        return [Poly.const(_DIM[geometry], 1)]  # l2_p0.jl:65-72

    def get_ndofs(self, geometry):
        return self.ncomponents


class _HDIVRT0(FEType):
    family = "Hdiv"
    handler2d = HANDLER_RT0_SIGNS
    handler3d = HANDLER_RT0_SIGNS

    def get_dofmap_pattern(self, geometry):
        return "f1"

    def reference_basis(self, geometry):
        return (_hdiv_rt0(_DIM[geometry]),)

    def get_ndofs(self, geometry):
        return _NFACES[geometry]


class _HCURLN0(FEType):
    """Lowest-order Nedelec space of the first kind (hcurl_n0.jl): one tangential flux per face (2-D) / edge (3-D)."""

    family = "Hcurl"
    handler2d = HANDLER_RT0_SIGNS  # CellFaceSigns on every function (hcurl_n0.jl:144-154), as for RT0
    handler3d = HANDLER_N0_EDGESIGNS3D

    def get_dofmap_pattern(self, geometry):
        return "f1" if _DIM[geometry] == 2 else "e1"  # hcurl_n0.jl:31-34

    def reference_basis(self, geometry):
        return (_hcurl_n0(_DIM[geometry]),)

    def get_ndofs(self, geometry):
        return 3 if _DIM[geometry] == 2 else 6


class _HDIVBDM1(FEType):
    family = "Hdiv"
    handler2d = HANDLER_BDM1_2D
    handler3d = HANDLER_BDM1_3D

    def get_dofmap_pattern(self, geometry):
        return "f2" if _DIM[geometry] == 2 else "f3"

    def reference_basis(self, geometry):
        return (_hdiv_bdm1(_DIM[geometry]),)

    def get_ndofs(self, geometry):
        return _DIM[geometry] * _NFACES[geometry]

    def get_ndofs_all(self, geometry):
        return 2 * _NFACES[geometry] if _DIM[geometry] == 2 else 4 * _NFACES[geometry]  # hdiv_bdm1.jl:24


@lru_cache(maxsize=None)
def H1P1(ncomponents=1):
    return _H1P1(f"H1P1{{{ncomponents}}}", ncomponents, None, 1)


@lru_cache(maxsize=None)
def H1P2(ncomponents=1, edim=2):
    return _H1P2(f"H1P2{{{ncomponents},{edim}}}", ncomponents, edim, 2)


@lru_cache(maxsize=None)
def H1P3(ncomponents=1, edim=2):
    return _H1P3(f"H1P3{{{ncomponents},{edim}}}", ncomponents, edim, 3)


@lru_cache(maxsize=None)
def H1Pk(ncomponents=1, edim=2, order=2):
    return _H1Pk(f"H1Pk{{{ncomponents},{edim},{order}}}", ncomponents, edim, order)


@lru_cache(maxsize=None)
def L2P0(ncomponents=1):
    return _L2P0(f"L2P0{{{ncomponents}}}", ncomponents, None, 0)


@lru_cache(maxsize=None)
def HDIVRT0(edim=2):
    return _HDIVRT0(f"HDIVRT0{{{edim}}}", edim, edim, 1)


@lru_cache(maxsize=None)
def HCURLN0(edim=2):
    return _HCURLN0(f"HCURLN0{{{edim}}}", edim, edim, 1)


@lru_cache(maxsize=None)
def HDIVBDM1(edim=2):
    return _HDIVBDM1(f"HDIVBDM1{{{edim}}}", edim, edim, 1)
