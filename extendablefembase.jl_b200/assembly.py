"""Batched ``assemble!`` drop-ins for the low-level loops of the reference examples.

* ``assemble_poisson(A, b, fe_space, rhs, mu)``  <->  ``Example200.assemble!(A, b, fe_space, rhs, mu)``
  (/root/reference/examples/Example200_LowLevelPoisson.jl:103-184)
* ``prepare_assembly(A, b, FESu, FESp, sol)`` -> ``update_system(linear, nonlinear)``  <->
  ``Example210.prepare_assembly!`` (/root/reference/examples/Example210_LowLevelNavierStokes.jl:218-394)
* ``assemble_hdiv_mixed`` : HDIVRT0/HDIVBDM1 mass + divergence blocks against L2P0 (config C4).
* ``assemble_mass`` : mass matrix / load vector of one space with the Identity push-forward of its family (H1, L2, Hdiv, Hcurl).

All arithmetic happens in libfemb.so (hand-written sm_100a CUDA) through the C ABI of
include/femb.h; this module only extracts the raw arrays from the FESpace / FEEvaluator /
QuadratureRule / FEMatrix / FEVector objects — exactly what julia/FEMBaseB200.jl does with
``ccall``.  There is no CPU path.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .feevaluator import FEEvaluator, Gradient, Identity, Divergence
from .fematrix import FEMatrix, FEVector
from .quadrature import QuadratureRule


class AffineFunction:
    """``f_c(x) = c0[c] + sum_d g[c,d] x_d`` — evaluated on the device (FEMB_RHS_AFFINE).
    Example200's default ``rhs = x -> x[1] - x[2]`` is ``AffineFunction([0.0], [[1.0, -1.0]])``."""

    def __init__(self, c0, grad):
        self.c0 = np.atleast_1d(np.asarray(c0, dtype=np.float64))
        self.g = np.atleast_2d(np.asarray(grad, dtype=np.float64))

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        out = self.c0 + np.tensordot(x, self.g.T, axes=([-1], [0]))
        return out[..., 0] if self.c0.shape[0] == 1 else out

    def params(self):
        return np.concatenate([np.concatenate([[self.c0[c]], self.g[c]]) for c in range(self.c0.shape[0])])


class Example210RHS:
    """``f!(fval, x, t)`` of Example210:44-48, evaluated on the device (FEMB_RHS_EX210)."""

    def __init__(self, mu, t=0.0):
        self.mu, self.t = float(mu), float(t)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        e = 8.0 * np.pi * np.pi * self.mu * np.exp(-8.0 * np.pi * np.pi * self.mu * self.t)
        return np.stack([e * np.sin(2 * np.pi * x[..., 0]) * np.sin(2 * np.pi * x[..., 1]), e * np.cos(2 * np.pi * x[..., 0]) * np.cos(2 * np.pi * x[..., 1])], axis=-1)

    def params(self):
        return np.array([self.mu, self.t])


def _register_space(ctx, sid, fes, device_dofmap=False):
    """``device_dofmap=True``: faces / edges, signs and CellDofs are generated on the device from the dofmap pattern
    (femb_space_set_from_pattern) — nothing but the pattern string crosses the boundary."""
    ft, grid = fes.fetype, fes.xgrid
    g = grid.geometry
    handler = ft.handler(g)
    if device_dofmap:
        from .fespace import parse_pattern

        ndofs, coffset = ctx.space_set_from_pattern(sid, _lib.FAMILY[ft.family], ft.ncomponents, ft.get_ndofs(g), ft.get_ndofs_all(g), handler,
                                                    parse_pattern(ft.get_dofmap_pattern(g)))
        if (ndofs, coffset) != (fes.ndofs, fes.coffset):
            raise RuntimeError(f"device dofmap: ndofs / coffset {ndofs, coffset} differ from the FESpace's {fes.ndofs, fes.coffset}")
        return
    signs = orient = None
    if handler in (1, 3, 4, 5):
        signs = grid["CellFaceSigns"]
    elif handler in (2, 6):
        signs = grid["CellEdgeSigns"]
    if handler == 5:
        orient = grid["CellFaceOrientations"]
    if ft.family == "H1" and ft.ncomponents == 1 and ft.order == 1:
        celldofs = None  # CellDofs of scalar P1 is CellNodes itself (pattern "N1"): no second copy in HBM
    else:
        celldofs = fes["CellDofs"]
    ctx.space_set(sid, _lib.FAMILY[ft.family], ft.ncomponents, fes.ndofs, ft.get_ndofs(g), ft.get_ndofs_all(g), handler, celldofs, signs, orient)


def _rhs_args(ctx, rhs, nqp, ncomp):
    """-> (kind, data) for the C ABI.  Closures are evaluated on the host at the physical
    quadrature points the device reports (eval_trafo!, Example200:171)."""
    if rhs is None:
        return _lib.RHS_NONE, None
    if isinstance(rhs, AffineFunction):
        return _lib.RHS_AFFINE, rhs.params()
    if isinstance(rhs, Example210RHS):
        return _lib.RHS_EX210, rhs.params()
    x = ctx.qp_coords(nqp)  # (ncells, nqp, dim)
    f = np.asarray(rhs(x), dtype=np.float64)
    return _lib.RHS_QPVALUES, np.ascontiguousarray(f.reshape(x.shape[0], nqp, ncomp))


class _Session:
    """Device-resident state for one (grid, FE spaces, quadrature rule) combination."""

    def __init__(self, spaces, qf, device=0, cellvolumes=True, device_dofmaps=False):
        """``cellvolumes=False`` passes NULL for xgrid[CellVolumes]: the kernels then use |det A| / d!.
        ``device_dofmaps=True`` generates CellDofs (and the face / edge components behind them) on the device."""
        self.ctx = _lib.Context(device)
        grid = spaces[0].xgrid
        self.grid, self.spaces, self.qf = grid, spaces, qf
        self.ctx.mesh_set(grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"] if cellvolumes else None)
        for sid, fes in enumerate(spaces):
            _register_space(self.ctx, sid, fes, device_dofmaps)
        self.ctx.quadrature_set(qf.xref, qf.w)
        self.ctx.matrix_layout(list(range(len(spaces))), list(range(len(spaces))))
        self.evals = {}
        self.have_pattern = False

    def evaluator(self, sid, op):
        key = (sid, op)
        if key not in self.evals:
            feb = FEEvaluator(self.spaces[sid], op, self.qf)
            feb.register(self.ctx, len(self.evals), sid)
            self.evals[key] = feb
        return self.evals[key]

    def ensure_pattern(self, A, blocks, extra_diag=None, track=False):
        """Adopt A's flushed pattern if it has one, else build the structural pattern on the device."""
        ctx = self.ctx
        if A.nnz > 0:
            if not self.have_pattern or self._adopted_nnz != A.nnz:
                ctx.pattern_adopt(A.m, A.n, A.colptr, A.rowval, blocks)
                self.have_pattern, self._adopted_nnz = True, A.nnz
            return False
        ctx.pattern_track(track)
        ctx.pattern_build(blocks, extra_diag)
        self.have_pattern, self._adopted_nnz = True, -1
        return True

    def download_pattern(self, A):
        colptr, rowval = self.ctx.pattern_get(A.n)
        A.set_csc(colptr, rowval, np.zeros(self.ctx.nnz))
        self._adopted_nnz = A.nnz


def _session_for(spaces, qf, device):
    key = "_femb_session_%d" % device
    holder = spaces[0]
    s = getattr(holder, key, None)
    if s is None or s.spaces != list(spaces) or not np.array_equal(s.qf.xref, qf.xref):
        s = _Session(list(spaces), qf, device)
        setattr(holder, key, s)
    return s


# --------------------------------------------------------------------------------------
# Example200
# --------------------------------------------------------------------------------------
def assemble_poisson(A, b, fe_space, rhs, mu=1.0, *, drop_tol=1.0e-15, device=0, scatter=None):
    """Drop-in for ``Example200.assemble!(A::ExtendableSparseMatrix, b::Vector, fe_space, rhs, mu)``.

    Adds the stiffness matrix ``mu (grad u, grad v)`` to ``A`` (upper triangle computed, mirrored,
    contributions with ``|Aloc[j,k]| <= 1e-15`` skipped — Example200:139-161) and the load vector
    ``(rhs, v)`` to ``b`` (Example200:165-178).  First call on an empty ``A``: the pattern is built
    on the device and compressed by the same value filter; later calls reuse ``A``'s pattern
    (the reference's allocation-free second assembly, Example200:205-211).
    """
    ft = fe_space.fetype
    g = fe_space.xgrid.geometry
    qf = QuadratureRule(g, 2 * (ft.get_polynomialorder(g) - 1))  # Example200:110
    S = _session_for([fe_space], qf, device)
    ctx = S.ctx
    if scatter is not None:
        ctx.set_scatter(scatter)
    eg, ei = S.evaluator(0, Gradient), S.evaluator(0, Identity)
    fresh = S.ensure_pattern(A, [(0, 0)], track=True)
    kind, data = _rhs_args(ctx, rhs, len(qf), ft.ncomponents)
    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg._eid, ei._eid, float(mu), kind, data, drop_tol)
    if fresh:
        ctx.pattern_compress()
        S.download_pattern(A)
    ctx.matrix_add_to(A.nzval)
    b += ctx.vector_get(b.shape[0])
    return 0  # loop_allocations of the reference


# --------------------------------------------------------------------------------------
# Example210
# --------------------------------------------------------------------------------------
def prepare_assembly(A: FEMatrix, b: FEVector, FESu, FESp, sol: FEVector, *, teval=0.0, mu=1.0e-2, f=None, device=0):
    """Drop-in for ``Example210.prepare_assembly!(A, b, FESu, FESp, sol; teval)``; returns
    ``update_system(linear, nonlinear)`` (Example210:389-393).  ``f`` defaults to the example's
    ``f!`` (Example210:44-48)."""
    Ae, be, sole = A.entries, b.entries, sol.entries
    g = FESu.xgrid.geometry
    qf = QuadratureRule(g, 3 * FESu.fetype.get_polynomialorder(g) - 1)  # Example210:237
    S = _session_for([FESu, FESp], qf, device)
    ctx = S.ctx
    egu, eiu, eip = S.evaluator(0, Gradient), S.evaluator(0, Identity), S.evaluator(1, Identity)
    rhs = Example210RHS(mu, teval) if f is None else f
    blocks = [(0, 0), (0, 1), (1, 0)]

    def update_system(linear: bool, nonlinear: bool):
        fresh = S.ensure_pattern(Ae, blocks, extra_diag=[FESu.ndofs + 1])  # pressure pin, Example210:172
        if fresh:
            S.download_pattern(Ae)
        kind, data = _rhs_args(ctx, rhs if linear else None, len(qf), 2)
        ctx.vector_set(1, sole)
        ctx.matrix_zero()
        ctx.vector_zero()
        ctx.assemble_navierstokes(egu._eid, eiu._eid, eip._eid, mu, linear, nonlinear, kind, data)
        ctx.matrix_add_to(Ae.nzval)
        np.add(be, ctx.vector_get(be.shape[0]), out=be)

    update_system.session = S
    return update_system


# --------------------------------------------------------------------------------------
# Hdiv mixed blocks (config C4): mass (v,w) and divergence (div v, q) against L2P0
# --------------------------------------------------------------------------------------
def assemble_hdiv_mixed(A: FEMatrix, FESv, FESq, *, device=0, scatter=None):
    """``A[1,1] += (v_i, v_j)``, ``A[1,2] += (div v_i, q_k)`` and its transpose in ``A[2,1]``, with
    contravariant-Piola evaluators (feevaluator_hdiv.jl:2-19,66-83) and the 4-point (3-D) /
    3-point (2-D) order-2 rule."""
    Ae = A.entries
    g = FESv.xgrid.geometry
    qf = QuadratureRule(g, 2)
    S = _session_for([FESv, FESq], qf, device)
    ctx = S.ctx
    if scatter is not None:
        ctx.set_scatter(scatter)
    eid, ediv, eq = S.evaluator(0, Identity), S.evaluator(0, Divergence), S.evaluator(1, Identity)
    fresh = S.ensure_pattern(Ae, [(0, 0), (0, 1), (1, 0)])
    if fresh:
        S.download_pattern(Ae)
    ctx.matrix_zero()
    ctx.assemble_hdiv_mixed(eid._eid, ediv._eid, eq._eid, 1.0, 1.0)
    ctx.matrix_add_to(Ae.nzval)
    return S


# --------------------------------------------------------------------------------------
# L2 mass system of one space (every family: H1 / L2 / Hdiv / Hcurl Identity push-forwards)
# --------------------------------------------------------------------------------------
def assemble_mass(A, b, fes, rhs=None, *, quadorder=None, device=0, scatter=None):
    """``A += (u_j, u_k)`` and, with ``rhs``, ``b += (rhs, u_j)`` on the cells of ``fes``: the same loop as Example200:129-178 with the
    Identity evaluator on both sides (feevaluator_h1.jl:2-14, feevaluator_hdiv.jl:2-19, feevaluator_hcurl.jl:2-16).  ``A`` is an
    ``ExtendableSparseMatrix`` (pattern built on the device on first use), ``b`` a vector of length ``fes.ndofs``."""
    ft = fes.fetype
    g = fes.xgrid.geometry
    qf = QuadratureRule(g, 2 * ft.get_polynomialorder(g) if quadorder is None else quadorder)
    S = _session_for([fes], qf, device)
    ctx = S.ctx
    if scatter is not None:
        ctx.set_scatter(scatter)
    ei = S.evaluator(0, Identity)
    fresh = S.ensure_pattern(A, [(0, 0)])
    if fresh:
        S.download_pattern(A)
    ctx.matrix_zero()
    ctx.assemble_bilinear(ei._eid, ei._eid, 0, 0, 1.0, 0, 0.0)
    ctx.matrix_add_to(A.nzval)
    if rhs is not None:
        kind, data = _rhs_args(ctx, rhs, len(qf), ei.resultdim)
        ctx.vector_zero()
        ctx.assemble_linear(ei._eid, 0, 1.0, kind, data)
        b += ctx.vector_get(b.shape[0])
    return S


# --------------------------------------------------------------------------------------
# compute_error (Example210:110-154)
# --------------------------------------------------------------------------------------
def compute_error(fes, coefficients, u=None, order=None, p=2, *, operator=Identity, device=0):
    """Drop-in for ``compute_error(uh::FEVectorBlock, u, order, p)``: ``error[cell, d] = sum_qp (uh_d - u_d)^p |T| w_qp`` over the
    ``VertexRule`` of the given order (default: the polynomial order of the element), evaluated on the device.  ``u`` maps the
    physical points ``x[..., dim]`` to values ``[..., ncomp]`` (or is None).  Returns ``(ncells, ncomp)`` (the bytes of Julia's
    ``ncomp x ncells``)."""
    from .quadrature import VertexRule

    g = fes.xgrid.geometry
    if order is None:
        order = fes.fetype.get_polynomialorder(g)
    qf = VertexRule(g, order)
    S = _Session([fes], qf, device)
    try:
        ev = S.evaluator(0, operator)
        exact = None
        if u is not None:
            x = S.ctx.qp_coords(len(qf))
            exact = np.ascontiguousarray(np.asarray(u(x), dtype=np.float64).reshape(x.shape[0], len(qf), ev.resultdim))
        return S.ctx.eval_error(ev._eid, np.asarray(coefficients, dtype=np.float64), exact, int(p), ev.resultdim)
    finally:
        S.ctx.close()
