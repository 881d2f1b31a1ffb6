"""FEEvaluator: host-side mirror of ``FEEvaluator(FES, operator, qf)``.

The reference constructor (/root/reference/src/feevaluator.jl:83-215) tabulates the reference
basis at all quadrature points (``refbasisvals``) and its reference derivatives via ForwardDiff
(``refbasisderivvals``, :233-291); ``update_basis!(FEB, cell)`` (:381-387 and
src/feevaluator_h1.jl / feevaluator_hdiv.jl) then fills ``FEB.cvals[r, dof, qp]`` cell by cell.

Here the constructor builds the same tables (closed-form polynomial derivatives, see fedefs.py);
the per-cell update is *batched* on the device: ``update_basis_all`` returns the ``cvals`` of a whole
cell range, and the assembly kernels consume the tables directly without ever materialising cvals
in HBM.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .fedefs import _DIM

Identity, Gradient, Divergence = "Identity", "Gradient", "Divergence"
_OPCODE = {Identity: _lib.OP_IDENTITY, Gradient: _lib.OP_GRADIENT, Divergence: _lib.OP_DIVERGENCE}


def Length4Operator(op, xdim, ncomponents):
    """functionoperators.jl:182-196"""
    if op == Identity:
        return ncomponents
    if op == Gradient:
        return xdim * ncomponents
    if op == Divergence:
        return -(-ncomponents // xdim)
    raise ValueError(op)


def NeededDerivative4Operator(op):
    return 0 if op == Identity else 1


class FEEvaluator:
    def __init__(self, fes, operator, qf):
        if operator not in _OPCODE:
            raise NotImplementedError(f"operator {operator} is not on the assembly path")
        self.FE, self.operator, self.qf = fes, operator, qf
        ft, g = fes.fetype, fes.xgrid.geometry
        self.edim = _DIM[g]
        self.ncomponents = ft.ncomponents
        self.ndofs4item = ft.get_ndofs(g)
        self.ndofs4item_all = ft.get_ndofs_all(g)
        self.resultdim = Length4Operator(operator, self.edim, self.ncomponents)
        vals, derivs = ft.tabulate(g, qf.xref)
        self.refbasisvals = vals      # [qp][dof_all][comp]
        self.refbasisderivs = derivs  # [qp][dof_all][comp][j]
        self.derivorder = NeededDerivative4Operator(operator)
        self.citem = 0
        self.cvals = np.zeros((self.resultdim, self.ndofs4item, len(qf)))  # Julia shape (resultdim, ndofs, nqp)
        if operator == Identity and ft.family not in ("Hdiv", "Hcurl") and ft.handler(g) == 0:
            # filled once in the constructor, never changes per cell (feevaluator.jl:158-161)
            self.cvals[:] = np.transpose(vals[:, : self.ndofs4item, :], (2, 1, 0))

    @property
    def refbasisderivvals(self):
        """Julia layout ``[dof_all + ndofs_all*(comp-1), j, qp]`` (feevaluator.jl:170-177)."""
        nq, nda, nc, d = self.refbasisderivs.shape
        return np.transpose(self.refbasisderivs, (2, 1, 3, 0)).reshape(nc * nda, d, nq)

    # -- device registration ------------------------------------------------
    def register(self, ctx, eval_id: int, space_id: int):
        ctx.evaluator_set(
            eval_id, space_id, _OPCODE[self.operator],
            self.refbasisvals if self.derivorder == 0 else None,
            self.refbasisderivs if self.derivorder == 1 else None,
        )
        self._ctx, self._eid = ctx, eval_id

    def update_basis_all(self, c0: int = 0, c1: int | None = None) -> np.ndarray:
        """Batched ``update_basis!`` on the device: ``out[cell, qp, dof, r] = cvals[r, dof, qp]``."""
        c1 = self.FE.xgrid.ncells if c1 is None else c1
        return self._ctx.evaluator_eval(self._eid, c0, c1, len(self.qf), self.ndofs4item, self.resultdim)

    def update_basis(self, cell: int):
        """``update_basis!(FEB, cell)`` for one 1-based cell (device round trip; for tests)."""
        if self.citem != cell:
            self.citem = cell
            self.cvals[:] = np.transpose(self.update_basis_all(cell - 1, cell)[0], (2, 1, 0))
        return self.cvals
