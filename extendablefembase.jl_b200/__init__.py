"""extendablefembase.jl_b200 — B200-native batched cell-wise FE assembly behind the API surface of
WIAS-PDELib/ExtendableFEMBase.jl (FESpace, CellDofs, QuadratureRule, FEEvaluator, FEMatrix, FEVector).

Only the assembly hot path is here: the host mirror of the reference interface in
this package, and ``csrc/`` = hand-written sm_100a CUDA kernels behind the C ABI of ``include/femb.h``
(``libfemb.so``).  Import name: ``femb_b200`` (the directory name contains a dot; ``femb_b200.py`` at
the repo root is the import shim).
"""
from . import _lib
from .assembly import AffineFunction, Example210RHS, assemble_hdiv_mixed, assemble_mass, assemble_poisson, compute_error, prepare_assembly
from .fedefs import H1P1, H1P2, H1P3, H1Pk, HCURLN0, HDIVBDM1, HDIVRT0, L2P0
from .feevaluator import Divergence, FEEvaluator, Gradient, Identity
from .fematrix import ExtendableSparseMatrixCSC, FEMatrix, FEVector
from .fespace import FESpace, boundarydofs
from .grids import ExtendableGrid, grid_tetrahedron, grid_triangle, jitter, simplexgrid
from .quadrature import QuadratureRule, VertexRule

CellDofs = "CellDofs"

__all__ = [
    "AffineFunction", "Example210RHS", "assemble_hdiv_mixed", "assemble_mass", "assemble_poisson", "compute_error", "prepare_assembly", "VertexRule",
    "H1P1", "H1P2", "H1P3", "H1Pk", "HCURLN0", "HDIVBDM1", "HDIVRT0", "L2P0",
    "Divergence", "FEEvaluator", "Gradient", "Identity",
    "ExtendableSparseMatrixCSC", "FEMatrix", "FEVector", "FESpace", "boundarydofs", "CellDofs",
    "ExtendableGrid", "grid_tetrahedron", "grid_triangle", "jitter", "simplexgrid", "QuadratureRule",
]
