"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck): tiny meshes, all paths."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa
import femb_b200 as fb
from femb_b200 import _lib
from femb_b200.assembly import _Session

rhs2 = fb.AffineFunction([0.0], [[1.0, -1.0]])
rhs3 = fb.AffineFunction([0.25], [[1.0, -1.0, 0.5]])
X = np.linspace(0, 1, 12)
g2 = fb.jitter(fb.simplexgrid(X, X))
g3 = fb.jitter(fb.simplexgrid(X[:6], X[:6], X[:6]), 0.1)
for ft, g, rhs in [(fb.H1Pk(1, 2, 1), g2, rhs2), (fb.H1Pk(1, 2, 2), g2, rhs2), (fb.H1Pk(1, 2, 3), g2, rhs2), (fb.H1P1(1), g3, rhs3), (fb.H1P2(1, 3), g3, rhs3), (fb.H1P3(1, 3), g3, rhs3)]:
    for sc in (_lib.SCATTER_GATHER, _lib.SCATTER_ATOMIC):
        fes = fb.FESpace(ft, g)
        A, b = fb.FEMatrix(fes, fes), fb.FEVector(fes)
        fb.assemble_poisson(A.entries, b.entries, fes, rhs, 1.0, scatter=sc)
        fb.assemble_poisson(A.entries, b.entries, fes, rhs, 1.0, scatter=sc)
        print(ft, sc, A.entries.nnz, float(np.abs(A.entries.nzval).max()))
# streamed path
fes = fb.FESpace(fb.H1P1(1), g2)
S = _Session([fes], fb.QuadratureRule(g2.geometry, 0))
eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
nnz = S.ctx.pattern_build([(0, 0)])
nz, bb = np.zeros(nnz), np.zeros(fes.ndofs)
S.ctx.assemble_poisson_host(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, rhs2.params(), 1e-15, g2["Coordinates"], g2["CellNodes"], g2["CellVolumes"], nz, bb, 3)
print("streamed", nnz, float(np.abs(nz).max()))
# Example210
FESu, FESp = fb.FESpace(fb.H1Pk(2, 2, 2), g2), fb.FESpace(fb.H1Pk(1, 2, 1), g2)
sol = fb.FEVector([FESu, FESp]); sol.entries[:] = np.random.default_rng(0).uniform(-1, 1, sol.entries.shape[0])
A, b = fb.FEMatrix([FESu, FESp]), fb.FEVector([FESu, FESp])
upd = fb.prepare_assembly(A, b, FESu, FESp, sol)
upd(True, True)
print("TH", A.entries.nnz)
# Hdiv
for ft in (fb.HDIVRT0(3), fb.HDIVBDM1(3), fb.HDIVRT0(2), fb.HDIVBDM1(2)):
    g = g3 if ft.edim == 3 else g2
    FESv, FESq = fb.FESpace(ft, g), fb.FESpace(fb.L2P0(1), g)
    A = fb.FEMatrix([FESv, FESq])
    fb.assemble_hdiv_mixed(A, FESv, FESq)
    print(ft, A.entries.nnz)
# block-wise lazy zero: one block claimed, the others settled by the consumer; then the separate divergence kernel with claims
FESv, FESq = fb.FESpace(fb.HDIVRT0(3), g3), fb.FESpace(fb.L2P0(1), g3)
S = _Session([FESv, FESq], fb.QuadratureRule(g3.geometry, 2))
eid, ediv, eq = S.evaluator(0, "Identity"), S.evaluator(0, "Divergence"), S.evaluator(1, "Identity")
S.ctx.pattern_build([(0, 0), (0, 1), (1, 0)])
S.ctx.matrix_zero(); S.ctx.assemble_bilinear(eid._eid, eid._eid, 0, 0, 1.0, 0, 0.0)
print("claimed mass only", float(np.abs(S.ctx.matrix_get()).max()))
S.ctx.matrix_zero(); S.ctx.assemble_bilinear(ediv._eid, eq._eid, 0, 1, 1.0, _lib.BIL_TRANSPOSE_COPY, 0.0); S.ctx.assemble_bilinear(eid._eid, eid._eid, 0, 0, 1.0, 0, 0.0)
print("div then mass", float(np.abs(S.ctx.matrix_get()).max()))
S.ctx.close()
# Hcurl mass system
for ft, g in ((fb.HCURLN0(2), g2), (fb.HCURLN0(3), g3)):
    fes = fb.FESpace(ft, g)
    A, b = fb.ExtendableSparseMatrixCSC(fes.ndofs, fes.ndofs), np.zeros(fes.ndofs)
    fb.assemble_mass(A, b, fes, fb.AffineFunction([1.0] * ft.edim, [[0.0] * ft.edim] * ft.edim))
    print(ft, A.nnz, float(np.abs(b).max()))
# coloured scatter, device topology / dofmaps, boundary dofs, penalisation, residual, block products, compute_error
for ft in (fb.HDIVBDM1(3), fb.H1P3(1, 3)):
    FESv, FESq = fb.FESpace(ft, g3), fb.FESpace(fb.L2P0(1), g3)
    A = fb.FEMatrix([FESv, FESq])
    fb.assemble_hdiv_mixed(A, FESv, FESq, scatter=_lib.SCATTER_COLORED) if ft.family == "Hdiv" else None
for g in (g2, g3):
    dev = fb.ExtendableGrid(g["Coordinates"], g["CellNodes"], g.geometry).instantiate_on_device()
    print("entities", g.geometry, dev["FaceNodes"].shape[0], dev["EdgeNodes"].shape[0])
from femb_b200.fespace import parse_pattern
fes = fb.FESpace(fb.H1P2(1, 3), g3)
S = _Session([fes], fb.QuadratureRule(g3.geometry, 2), device_dofmaps=True)
eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
S.ctx.pattern_build([(0, 0)])
S.ctx.matrix_zero(); S.ctx.vector_zero()
S.ctx.assemble_poisson(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, rhs3.params(), 0.0)
nb = S.ctx.boundary_dofs(0, parse_pattern(fes.fetype.get_dofmap_pattern(g3.geometry)), download=False)
S.ctx.dirichlet_apply(True, 0, None, None, 1.0e60)
x = np.random.default_rng(1).uniform(-1, 1, fes.ndofs)
print("residual", nb, S.ctx.residual(x, zero_fixed=True), S.ctx.lrmatmul(x, x))
S.ctx.block_matmul(np.zeros(fes.ndofs), x, 0, 0, transposed=True)
S.ctx.system_save(); S.ctx.system_add_saved()
print("compute_error", float(fb.compute_error(fes, x, None, order=2).sum()))
print("done")
