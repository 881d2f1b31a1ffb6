# Parity harness a maintainer runs inside the reference's environment (needs julia, ExtendableFEMBase and a B200 with
# libfemb.so built: `python __graft_entry__.py build`).  It compares the B200 drop-ins with the reference's own
# hand-written loops on the reference's own test set-ups:
#   * Example200.runtests (examples/Example200_LowLevelPoisson.jl:192-212): H1Pk{1,2,order} on simplexgrid(LinRange(0,1,64))²
#   * Example210.prepare_assembly! (examples/Example210_LowLevelNavierStokes.jl:218-394): linear and nonlinear update
#   * boundarydofs / penalisation / residual (Example200:86-92, Example210:172-209)
# NOT executed in the build container (no julia binary there): the same comparisons run in tests/test_gpu_parity.py against
# the CPU restatement of these loops (oracle/), call for call.  Bars: pattern and dof indices identical, values 1e-12
# relative to the largest local contribution.
#
#   julia --project=<ExtendableFEMBase checkout> julia/runtests_b200.jl
using Test
using LinearAlgebra
using SparseArrays
using ExtendableFEMBase
using ExtendableGrids
using ExtendableSparse

const REF = get(ENV, "EXTENDABLEFEMBASE_DIR", joinpath(dirname(pathof(ExtendableFEMBase)), ".."))
include(joinpath(REF, "examples", "Example200_LowLevelPoisson.jl"))
include(joinpath(REF, "examples", "Example210_LowLevelNavierStokes.jl"))
include(joinpath(@__DIR__, "FEMBaseB200.jl"))

const RTOL = 1.0e-12

function compare_csc(A::SparseMatrixCSC, B::SparseMatrixCSC, scale)
    @test A.colptr == B.colptr
    @test A.rowval == B.rowval
    @test maximum(abs.(A.nzval .- B.nzval)) <= RTOL * scale
end

@testset "Example200 assemble! order $order" for order in (1, 2, 3)
    FEType = H1Pk{1, 2, order}
    X = LinRange(0, 1, 64)                                   # Example200:199
    xgrid = simplexgrid(X, X)
    fe_space = FESpace{FEType}(xgrid)
    rhs = x -> x[1] - x[2]                                   # Example200:38
    mu = 1.0
    # the reference's loop
    A_ref = FEMatrix(fe_space, fe_space)
    b_ref = FEVector(fe_space)
    Example200.assemble!(A_ref.entries, b_ref.entries, fe_space, rhs, mu)
    flush!(A_ref.entries)
    # the B200 drop-in, first assembly (pattern built on the device) ...
    A = FEMatrix(fe_space, fe_space)
    b = FEVector(fe_space)
    FEMBaseB200.assemble!(A.entries, b.entries, fe_space, rhs, mu)
    scale = maximum(abs.(A_ref.entries.cscmatrix.nzval))
    compare_csc(A.entries.cscmatrix, A_ref.entries.cscmatrix, scale)
    @test maximum(abs.(b.entries .- b_ref.entries)) <= RTOL * maximum(abs.(b_ref.entries))
    # ... and the second assembly into the same matrix (pattern adopted; the reference ADDS, Example200:206-211)
    Example200.assemble!(A_ref.entries, b_ref.entries, fe_space, rhs, mu)
    FEMBaseB200.assemble!(A.entries, b.entries, fe_space, rhs, mu)
    compare_csc(A.entries.cscmatrix, A_ref.entries.cscmatrix, 2 * scale)
    @test issymmetric(A.entries.cscmatrix)                   # owner-computes gather: exactly symmetric
end

@testset "Example210 update_system!" begin
    X = LinRange(0, 1, 2^3 + 1)
    xgrid = simplexgrid(X, X)
    FES = [FESpace{H1Pk{2, 2, 2}}(xgrid), FESpace{H1Pk{1, 2, 1}}(xgrid)]
    sol = FEVector(FES)
    sol.entries .= sin.(1:length(sol.entries))               # any linearisation point
    for (linear, nonlinear) in ((true, false), (false, true), (true, true))
        A_ref, b_ref = FEMatrix(FES), FEVector(FES)
        update_ref! = Example210.prepare_assembly!(A_ref, b_ref, FES[1], FES[2], sol)
        update_ref!(linear, nonlinear)
        A, b = FEMatrix(FES), FEVector(FES)
        update! = FEMBaseB200.prepare_assembly!(A, b, FES[1], FES[2], sol; teval = 0, mu = Example210.μ)
        update!(linear, nonlinear)
        scale = maximum(abs.(A_ref.entries.cscmatrix.nzval))
        compare_csc(A.entries.cscmatrix, A_ref.entries.cscmatrix, scale)
        @test maximum(abs.(b.entries .- b_ref.entries)) <= RTOL * max(maximum(abs.(b_ref.entries)), scale)
    end
end

@testset "grid components, CellDofs, boundarydofs, penalisation, residual on the device" begin
    X = LinRange(0, 1, 9)
    for (xgrid, FEType) in ((simplexgrid(X, X), H1Pk{1, 2, 3}), (simplexgrid(X, X, X), H1P2{1, 3}), (simplexgrid(X, X, X), HDIVBDM1{3}))
        EG = xgrid[UniqueCellGeometries][1]
        FES = FESpace{FEType}(xgrid)
        ctx = FEMBaseB200.Context(0)
        FEMBaseB200.mesh_set!(ctx, xgrid)
        f = FEMBaseB200.grid_entities!(ctx, xgrid, 0)
        @test f.cell_ent == xgrid[CellFaces]
        @test f.ent_nodes == xgrid[FaceNodes]
        @test f.signs == xgrid[CellFaceSigns]
        @test f.facecells == xgrid[FaceCells]
        if dim_element(EG) == 3
            @test f.orient == xgrid[CellFaceOrientations]
            e = FEMBaseB200.grid_entities!(ctx, xgrid, 1)
            @test e.cell_ent == xgrid[CellEdges]
            @test e.ent_nodes == xgrid[EdgeNodes]
            @test e.signs == xgrid[CellEdgeSigns]
        end
        ndofs, _ = FEMBaseB200.space_set_from_pattern!(ctx, 0, FEType, EG)
        @test ndofs == FES.ndofs
        nd = get_ndofs(ON_CELLS, FEType, EG)
        celldofs = FES[CellDofs]
        @test FEMBaseB200.celldofs(ctx, 0, nd, num_cells(xgrid)) == reshape(celldofs.colentries, nd, :)
        if FEType <: ExtendableFEMBase.AbstractH1FiniteElement
            @test FEMBaseB200.boundarydofs(ctx, 0, FEType, EG) == sort(unique(boundarydofs(FES)))
        end
    end
end
