# FEMBaseB200.jl — Julia host side of libfemb (hand-written sm_100a CUDA kernels behind the C ABI of include/femb.h).
#
# Keeps the reference's API surface (FESpace, CellDofs, QuadratureRule, FEEvaluator, FEMatrix, FEVector): the
# batched `assemble!` below has the signature of Example200's hand-written loop
# (ExtendableFEMBase.jl/examples/Example200_LowLevelPoisson.jl:103) and `prepare_assembly!` that of Example210
# (examples/Example210_LowLevelNavierStokes.jl:218), so they are drop-ins for those loops.  No CUDA.jl kernels, no
# CPU fallback: every call goes through `ccall` into libfemb.so.
#
# NOTE: this file cannot be executed in the build container (no julia binary); every `ccall` here has a ctypes
# twin in extendablefembase.jl_b200/_lib.py which IS exercised by the test-suite, argument for argument.
module FEMBaseB200

using ExtendableFEMBase
using ExtendableGrids
using ExtendableSparse
using SparseArrays

const libfemb = get(ENV, "LIBFEMB", joinpath(@__DIR__, "..", "extendablefembase.jl_b200", "libfemb.so"))

# enums of include/femb.h
const FAMILY_H1, FAMILY_HDIV, FAMILY_L2, FAMILY_HCURL = Int32(0), Int32(1), Int32(2), Int32(3)
const OP_IDENTITY, OP_GRADIENT, OP_DIVERGENCE = Int32(0), Int32(1), Int32(2)
const RHS_NONE, RHS_AFFINE, RHS_QPVALUES, RHS_EX210 = Int32(0), Int32(1), Int32(2), Int32(3)
const BIL_TRANSPOSE_COPY = Int32(1)

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:femb_create, libfemb), Int32, (Int32, Ref{Ptr{Cvoid}}), device, ref)
        st == 0 || error("femb_create: " * unsafe_string(ccall((:femb_last_error, libfemb), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(ref[])
        finalizer(c -> ccall((:femb_destroy, libfemb), Int32, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end

check(ctx::Context, st::Int32) = st == 0 || error("libfemb status $st: " * unsafe_string(ccall((:femb_last_error, libfemb), Cstring, (Ptr{Cvoid},), ctx.h)))

# ---- raw-array extraction: Julia arrays are passed AS IS (column-major, 1-based) ---------------------------------
function mesh_set!(ctx::Context, xgrid::ExtendableGrid; cellvolumes::Bool = true)
    coords = xgrid[Coordinates]                 # dim x nnodes Float64
    cellnodes = xgrid[CellNodes]                # (dim+1) x ncells Int32 (Matrix for single-geometry grids)
    vols = cellvolumes ? xgrid[CellVolumes] : nothing
    GC.@preserve coords cellnodes vols begin
        check(ctx, ccall((:femb_mesh_set, libfemb), Int32,
            (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Int32, Int64, Ptr{Int32}, Ptr{Float64}),
            ctx.h, size(coords, 1), size(coords, 2), coords, size(cellnodes, 1), size(cellnodes, 2), cellnodes,
            vols === nothing ? C_NULL : pointer(vols)))
    end
end

family(::Type{<:ExtendableFEMBase.AbstractH1FiniteElement}) = FAMILY_H1
family(::Type{<:ExtendableFEMBase.AbstractHdivFiniteElement}) = FAMILY_HDIV
family(::Type{<:ExtendableFEMBase.AbstractHcurlFiniteElement}) = FAMILY_HCURL
family(::Type{<:L2P0}) = FAMILY_L2

# per-cell handlers (get_basissubset / get_coefficients of src/fedefs/*.jl) -> FEMB_HANDLER_* of include/femb.h
function handler(FEType, EG)
    dim = dim_element(EG)
    FEType <: H1P3 && return dim == 2 ? Int32(1) : Int32(2)
    (FEType <: H1Pk && get_polynomialorder(FEType, EG) >= 3) && return Int32(1)
    FEType <: HDIVRT0 && return Int32(3)
    FEType <: HDIVBDM1 && return dim == 2 ? Int32(4) : Int32(5)
    FEType <: HCURLN0 && return dim == 2 ? Int32(3) : Int32(6)   # face signs (hcurl_n0.jl:144-154) / edge signs (:156-166)
    return Int32(0)
end

function space_set!(ctx::Context, id::Integer, FES::FESpace)
    FEType = eltype(FES)
    xgrid = FES.xgrid
    EG = xgrid[UniqueCellGeometries][1]
    celldofs = FES[CellDofs]                       # VariableTargetAdjacency{Int32}: colentries is nd x ncells, cell-contiguous
    nd = get_ndofs(ON_CELLS, FEType, EG)
    nd_all = get_ndofs_all(ON_CELLS, FEType, EG)
    h = handler(FEType, EG)
    signs = h in (1, 3, 4, 5) ? xgrid[CellFaceSigns] : (h in (2, 6) ? xgrid[CellEdgeSigns] : nothing)
    orient = h == 5 ? xgrid[CellFaceOrientations] : nothing
    entries = celldofs isa Matrix ? celldofs : celldofs.colentries
    # scalar P1: CellDofs is CellNodes (pattern "N1"); pass NULL so the library aliases the mesh array
    p_dofs = (FEType <: H1P1 || (FEType <: H1Pk && get_polynomialorder(FEType, EG) == 1)) && get_ncomponents(FEType) == 1 ? C_NULL : pointer(entries)
    GC.@preserve entries signs orient begin
        check(ctx, ccall((:femb_space_set, libfemb), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int32, Int64, Int32, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
            ctx.h, id, family(FEType), get_ncomponents(FEType), FES.ndofs, nd, nd_all, h, p_dofs,
            signs === nothing ? C_NULL : pointer(signs isa Matrix ? signs : signs.colentries),
            orient === nothing ? C_NULL : pointer(orient isa Matrix ? orient : orient.colentries)))
    end
end

# ---- grid components and dofmaps generated on the device (optional: skips the host instantiation of
# xgrid[CellFaces] / xgrid[CellEdges] and fe_space[CellDofs], timed at Example200:55,59) ----------------------------
const SEGCODE = Dict('n' => Int32(0), 'f' => Int32(1), 'e' => Int32(2), 'i' => Int32(3), 'c' => Int32(3))

# "N1E1" -> Int32[type, each_component, q, ...]  (the parser of src/dofmaps.jl:284-301)
function pattern_segments(pattern::String)
    segs = Int32[]
    for m in eachmatch(r"([NnFfEeIiCc])(\d+)", pattern)
        c = m.captures[1][1]
        append!(segs, (SEGCODE[lowercase(c)], Int32(isuppercase(c)), parse(Int32, m.captures[2])))
    end
    return segs
end

"""
    grid_entities!(ctx, xgrid, kind) -> NamedTuple

Face (kind = 0) or edge (kind = 1) enumeration of the mesh registered with `mesh_set!`, on the device; the result can be
stored into `xgrid[CellFaces]`, `xgrid[FaceNodes]`, `xgrid[CellFaceSigns]`, `xgrid[CellFaceOrientations]`,
`xgrid[FaceCells]` / `xgrid[CellEdges]`, `xgrid[EdgeNodes]`, `xgrid[CellEdgeSigns]`.
"""
function grid_entities!(ctx::Context, xgrid::ExtendableGrid, kind::Integer)
    n = Ref{Int64}(0)
    check(ctx, ccall((:femb_grid_entities, libfemb), Int32, (Ptr{Cvoid}, Int32, Ref{Int64}), ctx.h, kind, n))
    dim = size(xgrid[Coordinates], 1)
    ncells = num_cells(xgrid)
    faces = kind == 0 || dim == 2
    per = faces ? dim + 1 : 6
    k = (faces && dim == 3) ? 3 : 2
    cell_ent, ent_nodes, signs = zeros(Int32, per, ncells), zeros(Int32, k, n[]), zeros(Int32, per, ncells)
    orient = faces ? zeros(Int32, per, ncells) : nothing
    facecells = faces ? zeros(Int32, 2, n[]) : nothing
    check(ctx, ccall((:femb_grid_entities_get, libfemb), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
        ctx.h, kind, cell_ent, ent_nodes, signs, faces ? orient : C_NULL, faces ? facecells : C_NULL))
    return (; cell_ent, ent_nodes, signs, orient, facecells)
end

# fe_space[CellDofs] generated on the device from get_dofmap_pattern (src/dofmaps.jl:419-627); returns (ndofs, coffset)
function space_set_from_pattern!(ctx::Context, id::Integer, FEType, EG)
    segs = pattern_segments(ExtendableFEMBase.get_dofmap_pattern(FEType, CellDofs, EG))
    ndofs, coffset = Ref{Int64}(0), Ref{Int64}(0)
    GC.@preserve segs check(ctx, ccall((:femb_space_set_from_pattern, libfemb), Int32,
        (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Int32, Int32, Int32, Ptr{Int32}, Ref{Int64}, Ref{Int64}),
        ctx.h, id, family(FEType), get_ncomponents(FEType), get_ndofs(ON_CELLS, FEType, EG), get_ndofs_all(ON_CELLS, FEType, EG),
        handler(FEType, EG), length(segs) ÷ 3, segs, ndofs, coffset))
    return ndofs[], coffset[]
end

function celldofs(ctx::Context, id::Integer, nd::Integer, ncells::Integer)
    out = zeros(Int32, nd, ncells)
    check(ctx, ccall((:femb_space_get_celldofs, libfemb), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}), ctx.h, id, out))
    return out
end

# ---- the steps after the assembly, on the device (Example200:86-92, Example210:172-173,180-185,192-209) -------------
# boundarydofs(FES) (src/dofmaps.jl:871-899), sorted and unique
function boundarydofs(ctx::Context, id::Integer, FEType, EG)
    segs = pattern_segments(ExtendableFEMBase.get_dofmap_pattern(FEType, CellDofs, EG))
    n = Ref{Int64}(0)
    GC.@preserve segs check(ctx, ccall((:femb_boundary_dofs, libfemb), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ref{Int64}), ctx.h, id, length(segs) ÷ 3, segs, n))
    dofs = zeros(Int32, n[])
    check(ctx, ccall((:femb_boundary_dofs_get, libfemb), Int32, (Ptr{Cvoid}, Ptr{Int32}), ctx.h, dofs))
    return dofs
end

# interpolate!(u_init[1], u!; time) on the device (node values + edge-mean preserving edge dofs); kind = :affine (params per
# component c0, g...) or :ex210_u (params [mu, t]); the result stays on the device as the boundary data of dirichlet!(values = nothing)
function interpolate_device!(ctx::Context, id::Integer, FES; kind = :ex210_u, params = Float64[], only_boundary = true, offset = 0, quadorder = 4)
    FEType = eltype(FES)
    EG = FES.xgrid[UniqueCellGeometries][1]
    segs = pattern_segments(ExtendableFEMBase.get_dofmap_pattern(FEType, CellDofs, EG))
    qf = QuadratureRule{Float64, Edge1D}(quadorder)
    qx = Float64[x[1] for x in qf.xref]
    qw = Vector{Float64}(qf.w)
    target = zeros(Float64, FES.ndofs)
    GC.@preserve segs params qx qw target check(ctx, ccall((:femb_interpolate, libfemb), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Int64, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}),
        ctx.h, id, length(segs) ÷ 3, segs, offset, kind == :affine ? 1 : 2, params, length(qw), qx, qw, only_boundary ? 1 : 0, target))
    return target
end

# A[dof,dof] = penalty; b[dof] = penalty * values[dof] (values = nothing: the device-resident data of interpolate_device!) for the boundary dofs of the last boundarydofs call (+ offset) and `extra`
function dirichlet!(ctx::Context; offset = 0, extra = Int64[], values = nothing, penalty = 1.0e60)
    GC.@preserve extra values check(ctx, ccall((:femb_dirichlet_apply, libfemb), Int32,
        (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Float64),
        ctx.h, 1, offset, length(extra), extra, values === nothing ? C_NULL : pointer(values), penalty))
end

# norm(A * x - b) with the fixed rows zeroed (x = nothing: the vector of the last femb_vector_set(1, ..))
function residual_norm(ctx::Context, x = nothing)
    nrm = Ref{Float64}(0)
    GC.@preserve x check(ctx, ccall((:femb_residual, libfemb), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}, Ref{Float64}),
        ctx.h, x === nothing ? C_NULL : pointer(x), 1, C_NULL, nrm))
    return nrm[]
end

# addblock_matmul!(a, B[brow, bcol], b; factor, transposed) (src/fematrix.jl:486-563), block numbers 0-based in the layout
function block_matmul!(ctx::Context, a::Vector{Float64}, b::Vector{Float64}; brow = -1, bcol = -1, factor = 1.0, transposed = false)
    GC.@preserve a b check(ctx, ccall((:femb_block_matmul, libfemb), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Float64, Int32, Ptr{Float64}),
        ctx.h, brow, bcol, b, factor, transposed, a))
    return a
end

# lrmatmul(a, A, b) / ldrdmatmul(a1, a2, A, b1, b2) (src/fematrix.jl:607-637)
function lrmatmul(ctx::Context, a1, b1; a2 = nothing, b2 = nothing, factor = 1.0)
    out = Ref{Float64}(0)
    GC.@preserve a1 a2 b1 b2 check(ctx, ccall((:femb_lrmatmul, libfemb), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Float64}),
        ctx.h, a1, a2 === nothing ? C_NULL : pointer(a2), b1, b2 === nothing ? C_NULL : pointer(b2), factor, out))
    return out[]
end

# compute_error(uh, u, order, p) (Example210:110-154): exact = values of u at the quadrature points (resultdim x nqp x ncells) or nothing
function eval_error(ctx::Context, eval_id::Integer, coefficients::Vector{Float64}, exact, p::Integer, resultdim::Integer, ncells::Integer)
    err = zeros(Float64, resultdim, ncells)
    GC.@preserve coefficients exact check(ctx, ccall((:femb_eval_error, libfemb), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}),
        ctx.h, eval_id, coefficients, exact === nothing ? C_NULL : pointer(exact), p, err))
    return err
end

system_save!(ctx::Context) = check(ctx, ccall((:femb_system_save, libfemb), Int32, (Ptr{Cvoid},), ctx.h))            # Alin = deepcopy(A); blin = deepcopy(b)
system_add_saved!(ctx::Context) = check(ctx, ccall((:femb_system_add_saved, libfemb), Int32, (Ptr{Cvoid},), ctx.h))  # A += Alin; b += blin

function quadrature_set!(ctx::Context, qf::QuadratureRule)
    xref = reduce(hcat, qf.xref)                 # dim x nqp
    w = collect(Float64, qf.w)
    GC.@preserve xref w check(ctx, ccall((:femb_quadrature_set, libfemb), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), ctx.h, length(w), xref, w))
end

# FEEvaluator tables -> C layout: refvals [qp][dof_all][comp], refderivs [qp][dof_all][comp][j]
# (Julia: FEB.refbasisvals[qp][dof_all, comp]; FEB.refbasisderivvals[dof_all + ndofs_all*(comp-1), j, qp], feevaluator.jl:123-126,170-177)
function evaluator_set!(ctx::Context, id::Integer, space_id::Integer, FEB, op::Int32)
    nqp = length(FEB.xref)
    nda, nc = size(FEB.refbasisvals[1])
    vals = Array{Float64}(undef, nc, nda, nqp)
    for qp in 1:nqp, d in 1:nda, c in 1:nc
        vals[c, d, qp] = FEB.refbasisvals[qp][d, c]
    end
    derivs = nothing
    if FEB.derivorder > 0
        edim = size(FEB.refbasisderivvals, 2)
        derivs = Array{Float64}(undef, edim, nc, nda, nqp)
        for qp in 1:nqp, d in 1:nda, c in 1:nc, j in 1:edim
            derivs[j, c, d, qp] = FEB.refbasisderivvals[d + nda * (c - 1), j, qp]
        end
    end
    GC.@preserve vals derivs check(ctx, ccall((:femb_evaluator_set, libfemb), Int32,
        (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}),
        ctx.h, id, space_id, op, op == OP_IDENTITY ? pointer(vals) : C_NULL, derivs === nothing ? C_NULL : pointer(derivs)))
end

# ---- device session cached per FESpace (the grid and the symbolic phase stay resident between assemblies) --------
mutable struct Session
    ctx::Context
    have_pattern::Bool
    nnz::Int64
end
const SESSIONS = IdDict{Any, Session}()

affine_params(c0::Real, g::AbstractVector) = Float64[c0; g]

"""
    assemble!(A::ExtendableSparseMatrix, b::Vector, fe_space, rhs, mu = 1; rhs_affine = nothing)

Drop-in for `Example200.assemble!` (Example200_LowLevelPoisson.jl:103-184).  `rhs` may be any Julia function of
`x` (evaluated on the host at the physical quadrature points the device reports, FEMB_RHS_QPVALUES) or, via
`rhs_affine = (c0, g)`, the affine function `c0 + g⋅x` evaluated on the device (Example200's default
`x -> x[1] - x[2]` is `(0.0, [1.0, -1.0])`).  Returns 0 (the reference returns its loop allocations).
"""
function assemble!(A::ExtendableSparseMatrix, b::Vector{Float64}, fe_space::FESpace, rhs, mu = 1; rhs_affine = nothing, device = 0)
    xgrid = fe_space.xgrid
    EG = xgrid[UniqueCellGeometries][1]
    FEType = eltype(fe_space)
    qf = QuadratureRule{Float64, EG}(2 * (get_polynomialorder(FEType, EG) - 1))        # Example200:110
    S = get!(SESSIONS, fe_space) do
        ctx = Context(device)
        mesh_set!(ctx, xgrid)
        space_set!(ctx, 0, fe_space)
        quadrature_set!(ctx, qf)
        evaluator_set!(ctx, 0, 0, FEEvaluator(fe_space, Gradient, qf), OP_GRADIENT)       # Example200:117
        evaluator_set!(ctx, 1, 0, FEEvaluator(fe_space, Identity, qf), OP_IDENTITY)       # Example200:119
        sp = Int32[0]
        check(ctx, ccall((:femb_matrix_layout, libfemb), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}), ctx.h, 1, sp, 1, sp))
        Session(ctx, false, 0)
    end
    ctx = S.ctx
    flush!(A)
    csc = A.cscmatrix
    blocks = Int32[0]
    fresh = false
    if nnz(csc) > 0 && (!S.have_pattern || S.nnz != nnz(csc))
        # steady state of the reference (second assembly, Example200:205-211): adopt the flushed pattern
        check(ctx, ccall((:femb_pattern_adopt, libfemb), Int32,
            (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int32}, Ptr{Int32}),
            ctx.h, size(csc, 1), size(csc, 2), nnz(csc), csc.colptr, csc.rowval, 1, blocks, blocks))
        S.have_pattern = true; S.nnz = nnz(csc)
    elseif nnz(csc) == 0
        fresh = true
        nnzref = Ref{Int64}(0)
        check(ctx, ccall((:femb_pattern_track, libfemb), Int32, (Ptr{Cvoid}, Int32), ctx.h, 1))
        check(ctx, ccall((:femb_pattern_build, libfemb), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{Int64}, Ref{Int64}), ctx.h, 1, blocks, blocks, 0, C_NULL, nnzref))
    end
    kind, data = RHS_NONE, Float64[]
    if rhs_affine !== nothing
        kind, data = RHS_AFFINE, affine_params(rhs_affine[1], rhs_affine[2])
    elseif rhs !== nothing
        nqp = length(qf.w); dim = size(xgrid[Coordinates], 1); nc = num_cells(xgrid)
        xq = Array{Float64}(undef, dim, nqp, nc)
        check(ctx, ccall((:femb_qp_coords, libfemb), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, xq))
        kind, data = RHS_QPVALUES, Float64[rhs(view(xq, :, qp, cell)) for qp in 1:nqp, cell in 1:nc][:]
    end
    check(ctx, ccall((:femb_matrix_zero, libfemb), Int32, (Ptr{Cvoid},), ctx.h))
    check(ctx, ccall((:femb_vector_zero, libfemb), Int32, (Ptr{Cvoid},), ctx.h))
    GC.@preserve data check(ctx, ccall((:femb_assemble_poisson, libfemb), Int32,
        (Ptr{Cvoid}, Int32, Int32, Float64, Int32, Ptr{Float64}, Float64), ctx.h, 0, 1, Float64(mu), kind, isempty(data) ? C_NULL : pointer(data), 1.0e-15))
    if fresh
        # Example200's |Aloc| > 1e-15 filter decides the pattern (Example200:153): compress, then hand the CSC to Julia
        nnzref = Ref{Int64}(0)
        check(ctx, ccall((:femb_pattern_compress, libfemb), Int32, (Ptr{Cvoid}, Ref{Int64}), ctx.h, nnzref))
        n = fe_space.ndofs
        colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nnzref[])
        check(ctx, ccall((:femb_pattern_get, libfemb), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), ctx.h, colptr, rowval))
        A.cscmatrix = SparseMatrixCSC{Float64, Int64}(n, n, colptr, rowval, zeros(Float64, nnzref[]))
        S.have_pattern = true; S.nnz = nnzref[]
    end
    check(ctx, ccall((:femb_matrix_add_to, libfemb), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, A.cscmatrix.nzval))   # A += assembled
    bdev = Vector{Float64}(undef, length(b))
    check(ctx, ccall((:femb_vector_get, libfemb), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, bdev))
    b .+= bdev
    return 0
end

"""
    prepare_assembly!(A, b, FESu, FESp, sol; teval = 0, mu = 1e-2) -> update_system!(linear, nonlinear)

Drop-in for `Example210.prepare_assembly!` (Example210_LowLevelNavierStokes.jl:218-394): Laplacian,
div-pressure coupling and rhs (linear part), Newton-linearised convection and its rhs (nonlinear part, around
`sol`).  The example's `f!` is evaluated on the device (FEMB_RHS_EX210).
"""
function prepare_assembly!(A, b, FESu, FESp, sol; teval = 0, mu = 1.0e-2, device = 0)
    Ae, be, sole = A.entries, b.entries, sol.entries
    xgrid = FESu.xgrid
    EG = xgrid[UniqueCellGeometries][1]
    qf = QuadratureRule{Float64, EG}(3 * get_polynomialorder(eltype(FESu), EG) - 1)    # Example210:237
    ctx = Context(device)
    mesh_set!(ctx, xgrid)
    space_set!(ctx, 0, FESu); space_set!(ctx, 1, FESp)
    quadrature_set!(ctx, qf)
    evaluator_set!(ctx, 0, 0, FEEvaluator(FESu, Gradient, qf), OP_GRADIENT)
    evaluator_set!(ctx, 1, 0, FEEvaluator(FESu, Identity, qf), OP_IDENTITY)
    evaluator_set!(ctx, 2, 1, FEEvaluator(FESp, Identity, qf), OP_IDENTITY)
    sp = Int32[0, 1]
    check(ctx, ccall((:femb_matrix_layout, libfemb), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}), ctx.h, 2, sp, 2, sp))
    brow, bcol = Int32[0, 0, 1], Int32[0, 1, 0]
    have_pattern = false
    function update_system!(linear::Bool, nonlinear::Bool)
        flush!(Ae)
        csc = Ae.cscmatrix
        if !have_pattern
            if nnz(csc) == 0
                nnzref = Ref{Int64}(0); pin = Int64[FESu.ndofs + 1]                    # pressure pin, Example210:172
                check(ctx, ccall((:femb_pattern_build, libfemb), Int32,
                    (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{Int64}, Ref{Int64}), ctx.h, 3, brow, bcol, 1, pin, nnzref))
                n = size(csc, 1)
                colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nnzref[])
                check(ctx, ccall((:femb_pattern_get, libfemb), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), ctx.h, colptr, rowval))
                Ae.cscmatrix = SparseMatrixCSC{Float64, Int64}(n, n, colptr, rowval, zeros(Float64, nnzref[]))
            else
                check(ctx, ccall((:femb_pattern_adopt, libfemb), Int32,
                    (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int32}, Ptr{Int32}),
                    ctx.h, size(csc, 1), size(csc, 2), nnz(csc), csc.colptr, csc.rowval, 3, brow, bcol))
            end
            have_pattern = true
        end
        params = Float64[mu, teval]
        check(ctx, ccall((:femb_vector_set, libfemb), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), ctx.h, 1, sole))   # current solution
        check(ctx, ccall((:femb_matrix_zero, libfemb), Int32, (Ptr{Cvoid},), ctx.h))
        check(ctx, ccall((:femb_vector_zero, libfemb), Int32, (Ptr{Cvoid},), ctx.h))
        check(ctx, ccall((:femb_assemble_navierstokes, libfemb), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int32, Float64, Int32, Int32, Int32, Ptr{Float64}),
            ctx.h, 0, 1, 2, Float64(mu), linear, nonlinear, linear ? RHS_EX210 : RHS_NONE, params))
        check(ctx, ccall((:femb_matrix_add_to, libfemb), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, Ae.cscmatrix.nzval))
        bdev = Vector{Float64}(undef, length(be))
        check(ctx, ccall((:femb_vector_get, libfemb), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, bdev))
        be .+= bdev
        return nothing
    end
    return update_system!
end

"""
    assemble_hdiv_mixed!(A::FEMatrix, FESv, FESq; device = 0)

Blocks of a mixed system over `[FESv; FESq]` (`FESv` an Hdiv space such as `HDIVRT0{3}`, `FESq = L2P0{1}`): `A[1,1] += (v_j, v_k)`,
`A[1,2] += (div v_j, q)`, `A[2,1] +=` its transpose — the cell loop of Example200:129-161 with the Identity / Divergence evaluators of
src/feevaluator_hdiv.jl (order-2 rule).  One `femb_assemble_hdiv_mixed` call; for HDIVRT0 on tetrahedra a fused write-once pass.
"""
function assemble_hdiv_mixed!(A, FESv, FESq; device = 0)
    Ae = A.entries
    xgrid = FESv.xgrid
    EG = xgrid[UniqueCellGeometries][1]
    qf = QuadratureRule{Float64, EG}(2)
    ctx = Context(device)
    mesh_set!(ctx, xgrid)
    space_set!(ctx, 0, FESv); space_set!(ctx, 1, FESq)
    quadrature_set!(ctx, qf)
    evaluator_set!(ctx, 0, 0, FEEvaluator(FESv, Identity, qf), OP_IDENTITY)
    evaluator_set!(ctx, 1, 0, FEEvaluator(FESv, Divergence, qf), OP_DIVERGENCE)
    evaluator_set!(ctx, 2, 1, FEEvaluator(FESq, Identity, qf), OP_IDENTITY)
    sp = Int32[0, 1]
    check(ctx, ccall((:femb_matrix_layout, libfemb), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}), ctx.h, 2, sp, 2, sp))
    brow, bcol = Int32[0, 0, 1], Int32[0, 1, 0]
    flush!(Ae)
    csc = Ae.cscmatrix
    if nnz(csc) == 0
        nnzref = Ref{Int64}(0)
        check(ctx, ccall((:femb_pattern_build, libfemb), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{Int64}, Ref{Int64}), ctx.h, 3, brow, bcol, 0, C_NULL, nnzref))
        n = size(csc, 1)
        colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nnzref[])
        check(ctx, ccall((:femb_pattern_get, libfemb), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), ctx.h, colptr, rowval))
        Ae.cscmatrix = SparseMatrixCSC{Float64, Int64}(n, n, colptr, rowval, zeros(Float64, nnzref[]))
    else
        check(ctx, ccall((:femb_pattern_adopt, libfemb), Int32,
            (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int32}, Ptr{Int32}),
            ctx.h, size(csc, 1), size(csc, 2), nnz(csc), csc.colptr, csc.rowval, 3, brow, bcol))
    end
    check(ctx, ccall((:femb_matrix_zero, libfemb), Int32, (Ptr{Cvoid},), ctx.h))
    check(ctx, ccall((:femb_assemble_hdiv_mixed, libfemb), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Float64, Float64), ctx.h, 0, 1, 2, 1.0, 1.0))
    check(ctx, ccall((:femb_matrix_add_to, libfemb), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, Ae.cscmatrix.nzval))
    return nothing
end

end # module
