/*
 * femb.h — C ABI of libfemb: batched cell-wise finite-element assembly on one B200 (sm_100a).
 *
 * This is the drop-in boundary for the hot path of WIAS-PDELib/ExtendableFEMBase.jl named in
 * BASELINE.json: the per-cell loops of
 *     examples/Example200_LowLevelPoisson.jl:103-184   (assemble!)
 *     examples/Example210_LowLevelNavierStokes.jl:218-394 (prepare_assembly!/update_system!)
 * and the per-cell evaluators they call (src/feevaluator.jl:344-387, src/feevaluator_h1.jl,
 * src/feevaluator_hdiv.jl) plus the sparse insert they end in (ExtendableSparse rawupdateindex!/
 * flush!, call sites Example200:155-157,182; Example210:373-379,391).
 *
 * The reference has no FFI of its own (it is pure Julia); each entry point below states which
 * reference object or call it takes its data from / replaces.  The Julia-side `ccall` stubs are in
 * julia/FEMBaseB200.jl and INTEGRATION.md; the Python ctypes mirror is extendablefembase.jl_b200/_lib.py.
 *
 * Conventions
 *   - every function returns int32 status (0 = FEMB_OK); femb_last_error() gives the message.
 *     Nothing throws, aborts or longjmps across this boundary.
 *   - Julia arrays are passed AS IS: column-major, 1-based Int32 grid/dof indices, 1-based Int64
 *     CSC indices; kernels subtract 1.  A Julia `k x n` matrix is `n` contiguous groups of `k`.
 *   - all pointers are HOST pointers unless the name ends in `_dev`.  The library copies inputs
 *     into device buffers it owns; outputs are written into caller-provided host buffers.
 *   - a femb_ctx is bound to one CUDA device and is not thread-safe; use one per thread / GPU.
 *   - there is no CPU fallback: every entry point that computes fails with FEMB_ERR_CUDA if no
 *     sm_100-class device is usable.
 */
#ifndef FEMB_H
#define FEMB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct femb_ctx femb_ctx;

enum {
    FEMB_OK = 0,
    FEMB_ERR_ARG = 1,      /* invalid argument / call order */
    FEMB_ERR_CUDA = 2,     /* CUDA runtime error (message has cudaGetErrorString) */
    FEMB_ERR_UNSUPPORTED = 3,
    FEMB_ERR_PATTERN = 4,  /* a contribution has no slot in the adopted pattern */
    FEMB_ERR_NCCL = 5
};

/* FE class of a space: selects the push-forward (src/feevaluator_h1.jl, feevaluator_hdiv.jl, feevaluator_hcurl.jl) */
enum { FEMB_FAMILY_H1 = 0, FEMB_FAMILY_HDIV = 1, FEMB_FAMILY_L2 = 2, FEMB_FAMILY_HCURL = 3 };

/* per-cell coefficient / subset handlers (get_coefficients, get_basissubset of src/fedefs) */
enum {
    FEMB_HANDLER_NONE = 0,
    FEMB_HANDLER_H1_FACESWAP2D = 1, /* h1_p3.jl:200-217, h1_pk.jl:280-304 */
    FEMB_HANDLER_H1_EDGESWAP3D = 2, /* h1_p3.jl:220-241 */
    FEMB_HANDLER_RT0_SIGNS = 3,     /* hdiv_rt0.jl:114-124 */
    FEMB_HANDLER_BDM1_2D = 4,       /* hdiv_bdm1.jl:200-212 */
    FEMB_HANDLER_BDM1_3D = 5,       /* hdiv_bdm1.jl:215-249 */
    FEMB_HANDLER_N0_EDGESIGNS3D = 6 /* hcurl_n0.jl:156-166 (CellEdgeSigns; HCURLN0{2} takes FEMB_HANDLER_RT0_SIGNS: face signs, :144-154) */
};

/* function operators implemented on the device (src/functionoperators.jl:132-196) */
enum { FEMB_OP_IDENTITY = 0, FEMB_OP_GRADIENT = 1, FEMB_OP_DIVERGENCE = 2 };

/* right-hand-side data kinds (the Julia closures rhs(x) / f!(fval,x,t) cannot be called from
 * the device; Example200:38,173 and Example210:44-48,311 are covered by these) */
enum {
    FEMB_RHS_NONE = 0,
    FEMB_RHS_AFFINE = 1, /* f_c(x) = params[c*(dim+1)] + sum_d params[c*(dim+1)+1+d] * x_d          */
    FEMB_RHS_QPVALUES = 2, /* fvals[ncomp x nqp x ncells] evaluated by the host at x_qp (femb_qp_coords) */
    FEMB_RHS_EX210 = 3   /* Example210 f!: params = {mu, t}                                         */
};

/* flags of femb_assemble_bilinear */
enum {
    FEMB_BIL_TRANSPOSE_COPY = 1, /* also add the transposed local matrix into block (bcol,brow): Example210:376-380 */
    FEMB_BIL_SYMMETRIC_UPPER = 2, /* Example200:139-158: compute j<=k, mirror, skip |v| <= drop_tol   */
    FEMB_BIL_NO_VOLUME = 4       /* do not multiply by |T|                                           */
};

/* scatter policy for the cell-parallel kernels */
enum {
    FEMB_SCATTER_ATOMIC = 0,   /* red.global.add.f64 (fastest; summation order not fixed)            */
    FEMB_SCATTER_COLORED = 1,  /* cell colouring, plain read-modify-write per colour (bit-reproducible) */
    FEMB_SCATTER_GATHER = 2    /* default: owner-computes column gather, write-once, bit-reproducible, where a gather
                                * kernel exists (scalar H1 P1 and P2 stiffness + rhs of femb_assemble_poisson, every
                                * linear form); the other bilinear forms use the atomic scatter under this policy  */
};

/* ---- lifetime ----------------------------------------------------------------------------- */
int32_t femb_version(void);
int32_t femb_create(int32_t device, femb_ctx** out);
int32_t femb_destroy(femb_ctx* ctx);
const char* femb_last_error(const femb_ctx* ctx); /* valid until the next call on ctx; ctx may be NULL */
int32_t femb_sync(femb_ctx* ctx);
/* pin / unpin a caller-owned host range so that copies run at full PCIe rate and overlap */
int32_t femb_host_register(femb_ctx* ctx, void* ptr, int64_t bytes);
int32_t femb_host_unregister(femb_ctx* ctx, void* ptr);

/* ---- stage-1 inputs: xgrid[Coordinates], xgrid[CellNodes], xgrid[CellVolumes] ----------------
 * (ExtendableGrids components read at Example200:107,114; Example210:227-228; feevaluator.jl:94-97)
 * coords: dim x nnodes Float64; cellnodes: nnpc x ncells Int32 1-based; cellvolumes: ncells or NULL
 * (NULL: |T| = |det A|/dim! is computed on the device).
 * A call with the same dim / nnodes / ncells compares the new CellNodes with the registered one on the device: if it is
 * identical only the geometry is replaced and everything derived from the connectivity (spaces, pattern, gather maps,
 * grid entities) stays valid — the steady state of a moving-mesh loop; if it differs anywhere (edge flips,
 * renumbering) the call behaves like a new mesh: spaces, pattern and entities are dropped and must be registered again. */
int32_t femb_mesh_set(femb_ctx* ctx, int32_t dim, int64_t nnodes, const double* coords, int32_t nnpc,
                      int64_t ncells, const int32_t* cellnodes, const double* cellvolumes_or_null);

/* ---- FESpace: fe_space[CellDofs] (+ xgrid[CellFaceSigns], xgrid[CellFaceOrientations]) --------
 * (finiteelements.jl:376-388, dofmaps.jl:419-627; handler inputs hdiv_rt0.jl:115, hdiv_bdm1.jl:201,235)
 * celldofs: nd_local x ncells Int32 1-based, constant stride (single cell geometry);
 * celldofs == NULL means "identical to cellnodes" (H1P1{1}).  signs/orient: nfaces x ncells Int32 or
 * NULL.  For FEMB_HANDLER_H1_EDGESWAP3D `signs` is xgrid[CellEdgeSigns] (6 x ncells). */
int32_t femb_space_set(femb_ctx* ctx, int32_t space_id, int32_t family, int32_t ncomp, int64_t ndofs,
                       int32_t nd_local, int32_t nd_all, int32_t handler, const int32_t* celldofs_or_null,
                       const int32_t* signs_or_null, const int32_t* orient_or_null);

/* ---- grid components on the device: face / edge enumeration -----------------------------------
 * Replaces the lazy instantiation of xgrid[CellFaces], xgrid[FaceNodes], xgrid[CellFaceSigns],
 * xgrid[CellFaceOrientations], xgrid[FaceCells] (kind FACES) and xgrid[CellEdges], xgrid[EdgeNodes],
 * xgrid[CellEdgeSigns] (kind EDGES; on triangles the edges are the faces) of ExtendableGrids for the mesh of
 * femb_mesh_set (read at dofmaps.jl:448-458, hdiv_rt0.jl:115, hdiv_bdm1.jl:201,235; Example200:55 times it).
 * Numbering = first appearance while looping cells and local entities in order; entity_nodes = node order of the
 * creating cell; face sign +1 creator / -1 neighbour; edge sign +1 when the local edge starts at the global edge's
 * first node; orientations 1..4 (3-D) as in hdiv_bdm1.jl:238-246.  All arrays Int32, 1-based, column-major
 * (k x nitems).  femb_grid_entities_get: any pointer may be NULL; sizes: cell_entities / signs / orientations
 * per_cell x ncells, entity_nodes nodes_per x n, facecells 2 x n (faces only). */
enum { FEMB_ENTITY_FACES = 0, FEMB_ENTITY_EDGES = 1 };
int32_t femb_grid_entities(femb_ctx* ctx, int32_t kind, int64_t* n_out);
int32_t femb_grid_entities_get(femb_ctx* ctx, int32_t kind, int32_t* cell_entities, int32_t* entity_nodes, int32_t* signs,
                               int32_t* orientations, int32_t* facecells);

/* ---- FESpace with the dofmap generated on the device: init_dofmap_from_pattern! (dofmaps.jl:419-627) ----
 * segs: nseg triples (type, each_component, q) of the parsed dofmap pattern (dofmaps.jl:263-301), type
 * 0 = N/n nodes, 1 = F/f faces, 2 = E/e edges, 3 = I/i/C/c cell interior.  Faces / edges and the sign /
 * orientation arrays the handler needs are taken from the device enumeration above.  Outputs ndofs and the
 * component offset (count_ndofs, finiteelements.jl:395-449).  femb_space_get_celldofs downloads fe_space[CellDofs]
 * (nd_local x ncells Int32, 1-based) of any registered space. */
int32_t femb_space_set_from_pattern(femb_ctx* ctx, int32_t space_id, int32_t family, int32_t ncomp, int32_t nd_local,
                                    int32_t nd_all, int32_t handler, int32_t nseg, const int32_t* segs, int64_t* ndofs_out,
                                    int64_t* coffset_out);
int32_t femb_space_get_celldofs(femb_ctx* ctx, int32_t space_id, int32_t* celldofs);

/* ---- after the assembly: Dirichlet penalisation and the residual, without a host round trip of A ---------
 * femb_boundary_dofs: boundarydofs(FES) (dofmaps.jl:871-899; Example200:86, Example210:172) for a registered space
 * from the device face enumeration: every dof attached to a face with one neighbour cell; segs = the dofmap pattern
 * as in femb_space_set_from_pattern.  The list is sorted and unique (the reference returns BFaceDofs.colentries,
 * which repeats shared dofs; its callers only loop over it), 1-based, local to the space.
 * femb_dirichlet_apply: A[d,d] = penalty; b[d] = penalty * values[d] (values == NULL: the device-resident boundary data of
 * femb_interpolate, zeros if there is none) for d in
 * (boundary list + dof_offset, when use_boundary_list) and the extra dofs (1-based, matrix numbering) —
 * Example200:87-90, Example210:197-200 (penalty 1e60, the extra dof fixes one pressure value, Example210:173).
 * femb_residual: res = A x - b (x == NULL: the vector of femb_vector_set(1, ..)), rows fixed by the last
 * femb_dirichlet_apply zeroed when zero_fixed; norm_out = ||res||_2 (Example210:180-185, :204-209).  Row-wise
 * sums in ascending column order, write-once: bit-reproducible. */
int32_t femb_boundary_dofs(femb_ctx* ctx, int32_t space_id, int32_t nseg, const int32_t* segs, int64_t* n_out);
int32_t femb_boundary_dofs_get(femb_ctx* ctx, int32_t* dofs);
int32_t femb_dirichlet_apply(femb_ctx* ctx, int32_t use_boundary_list, int64_t dof_offset, int64_t n_extra,
                             const int64_t* extra_dofs, const double* values_or_null, double penalty);
/* interpolate!(u_init[1], u!; time) (Example210:168-173) on the device, for nodal H1 spaces whose dofmap pattern is one
 * dof per node and, optionally, one dof per edge (2-D: face) and component — H1P1, H1P2, H1Pk of order <= 2, any number
 * of components.  Node dofs: point evaluation (NodalInterpolator, src/interpolators.jl:93-127).  Edge dofs: the edge mean
 * is preserved (MomentInterpolator of order 0, src/fedefs/h1_p2.jl:47-76, src/interpolators.jl:137-460):
 * value = (mean_e(u) - (u(a) + u(b)) / 6) * 3/2 with mean_e by the rule (qx in [0,1], qw, nq points) the caller passes —
 * the reference uses its Gauss rule of order 2 * order_FE (src/quadrature.jl:214-233, :609-628); nq = 0 evaluates at the
 * edge midpoint instead.  func_kind / params: FEMB_FUNC_AFFINE (per component c0, g_1..g_dim), FEMB_FUNC_EX210_U {mu, t}
 * (the exact velocity u! of Example210:51-57).  only_boundary: only the dofs of the last femb_boundary_dofs are written.
 * The result is kept on the device as the boundary-data vector (matrix numbering: dof_offset + dof, zero elsewhere) that
 * femb_dirichlet_apply uses when its values argument is NULL — no host round trip of u_init; target_or_null (ndofs of the
 * space) receives a copy.  Edges are oriented by their global node ids: the result is bit-reproducible. */
enum { FEMB_FUNC_AFFINE = 1, FEMB_FUNC_EX210_U = 2 };
int32_t femb_interpolate(femb_ctx* ctx, int32_t space_id, int32_t nseg, const int32_t* segs, int64_t dof_offset, int32_t func_kind,
                         const double* params, int32_t nq, const double* qx, const double* qw, int32_t only_boundary, double* target_or_null);
int32_t femb_residual(femb_ctx* ctx, const double* x_or_null, int32_t zero_fixed, double* res_or_null, double* norm_out);
/* compute_error (Example210:110-154) through a registered evaluator (register it after femb_quadrature_set with the
 * VertexRule, quadrature.jl:103-186): error_out[d, cell] = sum_qp (uh_d(x_qp) - exact_qp[d, qp, cell])^p |T| w_qp with
 * uh = eval_febe! of `coefficients` (the ndofs values of the evaluator's space); exact_qp may be NULL (u = nothing).
 * error_out: resultdim x ncells. */
int32_t femb_eval_error(femb_ctx* ctx, int32_t eval_id, const double* coefficients, const double* exact_qp_or_null, int32_t p,
                        double* error_out);
/* addblock_matmul!(a, B, b; factor, transposed) (src/fematrix.jl:486-563): a += factor * B b (or B' b) for the block
 * (brow, bcol) of the matrix layout, vectors in block-local numbering; brow = bcol = -1: the whole matrix (:570-602).
 * femb_lrmatmul: lrmatmul / ldrdmatmul (src/fematrix.jl:607-637): (a1 - a2)' A (b1 - b2) * factor over the whole matrix
 * (a2 / b2 may be NULL).  Row-wise / column-wise sums with a fixed reduction tree: bit-reproducible. */
int32_t femb_block_matmul(femb_ctx* ctx, int32_t brow, int32_t bcol, const double* b, double factor, int32_t transposed,
                          double* a_inout);
int32_t femb_lrmatmul(femb_ctx* ctx, const double* a1, const double* a2_or_null, const double* b1, const double* b2_or_null,
                      double factor, double* out);
/* Alin = deepcopy(A); blin = deepcopy(b) (Example210:164-165) and A += Alin; b += blin (Example210:192-193), on the device */
int32_t femb_system_save(femb_ctx* ctx);
int32_t femb_system_add_saved(femb_ctx* ctx);

/* ---- QuadratureRule: qf.xref (dim x nqp), qf.w (nqp)   (quadrature.jl:38-42) ------------------ */
int32_t femb_quadrature_set(femb_ctx* ctx, int32_t nqp, const double* xref, const double* w);

/* ---- FEEvaluator(FES, op, qf): the reference tables built once by the constructor -------------
 * (feevaluator.jl:83-215, :233-291).  refvals: [nqp][nd_all][ncomp]; refderivs:
 * [nqp][nd_all][ncomp][dim] = d(phi_{dof,comp})/d(xref_j) (either may be NULL if the operator does
 * not need it).  Uses the quadrature rule current at the time of the call. */
int32_t femb_evaluator_set(femb_ctx* ctx, int32_t eval_id, int32_t space_id, int32_t op,
                           const double* refvals_or_null, const double* refderivs_or_null);

/* batched update_basis!(FEB, cell) for cells [cell_begin, cell_end) (0-based range):
 * out[(cell-cell_begin)][qp][dof][r] = FEB.cvals[r, dof, qp] after update_basis!(FEB, cell+1)
 * (feevaluator.jl:381-387 -> feevaluator_h1.jl:2-14,61-74,119-130; feevaluator_hdiv.jl:2-19,66-83) */
int32_t femb_evaluator_eval(femb_ctx* ctx, int32_t eval_id, int64_t cell_begin, int64_t cell_end, double* out);

/* physical quadrature points eval_trafo!(x, L2G, xref[qp]) for all cells: out[cell][qp][dim]
 * (Example200:171, Example210:310) — lets the host evaluate rhs closures for FEMB_RHS_QPVALUES */
int32_t femb_qp_coords(femb_ctx* ctx, double* out);

/* ---- FEMatrix(FESX, FESY): block layout + CSC pattern ------------------------------------------
 * (fematrix.jl:226-290: row offsets follow space_rows, column offsets space_cols) */
int32_t femb_matrix_layout(femb_ctx* ctx, int32_t nrow_spaces, const int32_t* row_spaces,
                           int32_t ncol_spaces, const int32_t* col_spaces);
/* structural pattern  U_cells dofs_row(cell) x dofs_col(cell)  for the listed blocks — what the first
 * assembly + flush! of the reference produces when nothing is filtered (Example210; and
 * apply_nonzero_pattern!, fematrix.jl:108-118).  extra_diag: 1-based global dofs whose diagonal entry
 * must exist (Dirichlet penalties, Example200:86-93, Example210:172,197-199). */
int32_t femb_pattern_build(femb_ctx* ctx, int32_t nblocks, const int32_t* block_rows, const int32_t* block_cols,
                           int64_t n_extra_diag, const int64_t* extra_diag_or_null, int64_t* nnz_out);
/* use an already flushed Julia pattern A.cscmatrix.colptr / rowval (steady state, Example200:205-211) */
int32_t femb_pattern_adopt(femb_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* colptr,
                           const int64_t* rowval, int32_t nblocks, const int32_t* block_rows, const int32_t* block_cols);
int32_t femb_pattern_get(femb_ctx* ctx, int64_t* colptr_out, int64_t* rowval_out); /* either may be NULL */
/* rowval at the given 0-based nz positions (gathered on the device): the interface columns of a partitioned mesh
 * without downloading the whole pattern */
int32_t femb_pattern_get_slots(femb_ctx* ctx, int64_t n, const int64_t* slots, int64_t* rowval_out);
/* Example200's value filter (|Aloc[j,k]| > 1e-15, Example200:153) decides the pattern: drop every
 * slot that received no contribution in the assemblies since the last femb_matrix_zero, keeping
 * extra_diag entries.  Rebuilds the cell->slot maps.  */
int32_t femb_pattern_compress(femb_ctx* ctx, int64_t* nnz_out);
/* enable/disable the per-slot "received a contribution" flags that femb_pattern_compress reads */
int32_t femb_pattern_track(femb_ctx* ctx, int32_t enable);

/* ---- device-resident FEMatrix.entries.cscmatrix.nzval and FEVector.entries ---------------------- */
int32_t femb_matrix_zero(femb_ctx* ctx);
int32_t femb_vector_zero(femb_ctx* ctx);
int32_t femb_matrix_get(femb_ctx* ctx, double* nzval_out);           /* nnz doubles, Julia CSC order */
int32_t femb_matrix_add_to(femb_ctx* ctx, double* nzval_inout);      /* nzval_inout += device values  */
/* partial downloads: nzval[first .. first+count) / b[first .. first+count) (0-based) into out[0 .. count) —
 * e.g. the interface columns after femb_exchange */
int32_t femb_matrix_get_range(femb_ctx* ctx, int64_t first, int64_t count, double* out);
int32_t femb_vector_get_range(femb_ctx* ctx, int64_t first, int64_t count, double* out);
int32_t femb_vector_get(femb_ctx* ctx, double* b_out);               /* sum of row-space ndofs        */
int32_t femb_vector_set(femb_ctx* ctx, int32_t which, const double* v); /* which: 0 = b, 1 = current solution (Example210:328-331) */
int32_t femb_set_scatter(femb_ctx* ctx, int32_t policy);

/* ---- stages 3-5 ----------------------------------------------------------------------------------
 * generic quadrature-weighted contraction of two evaluators + scatter into block (brow,bcol):
 *   Aloc[j,k] = factor * |T| * sum_qp w_qp <op_row_j(qp), op_col_k(qp)>      (Example210:284-301)
 * brow/bcol index into the row/col space lists of femb_matrix_layout (0-based). */
int32_t femb_assemble_bilinear(femb_ctx* ctx, int32_t eval_row, int32_t eval_col, int32_t brow, int32_t bcol,
                               double factor, int32_t flags, double drop_tol);
/* The blocks of a mixed (Darcy-type) system over the layout [v; q], v an Hdiv space, q = L2P0, in one call (config C4 of BASELINE.json):
 *   A[0,0] += factor_mass (v_j, v_k),  A[0,1] += factor_div (div v_j, q),  A[1,0] += its transpose
 * — the cell loop of Example200:129-161 with the evaluators of src/feevaluator_hdiv.jl:2-19 (Identity) and :66-83 (Divergence); equal to
 * femb_assemble_bilinear(e_id, e_id, 0, 0, factor_mass, 0, 0) + femb_assemble_bilinear(e_div, e_q, 0, 1, factor_div,
 * FEMB_BIL_TRANSPOSE_COPY, 0).  For HDIVRT0 on tetrahedra right after femb_matrix_zero every column of the matrix is written once and in
 * one piece by two kernels (closed-form cell pre-pass + owner-computes column gather), without zeroing. */
int32_t femb_assemble_hdiv_mixed(femb_ctx* ctx, int32_t eval_id, int32_t eval_div, int32_t eval_q, double factor_mass, double factor_div);
/* b[off(brow)+dof_j] += factor * |T| * sum_qp w_qp <op_j(qp), f(x_qp)>   (Example200:165-178, Example210:304-318) */
int32_t femb_assemble_linear(femb_ctx* ctx, int32_t eval_id, int32_t brow, double factor, int32_t rhs_kind,
                             const double* params_or_fvals);
/* Example200.assemble!(A, b, fe_space, rhs, mu) in one call (Example200:103-184): stiffness with the
 * upper-triangle/mirror/drop_tol rule + rhs.  For H1P1 with FEMB_SCATTER_GATHER this is ONE fused
 * kernel (stages 1-5). */
int32_t femb_assemble_poisson(femb_ctx* ctx, int32_t eval_grad, int32_t eval_id, double mu, int32_t rhs_kind,
                              const double* params_or_fvals, double drop_tol);
/* The same assembly with HOST buffers end to end, software-pipelined over `nchunks` column ranges on three streams:
 * upload of Coordinates / CellVolumes (prefix needed by the next chunk) and CellNodes, the fused kernel per chunk,
 * and download of the finished nzval / b ranges overlap.  Equivalent to femb_mesh_set (same shapes: the symbolic
 * phase is reused) + femb_matrix_zero + femb_vector_zero + femb_assemble_poisson + femb_matrix_get +
 * femb_vector_get, bit for bit.  Scalar P1 with rhs_kind NONE / AFFINE, and scalar P2 on tetrahedra (closed-form gather:
 * the geometry is uploaded once, the finished nzval / b ranges of every launch chunk follow the kernels); pin the host
 * arrays with femb_host_register for the copies to overlap.  cellvol must be given iff the registered mesh has volumes.
 * cellnodes = NULL declares the connectivity unchanged (the numeric phase reads the geometry only: the fast, default way
 * to call this).  cellnodes != NULL is uploaded behind the geometry and compared with the registered connectivity on
 * the device; if it differs the call returns FEMB_ERR_ARG (the buffers then hold the matrix of the OLD topology and must
 * be discarded): call femb_mesh_set and rebuild the pattern. */
int32_t femb_assemble_poisson_host(femb_ctx* ctx, int32_t eval_grad, int32_t eval_id, double mu, int32_t rhs_kind, const double* params,
                                   double drop_tol, const double* coords, const int32_t* cellnodes, const double* cellvolumes_or_null,
                                   double* nzval_out, double* b_out, int32_t nchunks);
/* Example210 update_system!(linear, nonlinear) (Example210:276-385): Laplacian + div-pressure blocks +
 * rhs (linear) and the Newton-linearised convection term + its rhs (nonlinear; uses vector 1 =
 * current solution).  evals: grad u, id u, id p. */
int32_t femb_assemble_navierstokes(femb_ctx* ctx, int32_t eval_grad_u, int32_t eval_id_u, int32_t eval_id_p,
                                   double mu, int32_t linear, int32_t nonlinear, int32_t rhs_kind,
                                   const double* params_or_fvals);

/* ---- timing / introspection (bench + tests) ------------------------------------------------------ */
/* CUDA-event time of the kernels launched by the last femb_assemble_* call, in ms, and their count */
int32_t femb_last_kernel_ms(femb_ctx* ctx, float* ms_out, int32_t* nlaunches_out);
int64_t femb_launch_count(const femb_ctx* ctx); /* kernels launched by this ctx so far */
/* CUDA events on the library's own stream bracketing an arbitrary sequence of calls (bench.py's timed
 * region; the reference's counterpart is the @elapsed around assemble!, Example200:82-94).  _end
 * synchronises and returns the device time between the two events. */
int32_t femb_region_begin(femb_ctx* ctx);
int32_t femb_region_end(femb_ctx* ctx, float* ms_out);
/* raw device pointers (for harnesses that keep data resident): 0 nzval, 1 b, 2 colptr, 3 rowval */
void* femb_device_ptr(femb_ctx* ctx, int32_t which);

/* ---- multi-GPU: cells partitioned across ranks, interface columns exchanged over NCCL ------------
 * (no counterpart in the reference, which is single-process: SURVEY.md §8(e)).  One process per GPU.
 * Each rank registers its LOCAL mesh part with LOCAL dof numbering: the dofs of its own cells plus "halo" dofs
 * (rows that peers contribute to columns this rank owns).  A dof on a cut is owned by exactly one rank; the other
 * ranks hold it as a ghost: their contributions to its matrix column / vector entry are sent to the owner. */
/* extra structural entries (1-based row / column in the layout's numbering) merged into the next
 * femb_pattern_build: the rows a peer will contribute to columns owned here. */
int32_t femb_pattern_extra(femb_ctx* ctx, int64_t n, const int64_t* rows, const int64_t* cols);
int32_t femb_comm_unique_id(char id_out[128]);  /* ncclGetUniqueId; broadcast it to all ranks out of band */
int32_t femb_comm_init(femb_ctx* ctx, int32_t nranks, int32_t rank, const char id[128]);
/* exchange plan: for peer p (rank peers[p]) this rank SENDS nzval[send_slots[send_ptr[p] .. send_ptr[p+1])] and
 * b[send_b[sendb_ptr[p] ..)] and ADDS what it receives from p into nzval[recv_slots[recv_ptr[p] ..)] /
 * b[recv_b[recvb_ptr[p] ..)] (0-based positions; the two sides list the same entries in the same order). */
int32_t femb_comm_set_exchange(femb_ctx* ctx, int32_t npeers, const int32_t* peers, const int64_t* send_ptr, const int64_t* send_slots,
                               const int64_t* recv_ptr, const int64_t* recv_slots, const int64_t* sendb_ptr, const int64_t* send_b,
                               const int64_t* recvb_ptr, const int64_t* recv_b);
/* pack -> ncclSend/ncclRecv (one group) -> add in ascending peer order (reproducible).  Received non-zero matrix
 * values also mark their slot as touched (femb_pattern_compress).  Sent ghost entries are left in place. */
int32_t femb_exchange(femb_ctx* ctx);
/* enable = 1: femb_assemble_poisson performs the exchange itself.  On the scalar P1 gather path it is overlapped with
 * the bulk of the assembly: the slices holding ghost and owned-interface columns are launched first, the pack / NCCL /
 * add sequence runs on a second stream while the remaining columns are assembled; every other dispatch path runs the
 * same exchange right after its kernels.  Do not call femb_exchange then.  Each such assembly must start from
 * femb_matrix_zero + femb_vector_zero (the whole ghost entry is sent: FEMB_ERR_ARG otherwise); without an rhs the
 * vector part of the plan carries zeros.  The plan must have been installed for the current pattern
 * (femb_pattern_build / _adopt / _compress drop it). */
int32_t femb_set_fused_exchange(femb_ctx* ctx, int32_t enable);
int32_t femb_comm_destroy(femb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FEMB_H */
