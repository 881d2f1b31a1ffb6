// Internal state of libfemb (not part of the C ABI; see include/femb.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "femb.h"

#define FEMB_MAX_SPACES 8
#define FEMB_MAX_EVALS 16
#define FEMB_MAX_QP 32
#define FEMB_MAX_BLOCKS 16

// extra operator codes used internally by femb_assemble_navierstokes
#define FEMB_OP_GRADTRACE 3   // g[1] + g[4] of the Gradient (Example210:296)
#define FEMB_OP_CONVECTION 4  // J(u_h) * [phi; grad phi]  (Example210:339-350)

struct FembSpace {
    bool set = false;
    int family = 0, ncomp = 0, nd = 0, nd_all = 0, handler = 0;
    int64_t ndofs = 0;
    int32_t* celldofs = nullptr;  // device, nd x ncells, 1-based (may alias mesh.cellnodes)
    bool celldofs_owned = false;
    int32_t* signs = nullptr;     // device
    int32_t* orient = nullptr;    // device
    int nsign = 0;                // rows of signs (faces or edges per cell)
    // dof -> (cell, local index) adjacency, sorted by (dof, cell): adj[e] = cell*nd + k
    int32_t* adj_ptr = nullptr;   // ndofs+1
    uint32_t* adj = nullptr;      // ncells*nd
    int max_deg = 0;
};

// grid entities enumerated on the device (femb_grid.cu): faces [0] and, in 3-D, edges [1]
struct FembEntities {
    bool built = false;
    int64_t n = 0;                 // number of faces / edges
    int per_cell = 0, nodes_per = 0;
    int32_t* cell_ent = nullptr;   // per_cell x ncells, 1-based ids in order of first appearance (xgrid[CellFaces] / [CellEdges])
    int32_t* ent_nodes = nullptr;  // nodes_per x n, 1-based, local order of the creating cell (xgrid[FaceNodes] / [EdgeNodes])
    int32_t* signs = nullptr;      // per_cell x ncells (xgrid[CellFaceSigns] / [CellEdgeSigns])
    int32_t* orient = nullptr;     // faces: xgrid[CellFaceOrientations]
    int32_t* facecells = nullptr;  // faces: 2 x n, (creating cell, neighbour or 0) (xgrid[FaceCells])
};

#define FEMB_MAX_PATTERN_SEGS 8

struct FembEval {
    bool set = false;
    int space = -1, op = 0, nqp = 0, resultdim = 0;
    double* refvals = nullptr;    // device [nqp][nd_all][ncomp]
    double* refderivs = nullptr;  // device [nqp][nd_all][ncomp][dim]
    std::vector<double> h_refvals; // host copy (P1 gather kernel takes the Identity table by value)
    std::vector<double> h_refderivs; // host copy (reference tensor of the DMMA stiffness kernel)
    double* reftensor = nullptr;  // device [KS*4][NT*8], built lazily by the DMMA stiffness path
    int reftensor_nt = 0;
    double xref[FEMB_MAX_QP * 3];
    double w[FEMB_MAX_QP];
};

// H1 gather map: sliced ELL over (column dof, adjacent cell) pairs, SoA planes [row][lane]: plane 0 = cell, planes 1.. = positions
// of the cell's local rows in the dof's CSC column packed 4 x 8 bit per word (0xFF = not in the pattern); the spare bytes of the
// last word carry the local index k of the dof (within its component) and the first-touch flags
struct FembHMap {
    uint32_t* map = nullptr;
    size_t stride = 0;
    int32_t* slice_ptr = nullptr;
    int32_t *sub0 = nullptr, *sublen = nullptr;  // vector maps: the run [sub0, sub0 + sublen) of each column that belongs to the block (positions are relative to it)
    uint8_t* slice_zero = nullptr;  // per slice: some column has positions without a local contribution (zero the accumulators)
    int space = -1, words = 0;
    int64_t col_off = 0;            // first column of the block in the matrix
    bool complete = false;          // every (cell, row) of the map has a slot: the gather may skip its per-entry checks
    int twins = 1;                  // vector maps: number of components whose blocks are copies of each other (1 = not established)
    bool canon = false;             // positions / first-touch flags of every entry are in P2 role order (femb_p2.cu) instead of local dof order
    struct Range { int64_t s0, s1; int maxlen; int split; int mode = 2; int maxrows = 0; };  // split: threads per column (1, or 4 for high-degree columns)
    std::vector<Range> ranges;      // contiguous slice ranges with their longest CSC column (shared-memory sizing)
    struct XList { int32_t* ifc = nullptr; int64_t n_ifc = 0; int32_t* rest = nullptr; int64_t n_rest = 0; };  // per crange: interface slices / all others
    std::vector<XList> xlists;
    int xlists_version = -1;
    int32_t* vrow = nullptr;        // femb_p2.cu (3-D): per slice the first row of its vertex-dof sectors in vrec, -1 = none
    uint32_t* vinv = nullptr;       // [cell][4]: slot in vrec of the cell's vertex-dof visits
    double* vrec = nullptr;         // vertex-dof sectors in visit order (4 doubles per slot)
    uint32_t* xsrc = nullptr;       // femb_hdiv.cu (mixed RT0 pass): per visit the nzval slot of the (face dof, cell) entry of block (0,1), bit 31 = position of
                                    // its transposed copy in the (1,0) run of the face column; xsrc_state 0 = not built, 1 = usable, -1 = layout not covered
    int xsrc_state = 0;
    int th_state = 0;               // femb_ns.cu (owner-computes div-pressure blocks): 0 = column layout not checked, 1 = as expected, -1 = not covered
    std::vector<Range> cranges;     // femb_p2.cu: `ranges` cut where the kind of dof changes (mode 0 = vertex dofs, 1 = edge dofs, 2 = both)
};

struct FembBlock {
    int brow = 0, bcol = 0;       // indices into layout row/col space lists
    int32_t* slots = nullptr;     // device [ncells][ndr][ndc] -> 0-based nz index, -1 = absent
};

struct femb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    // streamed host-to-host assembly (femb_assemble_poisson_host)
    cudaStream_t stream_h2d = nullptr, stream_d2h = nullptr;
    cudaStream_t stream_aux = nullptr;          // second compute stream (femb_p2.cu: vertex-column and edge-column kernels side by side)
    cudaEvent_t ev_a0 = nullptr, ev_a1 = nullptr;
    std::vector<cudaEvent_t> chunk_ev;          // 2 per chunk: upload done, kernel done
    std::vector<int64_t> chunk_slice, chunk_nmax, chunk_cmax, chunk_cp;  // per chunk: first slice, max node / cell id referenced, colptr at the chunk start (0-based)
    int chunk_n = 0;
    std::string err;
    int64_t launches = 0;
    int last_launches = 0;
    float last_ms = 0.f;
    int scatter = FEMB_SCATTER_GATHER;
    int sm_count = 148;

    // mesh
    int dim = 0, nnpc = 0;
    int64_t nnodes = 0, ncells = 0;
    double* coords = nullptr;     // dim x nnodes
    int32_t* cellnodes = nullptr; // nnpc x ncells
    int32_t* cellnodes_tmp = nullptr;  // scratch copy for the "is the connectivity unchanged" compare
    double* cellvol = nullptr;    // ncells or null

    FembSpace spaces[FEMB_MAX_SPACES];
    FembEval evals[FEMB_MAX_EVALS];

    // quadrature
    int nqp = 0;
    double xref[FEMB_MAX_QP * 3];
    double w[FEMB_MAX_QP];

    // layout
    int nrs = 0, ncs = 0;
    int row_spaces[FEMB_MAX_SPACES], col_spaces[FEMB_MAX_SPACES];
    int64_t row_off[FEMB_MAX_SPACES + 1], col_off[FEMB_MAX_SPACES + 1];
    int64_t nrows = 0, ncols = 0;

    // pattern (device, 1-based int64 like Julia)
    int64_t nnz = 0;
    int64_t* colptr = nullptr;
    int64_t* rowval = nullptr;
    int max_collen = 0;
    std::vector<FembBlock> blocks;
    int64_t block_nnz[FEMB_MAX_SPACES][FEMB_MAX_SPACES];  // pattern entries per layout block, -1 = not counted yet (femb_block_entries)
    std::vector<int64_t> extra_diag;
    // P1 gather map, sliced ELL (slice = 32 consecutive dofs = one warp), structure of arrays: plane p lives at
    // gmap + p * gmap_stride, element [row][lane].  Planes: 0 = na, 1 = nb (the e-th adjacent cell's other nodes in
    // local order, 0-based), 2 = w (local index k of the dof + positions of the local rows in the dof's CSC
    // column, see femb_pattern.cu), 3 = cell, 4 = nc (tetrahedra only)
    uint32_t* gmap = nullptr;
    size_t gmap_stride = 0;
    int32_t* gslice_ptr = nullptr;   // nslices + 1 row offsets
    int64_t gmap_rows = 0;
    bool gmap_uniform = false;       // every slice padded to gmap_maxrows rows (row0 = slice * maxrows, no indirection)
    int gmap_maxrows = 0;            // longest slice (rows) = largest number of cells around a dof
    int gmap_space = -1;
    double* gvol = nullptr;          // scalar P1, 2-D: xgrid[CellVolumes] in visit order [row][lane] (one more plane of the gather map, femb_p1.cu)
    bool gvol_valid = false;         // false after the volumes or the map changed
    // general H1 gather maps (P2, ...; femb_pattern.cu): hm = the single scalar block of a one-space layout (Example200), hm_vec = the
    // component-diagonal part of block (0,0) of a vector-valued space in a multi-space layout (Example210's Laplacian)
    FembHMap hm, hm_vec;
    double* cellrhs = nullptr;     // [cell][qp][comp]: factor w_q |T| f(x_q), pre-pass of the Identity rhs gather
    size_t cellrhs_n = 0;
    double* p2vrhs = nullptr;      // [cell][6][2]: local entries of a two-component load on P2 triangles (fused rhs of Example210, femb_p2.cu)
    double* threc = nullptr;       // [cell][4]: |T| d_c lambda_a, a = 1, 2 (femb_ns.cu, owner-computes div-pressure blocks)
    double* rt0rec = nullptr;      // [cell][16]: local RT0 mass matrix, one 32-byte column per face dof (femb_hdiv.cu)
    double* p2rec = nullptr;       // [cell][28]: per-cell record of the closed-form P2 gather (femb_p2.cu)
    double* cellcoef = nullptr;    // [cell][FEMB_COEF_STRIDE]: Kc (3 / 6) then |T| f(x_q) per quadrature point (pre-pass of the gather)

    double* nzval = nullptr;      // nnz
    bool zero_pending_A = false;  // femb_matrix_zero is lazy: write-once kernels skip the memset
    uint64_t zero_mask = ~0ull;   // while zero_pending_A: layout blocks (bit brow * FEMB_MAX_SPACES + bcol) still awaiting the zero; a write-once
                                  // kernel that overwrites every entry of a block claims it (femb_claim_block), femb_flush_zero zeroes the rest
    bool zero_pending_b = false;
    bool dirty_A = false, dirty_b = false;  // something was assembled / set since the last femb_matrix_zero / femb_vector_zero
    uint8_t* touched = nullptr;   // nnz (only when tracking)
    bool track = false;
    double* bvec = nullptr;       // nrows
    double* sol = nullptr;        // ncols (current solution)
    int32_t* errflag = nullptr;   // device int: contributions that found no slot

    FembEntities ent[2];

    // Dirichlet data and the row-wise view of the pattern for A x - b (femb_system.cu)
    int32_t* bdofs = nullptr;      // last femb_boundary_dofs result: sorted, 1-based, local to its space
    int64_t nbdofs = 0;
    int64_t* fixed = nullptr;      // rows fixed by the last femb_dirichlet_apply (0-based, matrix numbering)
    int64_t nfixed = 0;
    int64_t* csr_ptr = nullptr;    // nrows + 1
    int32_t* csr_col = nullptr;    // nnz: column of each entry, rows contiguous
    uint32_t* csr_slot = nullptr;  // nnz: its position in nzval
    double* resvec = nullptr;      // nrows (+ 1 scratch)
    double* dirval = nullptr;      // nrows: boundary data of femb_interpolate (used by femb_dirichlet_apply when values == NULL)
    double* saved_nzval = nullptr; // femb_system_save: Alin / blin of Example210:164-165
    double* saved_b = nullptr;

    // colouring of cells (FEMB_SCATTER_COLORED)
    int ncolors = 0;
    int32_t* color_perm = nullptr;          // cells sorted by colour
    std::vector<int64_t> color_ptr;

    // pinned staging
    void* pinned = nullptr;
    size_t pinned_bytes = 0;

    // extra structural entries for the next pattern build (multi-GPU halo rows), keys col * nrows + row (0-based)
    std::vector<int64_t> extra_entries;

    // NCCL (bound at run time with dlopen, femb_comm.cu)
    void* comm = nullptr;
    int nranks = 1, rank = 0;
    std::vector<int> peers;
    std::vector<int64_t> send_ptr, recv_ptr, sendb_ptr, recvb_ptr;  // per peer, in elements
    int64_t* send_slots = nullptr;  // device
    int64_t* recv_slots = nullptr;
    int64_t* send_b = nullptr;
    int64_t* recv_b = nullptr;
    double* sendbuf = nullptr;      // [matrix part | vector part] per peer, contiguous
    double* recvbuf = nullptr;
    int64_t nsend = 0, nrecv = 0, nsendb = 0, nrecvb = 0;
    int64_t exchange_nnz = -1;      // nnz of the pattern the plan was validated against
    // fused (overlapped) exchange: dof ranges that must be assembled before the exchange may start
    bool fused_exchange = false;
    std::vector<int64_t> xslices;   // sorted: the 32-column slices that hold a ghost (sent) or an owned interface (received) dof
    int exchange_version = 0;       // bumped by every change of the plan (cached per-pattern launch lists check it)
    bool exchange_done = false;     // set by the assembly path that performed the (overlapped) exchange itself
    int64_t xdof_send_lo = 0, xdof_send_hi = -1, xdof_recv_lo = 0, xdof_recv_hi = -1;  // [lo, hi] of send_b / recv_b
    cudaStream_t stream_x = nullptr;
    cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr;
};

int femb_fail(femb_ctx* ctx, int code, const char* fmt, ...);

#define FEMB_CHECK_CTX(ctx) \
    do {                    \
        if (!(ctx)) return FEMB_ERR_ARG; \
    } while (0)

#define CUDA_TRY(ctx, expr)                                                                     \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return femb_fail((ctx), FEMB_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                             cudaGetErrorString(_e));                                            \
    } while (0)

#define FEMB_TRY(expr)              \
    do {                            \
        int _s = (expr);            \
        if (_s != FEMB_OK) return _s; \
    } while (0)

// after a kernel launch
#define KERNEL_CHECK(ctx)                          \
    do {                                           \
        (ctx)->launches++;                         \
        (ctx)->last_launches++;                    \
        CUDA_TRY((ctx), cudaGetLastError());       \
    } while (0)

template <typename T>
static inline int femb_alloc(femb_ctx* ctx, T** p, size_t n) {
    if (*p) {
        cudaFree(*p);
        *p = nullptr;
    }
    if (n == 0) return FEMB_OK;
    CUDA_TRY(ctx, cudaMalloc((void**)p, n * sizeof(T)));
    return FEMB_OK;
}

template <typename T>
static inline void femb_free(T** p) {
    if (*p) cudaFree(*p);
    *p = nullptr;
}

// host -> device copy of n elements through the ctx stream (pageable or registered memory)
template <typename T>
static inline int femb_h2d(femb_ctx* ctx, T* dst, const T* src, size_t n) {
    if (n == 0) return FEMB_OK;
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return FEMB_OK;
}
template <typename T>
static inline int femb_d2h(femb_ctx* ctx, T* dst, const T* src, size_t n) {
    if (n == 0) return FEMB_OK;
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return FEMB_OK;
}

// femb_api.cu
int femb_cellnodes_compare_async(femb_ctx* ctx, const int32_t* cellnodes, cudaStream_t stream);
int femb_cellnodes_differ(femb_ctx* ctx, bool* differ);
// pattern.cu
int femb_build_adjacency(femb_ctx* ctx, int space_id);
int femb_build_slotmaps(femb_ctx* ctx, int only_brow = -1, int only_bcol = -1);  // every block, or one
int femb_build_gmap(femb_ctx* ctx);
int femb_build_hmap(femb_ctx* ctx);
int femb_build_hmap_vec(femb_ctx* ctx);  // lazy: component-diagonal map of block (0,0) of a vector-valued H1 space
void femb_free_hmap(FembHMap& H);
int femb_hmap_set_canon(femb_ctx* ctx, FembHMap& H, bool canon);  // permutes the map entries in place between local dof order and P2 role order
void femb_p2_role_perm(uint8_t perm[12][12], int dim);             // femb_p2.cu: perm[k][role] = local dof
#define FEMB_COEF_STRIDE 16  // doubles per cell in femb_ctx::cellcoef (6 + up to 10 quadrature points)
void femb_free_pattern(femb_ctx* ctx);
void femb_free_space(FembSpace& s);
// femb_grid.cu
int femb_build_entities(femb_ctx* ctx, int kind);
void femb_free_entities(femb_ctx* ctx);
// femb_system.cu
void femb_free_csr(femb_ctx* ctx);
void femb_free_saved(femb_ctx* ctx);
void femb_free_system(femb_ctx* ctx);
// assemble.cu
int femb_flush_zero(femb_ctx* ctx, bool mat, bool vec);
bool femb_claim_block(femb_ctx* ctx, int brow, int bcol);   // true: the caller overwrites EVERY pattern entry of the block instead of adding to it
int femb_settle_partial_zero(femb_ctx* ctx);                // afterwards zero_pending_A means "the whole matrix" again (single-kernel paths)
int femb_block_entries(femb_ctx* ctx, int brow, int bcol, int64_t* n);
inline void femb_reset_block_counts(femb_ctx* ctx) {
    for (auto& r : ctx->block_nnz)
        for (auto& v : r) v = -1;
}
int femb_build_coloring(femb_ctx* ctx);
int femb_timer_begin(femb_ctx* ctx);
int femb_timer_end(femb_ctx* ctx);
// comm.cu: pack -> NCCL send/recv -> add on `stream` (no timing, no lazy-zero flush)
int femb_exchange_on(femb_ctx* ctx, cudaStream_t stream);
void femb_free_exchange(femb_ctx* ctx);
