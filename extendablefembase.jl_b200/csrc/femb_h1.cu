// Scalar (and component-blocked vector) H1 stiffness on affine cells for elements without per-cell handlers (P2, ...):
// the FP64 tensor-core reference-tensor kernel (atomic scatter) and the owner-computes gather (write-once).
#include "femb_kernels.cuh"

// ---------------------------------------------------------------------------------------------------
// scalar H1 stiffness on affine cells as a batched small GEMM on the FP64 tensor cores (P2, and P1 when the
// atomic scatter policy is chosen).  For an affine cell the push-forward factors out of the quadrature sum
// (feevaluator_h1.jl:69-72 with M = A^{-T} constant on the cell):
//     Aloc[j,k] = mu |T| sum_q w_q (M D_q e_j).(M D_q e_k) = sum_{a<=b} Kc[ab] * T[ab][j,k]
//     Kc = mu |T| (M^T M) = mu |T| (C^T C) / det^2  (3 or 6 coefficients per cell, C = cofactor matrix)
//     T[aa] = sum_q w_q D_q[.,a] D_q[.,a]^T,  T[ab] = sum_q w_q (D_q[.,a] D_q[.,b]^T + D_q[.,b] D_q[.,a]^T)
// i.e. Aloc (cells x unique entries) = Kc (cells x ncoef) * T (ncoef x unique entries): a dense GEMM with a
// reference tensor T that is the same for all cells.  One warp handles 32 cells per iteration as 4 m-tiles of
// mma.sync.m8n8k4.f64 (SASS DMMA): A fragments = Kc staged through shared memory, B fragments = T held in
// registers for the whole kernel, C = 8x8 tiles of local-matrix entries, scattered through the cell->slot map
// with red.global.add.f64 (upper triangle + mirror, Example200:149-158, |v| <= drop_tol skipped).
// ---------------------------------------------------------------------------------------------------
struct DmmaArgs {
    const double* coords;
    const int32_t* cellnodes;
    const double* cellvol;
    int64_t ncells;
    const double* T;       // [KS*4][NT*8]
    const int32_t* slots;  // [cell][nd][nd]
    double* nzval;
    uint8_t* touched;
    int32_t* errflag;
    double mu, drop_tol;
    int nd, ne;            // scalar local dofs nds, unique entries nds(nds+1)/2
    int ncomp;             // vector-valued H1: component-blocked basis (h1_p2.jl:133-141) -> the same scalar block on every diagonal block
};

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int DIM, int NT>
__global__ void __launch_bounds__(128) k_h1_stiffness_dmma(const DmmaArgs a) {
    constexpr int KS = (DIM == 2) ? 1 : 2;  // coefficient count 3 / 6 padded to 4 / 8
    constexpr int KP = KS * 4;
    __shared__ double Kst[4][32][KP + 1];   // per warp: Kc of its 32 cells (+1: no bank conflicts on the column reads)
    __shared__ uint8_t ejk[NT * 8][2];      // unique entry -> (j, k), j <= k
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int e = threadIdx.x; e < NT * 8; e += blockDim.x) {
        int j = 0, k = 0, idx = e;
        if (e < a.ne) {  // row-major upper triangle: (0,0),(0,1),...,(0,nd-1),(1,1),...
            while (idx >= a.nd - j) {
                idx -= a.nd - j;
                j++;
            }
            k = j + idx;
        }
        ejk[e][0] = (uint8_t)j;
        ejk[e][1] = (uint8_t)k;
    }
    double B[KS][NT];
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) B[ks][nt] = __ldg(a.T + (size_t)(ks * 4 + t) * (NT * 8) + nt * 8 + g);
    __syncthreads();
    const int64_t ntiles = (a.ncells + 31) / 32;
    const int per = a.nd * a.nd * a.ncomp * a.ncomp;
    for (int64_t tile = (int64_t)blockIdx.x * 4 + warp; tile < ntiles; tile += (int64_t)gridDim.x * 4) {
        const int64_t cell = tile * 32 + lane;
        double kc[KP];
#pragma unroll
        for (int i = 0; i < KP; i++) kc[i] = 0.0;
        if (cell < a.ncells) {
            double x[DIM + 1][DIM];
            load_nodes<DIM>(a.coords, a.cellnodes, cell, x);
            double A[DIM][DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++)
#pragma unroll
                for (int i = 0; i < DIM; i++) A[d][i] = x[i + 1][d] - x[0][d];
            double C[DIM][DIM], det;  // cofactors: M[k][j] = C[k][j] / det
            if (DIM == 2) {
                det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
                C[0][0] = A[1][1]; C[0][1] = -A[1][0];
                C[1][0] = -A[0][1]; C[1][1] = A[0][0];
            } else {
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const int ra = (r + 1) % 3, rb = (r + 2) % 3, ja = (j + 1) % 3, jb = (j + 2) % 3;
                        C[r][j] = A[ra][ja] * A[rb][jb] - A[ra][jb] * A[rb][ja];
                    }
                det = A[0][0] * C[0][0] + A[0][1] * C[0][1] + A[0][2] * C[0][2];
            }
            const double vol = a.cellvol ? __ldg(a.cellvol + cell) : fabs(det) * (DIM == 2 ? 0.5 : 1.0 / 6.0);
            const double sc = (a.mu * vol) / (det * det);
            // Kmat[p][q] = sc * sum_k C[k][p] C[k][q]; coefficient order (00, 11, [22,] 01, [02, 12])
            int ci = 0;
#pragma unroll
            for (int p = 0; p < DIM; p++) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; k++) s += C[k][p] * C[k][p];
                kc[ci++] = s * sc;
            }
#pragma unroll
            for (int p = 0; p < DIM; p++)
#pragma unroll
                for (int q = p + 1; q < DIM; q++) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; k++) s += C[k][p] * C[k][q];
                    kc[ci++] = s * sc;
                }
        }
#pragma unroll
        for (int i = 0; i < KP; i++) Kst[warp][lane][i] = kc[i];
        __syncwarp();
#pragma unroll 1
        for (int mt = 0; mt < 4; mt++) {
            double acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                const double af = Kst[warp][mt * 8 + g][ks * 4 + t];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) dmma_m8n8k4(acc[nt][0], acc[nt][1], af, B[ks][nt]);
            }
            const int64_t c2 = tile * 32 + mt * 8 + g;
            if (c2 < a.ncells) {
                const int ndl = a.nd * a.ncomp;  // local dofs of the (vector) space
                for (int comp = 0; comp < a.ncomp; comp++) {
                    const int32_t* sl = a.slots + c2 * per + (comp * a.nd) * ndl + comp * a.nd;
                    // all slot loads of the m-tile first (independent, in flight together), then the reductions
                    int32_t s_up[NT][2], s_lo[NT][2];
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int e = nt * 8 + 2 * t + h;
                            const int j = ejk[e][0], k = ejk[e][1];
                            const bool on = e < a.ne && fabs(acc[nt][h]) > a.drop_tol;
                            s_up[nt][h] = on ? __ldg(sl + j * ndl + k) : INT32_MIN;
                            s_lo[nt][h] = (on && k != j) ? __ldg(sl + k * ndl + j) : INT32_MIN;
                        }
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const double v = acc[nt][h];
                            if (s_up[nt][h] != INT32_MIN) scatter_add(a.nzval, a.touched, a.errflag, s_up[nt][h], v, 1);
                            if (s_lo[nt][h] != INT32_MIN) scatter_add(a.nzval, a.touched, a.errflag, s_lo[nt][h], v, 1);
                        }
                }
            }
        }
        __syncwarp();
    }
}

// reference tensor T[coef][entry] from the evaluator's reference-gradient table and weights (host, once)
static int build_reftensor(femb_ctx* ctx, FembEval& E, int nd, int ncomp, int nt) {
    if (E.reftensor && E.reftensor_nt == nt) return FEMB_OK;
    const int dim = ctx->dim, ks = dim == 2 ? 1 : 2, ncol = nt * 8;
    const FembSpace& s = ctx->spaces[E.space];
    if (E.h_refderivs.size() != (size_t)E.nqp * s.nd_all * ncomp * dim) return femb_fail(ctx, FEMB_ERR_ARG, "Gradient evaluator has no reference-derivative table");
    std::vector<double> T((size_t)ks * 4 * ncol, 0.0);
    // coefficient order (00, 11, [22,] 01, [02, 12])
    std::vector<std::pair<int, int>> co;
    for (int p = 0; p < dim; p++) co.push_back({p, p});
    for (int p = 0; p < dim; p++)
        for (int q = p + 1; q < dim; q++) co.push_back({p, q});
    int e = 0;
    for (int j = 0; j < nd; j++)
        for (int k = j; k < nd; k++, e++)
            for (size_t c = 0; c < co.size(); c++) {
                const int p = co[c].first, q = co[c].second;
                double sum = 0.0;
                for (int qp = 0; qp < E.nqp; qp++) {
                    // component 0 of the first nd functions = the scalar basis (component-blocked vector spaces)
                    const double* Dj = &E.h_refderivs[((size_t)qp * s.nd_all + j) * ncomp * dim];
                    const double* Dk = &E.h_refderivs[((size_t)qp * s.nd_all + k) * ncomp * dim];
                    sum += E.w[qp] * (p == q ? Dj[p] * Dk[p] : Dj[p] * Dk[q] + Dj[q] * Dk[p]);
                }
                T[c * ncol + e] = sum;
            }
    FEMB_TRY(femb_alloc(ctx, &E.reftensor, T.size()));
    FEMB_TRY(femb_h2d(ctx, E.reftensor, T.data(), T.size()));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    E.reftensor_nt = nt;
    return FEMB_OK;
}

// returns FEMB_ERR_UNSUPPORTED (without setting an error) when the DMMA path does not cover the space
int launch_stiffness_dmma(femb_ctx* ctx, int eg, double mu, double drop_tol) {
    FembEval& EG = ctx->evals[eg];
    const FembSpace& s = ctx->spaces[EG.space];
    if (s.family != FEMB_FAMILY_H1 || s.handler != FEMB_HANDLER_NONE || s.nd != s.nd_all || s.nd % s.ncomp) return FEMB_ERR_UNSUPPORTED;
    const int nds = s.nd / s.ncomp;
    if (nds > 10) return FEMB_ERR_UNSUPPORTED;
    int brow = -1;
    for (int i = 0; i < ctx->nrs; i++)
        if (ctx->row_spaces[i] == EG.space && i < ctx->ncs && ctx->col_spaces[i] == EG.space) brow = i;
    if (brow < 0 || !find_block(ctx, brow, brow)) return FEMB_ERR_UNSUPPORTED;
    FEMB_TRY(femb_build_slotmaps(ctx));
    const FembBlock* blk = find_block(ctx, brow, brow);
    if (!blk || !blk->slots) return FEMB_ERR_UNSUPPORTED;
    const int ne = nds * (nds + 1) / 2, nt = (ne + 7) / 8;
    FEMB_TRY(build_reftensor(ctx, EG, nds, s.ncomp, nt));
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    DmmaArgs a;
    a.coords = ctx->coords; a.cellnodes = ctx->cellnodes; a.cellvol = ctx->cellvol; a.ncells = ctx->ncells; a.T = EG.reftensor; a.slots = blk->slots;
    a.nzval = ctx->nzval; a.touched = ctx->touched; a.errflag = ctx->errflag; a.mu = mu; a.drop_tol = drop_tol; a.nd = nds; a.ne = ne; a.ncomp = s.ncomp;
    if (ctx->ncells == 0) return FEMB_OK;
    const int64_t ntiles = (ctx->ncells + 31) / 32;
    const unsigned grid = (unsigned)std::min<int64_t>((ntiles + 3) / 4, (int64_t)ctx->sm_count * 16);
#define FEMB_DMMA(D, N) k_h1_stiffness_dmma<D, N><<<grid, 128, 0, ctx->stream>>>(a)
    if (ctx->dim == 2 && nt == 1) FEMB_DMMA(2, 1);
    else if (ctx->dim == 2 && nt == 3) FEMB_DMMA(2, 3);
    else if (ctx->dim == 3 && nt == 2) FEMB_DMMA(3, 2);
    else if (ctx->dim == 3 && nt == 7) FEMB_DMMA(3, 7);
    else return FEMB_ERR_UNSUPPORTED;
#undef FEMB_DMMA
    KERNEL_CHECK(ctx);
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// scalar H1 stiffness + rhs, owner-computes by matrix column for general elements (P2, ...): write-once,
// bit-reproducible, no atomics.  Two launches per assembly:
//   k_cell_coef   (one thread per cell)  Kc = mu |T| (C^T C) / det^2 (3 / 6 coefficients) and |T| f(x_q): 80 B / cell
//   k_h1_gather   (one thread per column c) for every cell e adjacent to c (gather map, femb_pattern.cu) with local
//                 index k:  h[q] = w_q Kmat D_q[:,k]  (the cell-dependent factor moved to the column side), then
//                 Aloc[j,k] = sum_q D_q[:,j] . h[q]  with UNIFORM table operands D_q[:,j] (compile-time indices into the
//                 kernel-parameter table), accumulated in shared memory rows and written once.
// The slices are launched in ranges of similar column length (vertex vs edge dofs) so that shared memory is sized per
// range.  Column positions are 8-bit, the base comes from colptr (int64): nnz may exceed 2^31.
// ---------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128) k_cell_coef(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol,
                                                   int64_t ncells, double mu, RhsDev rhs, QuadDev q, double* __restrict__ out) {
    __shared__ double tile[128][FEMB_COEF_STRIDE + 1];  // staged so that the 128-byte records leave the SM fully coalesced
    const int64_t cell0 = blockIdx.x * (int64_t)blockDim.x;
    const int64_t cell = cell0 + threadIdx.x;
    if (cell < ncells) {
        double x[DIM + 1][DIM];
        load_nodes<DIM>(coords, cellnodes, cell, x);
        double A[DIM][DIM], C[DIM][DIM], det;
#pragma unroll
        for (int d = 0; d < DIM; d++)
#pragma unroll
            for (int i = 0; i < DIM; i++) A[d][i] = x[i + 1][d] - x[0][d];
        if (DIM == 2) {
            det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
            C[0][0] = A[1][1]; C[0][1] = -A[1][0];
            C[1][0] = -A[0][1]; C[1][1] = A[0][0];
        } else {
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const int ra = (r + 1) % 3, rb = (r + 2) % 3, ja = (j + 1) % 3, jb = (j + 2) % 3;
                    C[r][j] = A[ra][ja] * A[rb][jb] - A[ra][jb] * A[rb][ja];
                }
            det = A[0][0] * C[0][0] + A[0][1] * C[0][1] + A[0][2] * C[0][2];
        }
        const double vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * (DIM == 2 ? 0.5 : 1.0 / 6.0);
        const double sc = (mu * vol) / (det * det);
        double* o = tile[threadIdx.x];
        int ci = 0;
#pragma unroll
        for (int p = 0; p < DIM; p++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; k++) s += C[k][p] * C[k][p];
            o[ci++] = s * sc;
        }
#pragma unroll
        for (int p = 0; p < DIM; p++)
#pragma unroll
            for (int r = p + 1; r < DIM; r++) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; k++) s += C[k][p] * C[k][r];
                o[ci++] = s * sc;
            }
        for (; ci < 6; ci++) o[ci] = 0.0;
        for (int qp = 0; qp < FEMB_COEF_STRIDE - 6; qp++) {
            double f = 0.0;
            if (qp < q.nqp && rhs.kind == FEMB_RHS_AFFINE) {  // f(x_qp), x_qp = A xref_qp + x_1
                f = rhs.p[0];
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                    double sx = x[0][d];
#pragma unroll
                    for (int i = 0; i < DIM; i++) sx += A[d][i] * q.xref[qp * DIM + i];
                    f += rhs.p[1 + d] * sx;
                }
            } else if (qp < q.nqp && rhs.kind == FEMB_RHS_QPVALUES) {
                f = __ldg(rhs.fvals + (size_t)cell * q.nqp + qp);
            }
            o[6 + qp] = f * vol;
        }
    }
    __syncthreads();
    const int64_t nloc = min((int64_t)128, ncells - cell0);
    double* dst = out + cell0 * FEMB_COEF_STRIDE;
    for (int i = threadIdx.x; i < nloc * FEMB_COEF_STRIDE; i += 128) dst[i] = tile[i / FEMB_COEF_STRIDE][i % FEMB_COEF_STRIDE];
}

// 256-bit read-only load (sm_100: LDG.E.256).  A scattered 128-byte coefficient record costs one L1 tag look-up per
// instruction and lane, and the L1 pipe is what bounds the gather: fewer, wider loads relieve it.
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// coefficient record of one cell (k_cell_coef): Kc[0..5], |T| f(x_q) at [6, 6 + NQ)
template <int DIM, int KP, int NQ>
__device__ __forceinline__ void load_coef(const double* __restrict__ cc, bool has_rhs, double (&kc)[KP], double (&f)[NQ]) {
    if (DIM == 3 && NQ == 4) {
        ldg256(cc, kc[0], kc[1], kc[2], kc[3]);
        if (has_rhs) {
            ldg256(cc + 4, kc[4], kc[5], f[0], f[1]);
            const double2 fv = __ldg(reinterpret_cast<const double2*>(cc + 8));
            f[2] = fv.x; f[3] = fv.y;
        } else {
            const double2 cv = __ldg(reinterpret_cast<const double2*>(cc + 4));
            kc[4] = cv.x; kc[5] = cv.y;
        }
    } else {
        if (DIM == 2) {
            double unused;
            ldg256(cc, kc[0], kc[1], kc[2], unused);
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double2 cv = __ldg(reinterpret_cast<const double2*>(cc) + i);
                kc[2 * i] = cv.x; kc[2 * i + 1] = cv.y;
            }
        }
        if (has_rhs) {
#pragma unroll
            for (int qp = 0; qp < NQ; qp++) f[qp] = __ldg(cc + 6 + qp);
        }
    }
}

template <int DIM, int ND, int NQ>
struct HTab {
    double D[NQ][DIM][ND];  // reference gradients d phi_j / d xref_a at xref_q
    double wphi[NQ][ND];    // w_q phi_k(xref_q)
    double w[NQ];
};

struct HArgs {
    const double* coef;
    const int32_t* hslice_ptr;
    const uint32_t* hmap;
    size_t hstride;
    const int64_t* colptr;
    int64_t ndofs;
    double* nzval;
    uint8_t* touched;
    double* bvec;
    int32_t* errflag;
    double drop_tol;
    int acc_A, acc_b, has_rhs;
    int64_t slice_begin, slice_end;
};

// SPLIT = 1: one thread per column, 4 slices (warps) per CTA.  SPLIT = 4: one slice per CTA, the 4 warps share the
// columns of the slice — warp w takes the visits r = w (mod 4), computes them in parallel and the shared-memory
// accumulation is serialised in visit order between barriers, so the result is bit-identical to SPLIT = 1.
// FAST: no value filter (drop_tol <= 0), no slot tracking and a complete map (no missing positions): the accumulation is a
// branch-free load / add / store per row instead of ~23 instructions of filter, tracking and error bookkeeping.
template <int DIM, int ND, int NQ, int SPLIT, bool FAST>
__global__ void __launch_bounds__(128) k_h1_gather(const HArgs a, const HTab<DIM, ND, NQ> tab) {
    constexpr unsigned ncol = 128 / SPLIT;  // columns per CTA
    constexpr int NW = (ND + 3) / 4;
    constexpr int KP = (DIM == 2) ? 3 : 6;
    extern __shared__ double acc[];  // [maxlen of the range][ncol]
    __shared__ double bpart[SPLIT == 1 ? 1 : 4][32];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int64_t slice = a.slice_begin + (SPLIT == 1 ? (int64_t)blockIdx.x * 4 + warp : (int64_t)blockIdx.x);
    if (SPLIT == 1 && slice >= a.slice_end) return;  // (SPLIT == 4: the grid is exact)
    const int64_t c = slice * 32 + lane;
    const bool valid = c < a.ndofs;
    const int row0 = __ldg(a.hslice_ptr + slice), nrows = __ldg(a.hslice_ptr + slice + 1) - row0;
    int64_t base = 0;
    int len = 0;
    if (valid) {
        base = a.colptr[c] - 1;
        len = (int)(a.colptr[c + 1] - 1 - base);
    }
    const unsigned sacc = (unsigned)__cvta_generic_to_shared(acc) + (SPLIT == 1 ? tid : lane) * 8u;
    for (int i = (SPLIT == 1 ? 0 : (int)warp); i < len; i += SPLIT) sts64(sacc + i * (ncol * 8u), 0.0);
    if (SPLIT > 1) __syncthreads();
    double bsum = 0.0;
    unsigned err = 0u;
    const int niter = (nrows + SPLIT - 1) / SPLIT;
    // software pipeline: the gather-map entry two visits ahead and the cell coefficients of the next visit are in
    // flight while the current visit is computed (padding entries point at cell 0: their loads stay in bounds)
    const int rfirst = (SPLIT == 1) ? 0 : (int)warp;
    const uint32_t* __restrict__ hbase = a.hmap + ((size_t)row0 << 5) + lane;
    unsigned k1 = 0xFFFFFFFFu, cell1 = 0u, k2 = 0xFFFFFFFFu, cell2 = 0u, w1[NW], w2[NW];
#pragma unroll
    for (int wd = 0; wd < NW; wd++) w1[wd] = w2[wd] = 0xFFFFFFFFu;
    if (rfirst < nrows) {
        const uint32_t* hp = hbase + ((size_t)rfirst << 5);
        k1 = __ldg(hp + a.hstride);
        cell1 = __ldg(hp);
#pragma unroll
        for (int wd = 0; wd < NW; wd++) w1[wd] = __ldg(hp + (2 + wd) * a.hstride);
    }
    if (rfirst + SPLIT < nrows) {
        const uint32_t* hp = hbase + ((size_t)(rfirst + SPLIT) << 5);
        k2 = __ldg(hp + a.hstride);
        cell2 = __ldg(hp);
#pragma unroll
        for (int wd = 0; wd < NW; wd++) w2[wd] = __ldg(hp + (2 + wd) * a.hstride);
    }
    double kc1[KP], f1[NQ];
    {
#pragma unroll
        for (int qp = 0; qp < NQ; qp++) f1[qp] = 0.0;
        load_coef<DIM, KP, NQ>(a.coef + (size_t)cell1 * FEMB_COEF_STRIDE, a.has_rhs != 0, kc1, f1);
    }
    for (int it = 0; it < niter; it++) {
        const int r = it * SPLIT + rfirst;
        const unsigned k = k1;
        const bool on = k != 0xFFFFFFFFu;
        double v[ND], kc[KP], fq[NQ];
        unsigned words[NW];
#pragma unroll
        for (int wd = 0; wd < NW; wd++) words[wd] = w1[wd];
#pragma unroll
        for (int i = 0; i < KP; i++) kc[i] = kc1[i];
#pragma unroll
        for (int qp = 0; qp < NQ; qp++) fq[qp] = f1[qp];
        // rotate the pipeline
        k1 = k2; cell1 = cell2;
#pragma unroll
        for (int wd = 0; wd < NW; wd++) w1[wd] = w2[wd];
        if (r + 2 * SPLIT < nrows) {
            const uint32_t* hp = hbase + ((size_t)(r + 2 * SPLIT) << 5);
            k2 = __ldg(hp + a.hstride);
            cell2 = __ldg(hp);
#pragma unroll
            for (int wd = 0; wd < NW; wd++) w2[wd] = __ldg(hp + (2 + wd) * a.hstride);
        } else {
            k2 = 0xFFFFFFFFu; cell2 = 0u;
        }
        load_coef<DIM, KP, NQ>(a.coef + (size_t)cell1 * FEMB_COEF_STRIDE, a.has_rhs != 0, kc1, f1);
        if (on) {
            double Km[DIM][DIM];  // symmetric Kmat from the coefficient order (00, 11, [22,] 01, [02, 12])
            if (DIM == 2) {
                Km[0][0] = kc[0]; Km[1][1] = kc[1]; Km[0][1] = Km[1][0] = kc[2];
            } else {
                Km[0][0] = kc[0]; Km[1][1] = kc[1]; Km[2][2] = kc[2];
                Km[0][1] = Km[1][0] = kc[3]; Km[0][2] = Km[2][0] = kc[4]; Km[1][2] = Km[2][1] = kc[5];
            }
            double h[NQ][DIM];
#pragma unroll
            for (int qp = 0; qp < NQ; qp++) {
                double dk[DIM];
#pragma unroll
                for (int b = 0; b < DIM; b++) dk[b] = tab.D[qp][b][k];  // per-lane k: indexed constant-bank load
#pragma unroll
                for (int aa = 0; aa < DIM; aa++) {
                    double s2 = 0.0;
#pragma unroll
                    for (int b = 0; b < DIM; b++) s2 += Km[aa][b] * dk[b];
                    h[qp][aa] = s2 * tab.w[qp];
                }
            }
#pragma unroll
            for (int j = 0; j < ND; j++) {
                double vj = 0.0;
#pragma unroll
                for (int qp = 0; qp < NQ; qp++)
#pragma unroll
                    for (int aa = 0; aa < DIM; aa++) vj += tab.D[qp][aa][j] * h[qp][aa];
                v[j] = vj;
            }
            if (a.has_rhs) {
                double t = 0.0;
#pragma unroll
                for (int qp = 0; qp < NQ; qp++) t += tab.wphi[qp][k] * fq[qp];
                bsum += t;
            }
        }
#pragma unroll 1
        for (int ph = 0; ph < SPLIT; ph++) {
            if (FAST && on && (SPLIT == 1 || (int)warp == ph)) {
#pragma unroll
                for (int j = 0; j < ND; j++) {
                    const unsigned ad = sacc + ((words[j >> 2] >> (8 * (j & 3))) & 0xFFu) * (ncol * 8u);
                    sts64(ad, lds64(ad) + v[j]);
                }
            }
            if (!FAST && on && (SPLIT == 1 || (int)warp == ph)) {
#pragma unroll
                for (int j = 0; j < ND; j++) {
                    const unsigned pos = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                    if (fabs(v[j]) > a.drop_tol) {
                        if (pos == 0xFFu) {
                            err = 1u;
                        } else {
                            const unsigned ad = sacc + pos * (ncol * 8u);
                            sts64(ad, lds64(ad) + v[j]);
                            if (a.touched) a.touched[base + pos] = 1;
                        }
                    }
                }
            }
            if (SPLIT > 1) __syncthreads();
        }
    }
    if (err) atomicAdd(a.errflag, 1);
    if (SPLIT > 1) {
        bpart[warp][lane] = bsum;
        __syncthreads();
        bsum = ((bpart[0][lane] + bpart[1][lane]) + bpart[2][lane]) + bpart[3][lane];
    }
    if (!valid) return;
    double* out = a.nzval + base;
    if (a.acc_A) {
        for (int i = (SPLIT == 1 ? 0 : (int)warp); i < len; i += SPLIT) out[i] += lds64(sacc + i * (ncol * 8u));
    } else {
        for (int i = (SPLIT == 1 ? 0 : (int)warp); i < len; i += SPLIT) out[i] = lds64(sacc + i * (ncol * 8u));
    }
    if (a.has_rhs && (SPLIT == 1 || warp == 0)) {
        if (a.acc_b) a.bvec[c] += bsum;
        else a.bvec[c] = bsum;
    }
}

template <int DIM, int ND, int NQ>
static int launch_h1_gather_t(femb_ctx* ctx, const FembEval& EG, const FembEval* EI, HArgs a) {
    const FembSpace& s = ctx->spaces[EG.space];
    HTab<DIM, ND, NQ> tab;
    for (int q = 0; q < NQ; q++) {
        tab.w[q] = EG.w[q];
        for (int j = 0; j < ND; j++) {
            for (int d = 0; d < DIM; d++) tab.D[q][d][j] = EG.h_refderivs[((size_t)q * s.nd_all + j) * DIM + d];
            tab.wphi[q][j] = EI ? EG.w[q] * EI->h_refvals[(size_t)q * s.nd_all + j] : 0.0;
        }
    }
    const bool fast = ctx->hmap_complete && !a.touched && !(a.drop_tol > 0.0);
    auto kern1 = fast ? k_h1_gather<DIM, ND, NQ, 1, true> : k_h1_gather<DIM, ND, NQ, 1, false>;
    auto kern4 = fast ? k_h1_gather<DIM, ND, NQ, 4, true> : k_h1_gather<DIM, ND, NQ, 4, false>;
    size_t max1 = 0, max4 = 0;
    for (auto& r : ctx->hranges) {
        if (r.split == 4) max4 = std::max(max4, (size_t)std::max(1, r.maxlen) * 32 * sizeof(double));
        else max1 = std::max(max1, (size_t)std::max(1, r.maxlen) * 128 * sizeof(double));
    }
    if (max1 > 200 * 1024) return FEMB_ERR_UNSUPPORTED;
    // the accumulator rows are the occupancy limiter: ask for the largest shared-memory carve-out
    if (max1) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max1));
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    if (max4) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max4));
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern4, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    for (auto& r : ctx->hranges) {
        if (r.s1 <= r.s0) continue;
        a.slice_begin = r.s0;
        a.slice_end = r.s1;
        if (r.split == 4) {
            const size_t smem = (size_t)std::max(1, r.maxlen) * 32 * sizeof(double);
            kern4<<<(unsigned)(r.s1 - r.s0), 128, smem, ctx->stream>>>(a, tab);
        } else {
            const size_t smem = (size_t)std::max(1, r.maxlen) * 128 * sizeof(double);
            kern1<<<(unsigned)((r.s1 - r.s0 + 3) / 4), 128, smem, ctx->stream>>>(a, tab);
        }
        KERNEL_CHECK(ctx);
    }
    return FEMB_OK;
}

// returns FEMB_ERR_UNSUPPORTED (no error message) when the gather path does not cover the space / rule
int launch_h1_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol) {
    const FembEval& EG = ctx->evals[eg];
    const FembSpace& s = ctx->spaces[EG.space];
    if (!ctx->hmap || ctx->hmap_space != EG.space || EG.nqp > 10) return FEMB_ERR_UNSUPPORTED;
    if (rhs.kind != FEMB_RHS_NONE && rhs.kind != FEMB_RHS_AFFINE && rhs.kind != FEMB_RHS_QPVALUES) return FEMB_ERR_UNSUPPORTED;
    if (EG.h_refderivs.size() != (size_t)EG.nqp * s.nd_all * ctx->dim) return FEMB_ERR_UNSUPPORTED;
    const FembEval* EI = rhs.kind != FEMB_RHS_NONE ? &ctx->evals[ei] : nullptr;
    if (EI && EI->h_refvals.size() != (size_t)EG.nqp * s.nd_all) return FEMB_ERR_UNSUPPORTED;
    const int key = ctx->dim * 10000 + s.nd * 100 + EG.nqp;
    if (key != 20603 && key != 31004) return FEMB_ERR_UNSUPPORTED;  // H1P2 on triangles (3-point rule) / tetrahedra (4-point rule)
    if (!ctx->cellcoef) FEMB_TRY(femb_alloc(ctx, &ctx->cellcoef, (size_t)ctx->ncells * FEMB_COEF_STRIDE));
    if (ctx->ncells == 0) return FEMB_OK;
    QuadDev q = quad_dev(EG, ctx->dim);
    const unsigned gridc = (unsigned)((ctx->ncells + 127) / 128);
    if (ctx->dim == 2) k_cell_coef<2><<<gridc, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, mu, rhs, q, ctx->cellcoef);
    else k_cell_coef<3><<<gridc, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, mu, rhs, q, ctx->cellcoef);
    KERNEL_CHECK(ctx);
    HArgs a;
    a.coef = ctx->cellcoef; a.hslice_ptr = ctx->hslice_ptr; a.hmap = ctx->hmap; a.hstride = ctx->hmap_stride; a.colptr = ctx->colptr; a.ndofs = s.ndofs;
    a.nzval = ctx->nzval; a.touched = ctx->touched; a.bvec = ctx->bvec; a.errflag = ctx->errflag; a.drop_tol = drop_tol;
    a.acc_A = ctx->zero_pending_A ? 0 : 1;
    a.acc_b = ctx->zero_pending_b ? 0 : 1;
    a.has_rhs = rhs.kind != FEMB_RHS_NONE;
    a.slice_begin = a.slice_end = 0;
    ctx->zero_pending_A = false;
    if (a.has_rhs) ctx->zero_pending_b = false;
    if (key == 20603) return launch_h1_gather_t<2, 6, 3>(ctx, EG, EI, a);
    return launch_h1_gather_t<3, 10, 4>(ctx, EG, EI, a);
}
