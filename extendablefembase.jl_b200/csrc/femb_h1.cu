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

__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
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
    double D[NQ][DIM][ND];   // reference gradients d phi_j / d xref_a at xref_q (row side: uniform operands)
    double wD[NQ][DIM][ND];  // w_q D (column side: read with the per-lane local index k)
    double wphi[NQ][ND];     // w_q phi_k(xref_q)
    double L4[NQ][DIM + 1];  // P2 only: 4 lambda_i(xref_q), the barycentric coordinates of the quadrature points
};

#ifndef FEMB_H1_KTAB_SMEM
#define FEMB_H1_KTAB_SMEM 0
#endif

struct HArgs {
    const double* coef;
    const int32_t* hslice_ptr;
    const uint8_t* hslice_zero;
    const uint32_t* hmap;
    size_t hstride;
    const int64_t* colptr;
    const int32_t *sub0, *sublen;
    int64_t ndofs;
    double* nzval;
    uint8_t* touched;
    double* bvec;
    int32_t* errflag;
    double drop_tol;
    int acc_A, acc_b, has_rhs;
    int64_t slice_begin, slice_end;
};

// local edges of the P2 dofs DIM+1 .. (ExtendableGrids local edge / face numbering, h1_p2.jl): vertices (a, b), a < b
template <int DIM>
__device__ __host__ constexpr int p2_edge_a(int e) {
    return DIM == 2 ? (e == 0 ? 0 : (e == 1 ? 1 : 0)) : (e < 3 ? 0 : (e < 5 ? 1 : 2));
}
template <int DIM>
__device__ __host__ constexpr int p2_edge_b(int e) {
    return DIM == 2 ? (e == 0 ? 1 : 2) : (e < 3 ? e + 1 : (e == 3 ? 2 : 3));
}

// SPLIT = 1: one thread per column, 4 slices (warps) per CTA.  SPLIT = 4: one slice per CTA, the 4 warps share the
// columns of the slice — warp w takes the visits r = w (mod 4), computes them in parallel and the shared-memory
// accumulation is serialised in visit order between barriers (same sums in the same order as one thread per column).
// FAST: no value filter (drop_tol <= 0), no slot tracking and a complete map (no missing positions): the accumulation is
// branch-free, uses the map's first-touch flags (store instead of add: no zeroed accumulator, fewer shared-memory reads)
// and, with one thread per column, reads the accumulator rows of a visit before the FP64 work and sums on top of them.
// What bounds this kernel is the L1 data pipe (ncu: l1tex wavefronts 74-76 %): one wavefront per 32-byte sector of a
// scattered access, two per 64-bit shared-memory access of a warp — so the design minimises those: 256-bit loads of the
// coefficient record, 256-bit stores of the finished column, k and the flags packed into the position words.
template <int DIM, int ND, int NQ, int SPLIT, bool FAST, bool P2S>
__global__ void __launch_bounds__(128, 4) k_h1_gather(const HArgs a, const HTab<DIM, ND, NQ> tab) {
    constexpr unsigned ncol = 128 / SPLIT;  // columns per CTA
    constexpr int NW = (ND + 3) / 4;
    constexpr int KP = (DIM == 2) ? 3 : 6;
    constexpr unsigned PAD = 0xFFFFFFFFu;
    extern __shared__ double acc[];  // [maxlen of the range][ncol]
    __shared__ double bpart[SPLIT == 1 ? 1 : 4][32];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
#if FEMB_H1_KTAB_SMEM
    constexpr int KT_N = (NQ * DIM + NQ + 1) & ~1;  // column-side table row: w_q D_q[b][k] (NQ * DIM), w_q phi_k(x_q) (NQ)
    constexpr int KT_STRIDE = KT_N + 2;             // 16-byte aligned rows that start in different banks
    __shared__ __align__(16) double ktab[ND * KT_STRIDE];
    for (int i = tid; i < ND * KT_N; i += 128) {
        const int kk = i / KT_N, e = i % KT_N;
        double val = 0.0;
        if (e < NQ * DIM) val = tab.wD[e / DIM][e % DIM][kk];
        else if (e < NQ * DIM + NQ) val = tab.wphi[e - NQ * DIM][kk];
        ktab[kk * KT_STRIDE + e] = val;
    }
    __syncthreads();  // (before any early exit)
    const unsigned stab = (unsigned)__cvta_generic_to_shared(ktab);
#endif
    const int64_t slice = a.slice_begin + (SPLIT == 1 ? (int64_t)blockIdx.x * 4 + warp : (int64_t)blockIdx.x);
    if (SPLIT == 1 && slice >= a.slice_end) return;  // (SPLIT == 4: the grid is exact)
    const int64_t c = slice * 32 + lane;
    const bool valid = c < a.ndofs;
    const int row0 = __ldg(a.hslice_ptr + slice), nrows = __ldg(a.hslice_ptr + slice + 1) - row0;
    const bool zero_first = !FAST || __ldg(a.hslice_zero + slice) != 0;
    int64_t base = 0;
    int len = 0;
    if (valid) {
        base = a.colptr[c] - 1;
        len = (int)(a.colptr[c + 1] - 1 - base);
        if (a.sub0) {  // vector maps: the block's run of the column
            base += a.sub0[c];
            len = a.sublen[c];
        }
    }
    const unsigned sacc = (unsigned)__cvta_generic_to_shared(acc) + (SPLIT == 1 ? tid : lane) * 8u;
    if (zero_first) {
        for (int i = (SPLIT == 1 ? 0 : (int)warp); i < len; i += SPLIT) sts64(sacc + i * (ncol * 8u), 0.0);
    }
    if (SPLIT > 1) __syncthreads();
    const unsigned fmask = zero_first ? 0u : 0xFFFFFFFFu;  // zeroed accumulators: every visit adds
    double bsum = 0.0;
    unsigned err = 0u;
    const int niter = (nrows + SPLIT - 1) / SPLIT;
    // software pipeline: the gather-map entry two visits ahead and the cell coefficients of the next visit are in
    // flight while the current visit is computed (padding entries point at cell 0: their loads stay in bounds)
    const int rfirst = (SPLIT == 1) ? 0 : (int)warp;
    const uint32_t* __restrict__ hbase = a.hmap + ((size_t)row0 << 5) + lane;
    unsigned cell1 = 0u, cell2 = 0u, w1[NW], w2[NW];
#pragma unroll
    for (int wd = 0; wd < NW; wd++) w1[wd] = w2[wd] = PAD;
    if (rfirst < nrows) {
        const uint32_t* hp = hbase + ((size_t)rfirst << 5);
        cell1 = __ldg(hp);
#pragma unroll
        for (int wd = 0; wd < NW; wd++) w1[wd] = __ldg(hp + (1 + wd) * a.hstride);
    }
    if (rfirst + SPLIT < nrows) {
        const uint32_t* hp = hbase + ((size_t)(rfirst + SPLIT) << 5);
        cell2 = __ldg(hp);
#pragma unroll
        for (int wd = 0; wd < NW; wd++) w2[wd] = __ldg(hp + (1 + wd) * a.hstride);
    }
    double kc1[KP], f1[NQ];
    {
#pragma unroll
        for (int qp = 0; qp < NQ; qp++) f1[qp] = 0.0;
        load_coef<DIM, KP, NQ>(a.coef + (size_t)cell1 * FEMB_COEF_STRIDE, a.has_rhs != 0, kc1, f1);
    }
    for (int it = 0; it < niter; it++) {
        const int r = it * SPLIT + rfirst;
        double v[ND], kc[KP], fq[NQ];
        unsigned words[NW];
#pragma unroll
        for (int wd = 0; wd < NW; wd++) words[wd] = w1[wd];
        const unsigned k = (words[NW - 1] >> 16) & 0xFu;
        const unsigned first = (words[NW - 1] >> 20) & fmask;
        const bool on = k != 0xFu;
#pragma unroll
        for (int i = 0; i < KP; i++) kc[i] = kc1[i];
#pragma unroll
        for (int qp = 0; qp < NQ; qp++) fq[qp] = f1[qp];
        // rotate the pipeline
        cell1 = cell2;
#pragma unroll
        for (int wd = 0; wd < NW; wd++) w1[wd] = w2[wd];
        if (r + 2 * SPLIT < nrows) {
            const uint32_t* hp = hbase + ((size_t)(r + 2 * SPLIT) << 5);
            cell2 = __ldg(hp);
#pragma unroll
            for (int wd = 0; wd < NW; wd++) w2[wd] = __ldg(hp + (1 + wd) * a.hstride);
        } else {
            cell2 = 0u;
#pragma unroll
            for (int wd = 0; wd < NW; wd++) w2[wd] = PAD;
        }
        load_coef<DIM, KP, NQ>(a.coef + (size_t)cell1 * FEMB_COEF_STRIDE, a.has_rhs != 0, kc1, f1);
        constexpr bool HOIST = FAST && SPLIT == 1;
        unsigned ad[ND];
        if (FAST) {
#pragma unroll
            for (int j = 0; j < ND; j++) ad[j] = sacc + ((words[j >> 2] >> (8 * (j & 3))) & 0xFFu) * (ncol * 8u);
        }
        if (HOIST && on) {
#pragma unroll
            for (int j = 0; j < ND; j++) v[j] = lds64_unless(ad[j], (first >> j) & 1u);
        } else {
#pragma unroll
            for (int j = 0; j < ND; j++) v[j] = 0.0;
        }
        if (on) {
            double Km[DIM][DIM];  // symmetric Kmat from the coefficient order (00, 11, [22,] 01, [02, 12])
            if (DIM == 2) {
                Km[0][0] = kc[0]; Km[1][1] = kc[1]; Km[0][1] = Km[1][0] = kc[2];
            } else {
                Km[0][0] = kc[0]; Km[1][1] = kc[1]; Km[2][2] = kc[2];
                Km[0][1] = Km[1][0] = kc[3]; Km[0][2] = Km[2][0] = kc[4]; Km[1][2] = Km[2][1] = kc[5];
            }
#if FEMB_H1_KTAB_SMEM
            double g[KT_N];
            {
                const unsigned ka = stab + k * (KT_STRIDE * 8u);
#pragma unroll
                for (int i = 0; i < KT_N; i += 2) lds128(ka + i * 8u, g[i], g[i + 1]);
            }
#endif
            double h[NQ][DIM];
#pragma unroll
            for (int qp = 0; qp < NQ; qp++) {
                double dk[DIM];
#pragma unroll
#if FEMB_H1_KTAB_SMEM
                for (int b = 0; b < DIM; b++) dk[b] = g[qp * DIM + b];
#else
                for (int b = 0; b < DIM; b++) dk[b] = tab.wD[qp][b][k];  // per-lane k: indexed constant-bank load
#endif
#pragma unroll
                for (int aa = 0; aa < DIM; aa++) {
                    double s2 = 0.0;
#pragma unroll
                    for (int b = 0; b < DIM; b++) s2 += Km[aa][b] * dk[b];
                    h[qp][aa] = s2;
                }
            }
            if (P2S) {
                // P2 row side in barycentric form (verified against the evaluator's table at launch): with L = 4 lambda(x_q),
                // s_q = sum_a h[q][a], H[a] = sum_q h[q][a]:  vertex 0: -sum_q L_0 s_q + sum_a H[a];  vertex i: sum_q L_i
                // h[q][i-1] - H[i-1];  edge (0,b): sum_q L_0 h[q][b-1] - L_b s_q;  edge (a,b): sum_q L_a h[q][b-1] + L_b h[q][a-1]
                // — 16 uniform table values instead of 120, and 0.6 of the FMAs of the dense product
                double sq[NQ], H[DIM];
#pragma unroll
                for (int aa = 0; aa < DIM; aa++) H[aa] = 0.0;
#pragma unroll
                for (int qp = 0; qp < NQ; qp++) {
                    double t = h[qp][0];
#pragma unroll
                    for (int aa = 1; aa < DIM; aa++) t += h[qp][aa];
                    sq[qp] = t;
#pragma unroll
                    for (int aa = 0; aa < DIM; aa++) H[aa] += h[qp][aa];
                }
                {
                    double t = v[0];
#pragma unroll
                    for (int aa = 0; aa < DIM; aa++) t += H[aa];
#pragma unroll
                    for (int qp = 0; qp < NQ; qp++) t -= tab.L4[qp][0] * sq[qp];
                    v[0] = t;
                }
#pragma unroll
                for (int i = 1; i <= DIM; i++) {
                    double t = v[i] - H[i - 1];
#pragma unroll
                    for (int qp = 0; qp < NQ; qp++) t += tab.L4[qp][i] * h[qp][i - 1];
                    v[i] = t;
                }
#pragma unroll
                for (int e = 0; e < ND - DIM - 1; e++) {
                    const int ea = p2_edge_a<DIM>(e), eb = p2_edge_b<DIM>(e);
                    double t = v[DIM + 1 + e];
#pragma unroll
                    for (int qp = 0; qp < NQ; qp++) {
                        if (ea == 0) {
                            t += tab.L4[qp][0] * h[qp][eb - 1];
                            t -= tab.L4[qp][eb] * sq[qp];
                        } else {
                            t += tab.L4[qp][ea] * h[qp][eb - 1];
                            t += tab.L4[qp][eb] * h[qp][ea - 1];
                        }
                    }
                    v[DIM + 1 + e] = t;
                }
            } else {
#pragma unroll
                for (int j = 0; j < ND; j++) {
                    double vj = v[j];
#pragma unroll
                    for (int qp = 0; qp < NQ; qp++)
#pragma unroll
                        for (int aa = 0; aa < DIM; aa++) vj += tab.D[qp][aa][j] * h[qp][aa];
                    v[j] = vj;
                }
            }
            if (a.has_rhs) {
                double t = 0.0;
#pragma unroll
#if FEMB_H1_KTAB_SMEM
                for (int qp = 0; qp < NQ; qp++) t += g[NQ * DIM + qp] * fq[qp];
#else
                for (int qp = 0; qp < NQ; qp++) t += tab.wphi[qp][k] * fq[qp];
#endif
                bsum += t;
            }
        }
        if (HOIST) {
            if (on) {
#pragma unroll
                for (int j = 0; j < ND; j++) sts64(ad[j], v[j]);
            }
        } else {
#pragma unroll 1
            for (int ph = 0; ph < SPLIT; ph++) {
                if (FAST && on && (SPLIT == 1 || (int)warp == ph)) {
                    double o[ND];
#pragma unroll
                    for (int j = 0; j < ND; j++) o[j] = lds64_unless(ad[j], (first >> j) & 1u);
#pragma unroll
                    for (int j = 0; j < ND; j++) sts64(ad[j], o[j] + v[j]);
                }
                if (!FAST && on && (SPLIT == 1 || (int)warp == ph)) {
#pragma unroll
                    for (int j = 0; j < ND; j++) {
                        const unsigned pos = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                        if (fabs(v[j]) > a.drop_tol) {
                            if (pos == 0xFFu) {
                                err = 1u;
                            } else {
                                const unsigned adr = sacc + pos * (ncol * 8u);
                                sts64(adr, lds64(adr) + v[j]);
                                if (a.touched) a.touched[base + pos] = 1;
                            }
                        }
                    }
                }
                if (SPLIT > 1) __syncthreads();
            }
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
    {
        // write-out: a lane owns the contiguous entries of its column, so it stores them as whole 32-byte sectors
        // (st.global.v4.f64) between a scalar head and tail — a quarter of the store wavefronts of 8-byte stores, and
        // no partially written sectors for L2 to merge.
        const int w0 = (SPLIT == 1 ? 0 : (int)warp);
        const int head = min(len, (int)((4 - (base & 3)) & 3));
        const int ng = (len - head) >> 2;
        const int tail0 = head + 4 * ng;
        for (int i = w0; i < head; i += SPLIT) {
            const double t = lds64(sacc + i * (ncol * 8u));
            out[i] = a.acc_A ? out[i] + t : t;
        }
        for (int i = tail0 + w0; i < len; i += SPLIT) {
            const double t = lds64(sacc + i * (ncol * 8u));
            out[i] = a.acc_A ? out[i] + t : t;
        }
        for (int g = w0; g < ng; g += SPLIT) {
            const int i = head + 4 * g;
            double t[4];
#pragma unroll
            for (int u = 0; u < 4; u++) t[u] = lds64(sacc + (i + u) * (ncol * 8u));
            if (a.acc_A) {
                double o0, o1, o2, o3;
                asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o0), "=d"(o1), "=d"(o2), "=d"(o3) : "l"(out + i));
                t[0] += o0; t[1] += o1; t[2] += o2; t[3] += o3;
            }
            stg256(out + i, t[0], t[1], t[2], t[3]);
        }
    }
    if (a.has_rhs && (SPLIT == 1 || warp == 0)) {
        if (a.acc_b) a.bvec[c] += bsum;
        else a.bvec[c] = bsum;
    }
}

template <int DIM, int ND, int NQ>
static int launch_h1_gather_t(femb_ctx* ctx, const FembHMap& H, const FembEval& EG, const FembEval* EI, HArgs a) {
    const FembSpace& s = ctx->spaces[EG.space];
    const int nc = s.ncomp;  // component-blocked vector spaces: the scalar basis is component 0 of the first ND functions
    HTab<DIM, ND, NQ> tab;
    for (int q = 0; q < NQ; q++) {
        for (int j = 0; j < ND; j++) {
            for (int d = 0; d < DIM; d++) {
                tab.D[q][d][j] = EG.h_refderivs[((size_t)q * s.nd_all + j) * nc * DIM + d];
                tab.wD[q][d][j] = EG.w[q] * tab.D[q][d][j];
            }
            tab.wphi[q][j] = EI ? EG.w[q] * EI->h_refvals[(size_t)q * s.nd_all + j] : 0.0;
        }
    }
    // P2 in barycentric form: L4 from the quadrature points, accepted only if it reproduces the evaluator's table
    bool p2s = ND == (DIM + 1) * (DIM + 2) / 2;
    if (p2s) {
        for (int q = 0; q < NQ; q++) {
            double l0 = 1.0;
            for (int d = 0; d < DIM; d++) {
                tab.L4[q][d + 1] = 4.0 * EG.xref[q * DIM + d];
                l0 -= EG.xref[q * DIM + d];
            }
            tab.L4[q][0] = 4.0 * l0;
        }
        for (int q = 0; q < NQ && p2s; q++)
            for (int d = 0; d < DIM && p2s; d++)
                for (int j = 0; j < ND && p2s; j++) {
                    double ref;  // d phi_j / d xref_d at x_q from the barycentric formulas
                    if (j == 0) ref = -(tab.L4[q][0] - 1.0);
                    else if (j <= DIM) ref = (j - 1 == d) ? tab.L4[q][j] - 1.0 : 0.0;
                    else {
                        const int ea = DIM == 2 ? p2_edge_a<2>(j - DIM - 1) : p2_edge_a<3>(j - DIM - 1), eb = DIM == 2 ? p2_edge_b<2>(j - DIM - 1) : p2_edge_b<3>(j - DIM - 1);
                        if (ea == 0) ref = ((eb - 1 == d) ? tab.L4[q][0] : 0.0) - tab.L4[q][eb];
                        else ref = ((eb - 1 == d) ? tab.L4[q][ea] : 0.0) + ((ea - 1 == d) ? tab.L4[q][eb] : 0.0);
                    }
                    if (fabs(ref - tab.D[q][d][j]) > 1e-13 * (1.0 + fabs(ref))) p2s = false;
                }
    }
    if (!p2s)
        for (int q = 0; q < NQ; q++)
            for (int i = 0; i <= DIM; i++) tab.L4[q][i] = 0.0;
    const bool fast = H.complete && !a.touched && !(a.drop_tol > 0.0);
    auto kern1 = fast ? (p2s ? k_h1_gather<DIM, ND, NQ, 1, true, true> : k_h1_gather<DIM, ND, NQ, 1, true, false>) : k_h1_gather<DIM, ND, NQ, 1, false, false>;
    auto kern4 = fast ? (p2s ? k_h1_gather<DIM, ND, NQ, 4, true, true> : k_h1_gather<DIM, ND, NQ, 4, true, false>) : k_h1_gather<DIM, ND, NQ, 4, false, false>;
    size_t max1 = 0, max4 = 0;
    for (auto& r : H.ranges) {
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
    for (auto& r : H.ranges) {
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
    {
        // P2 on tetrahedra without value filter / slot tracking: closed-form kernel (femb_p2.cu)
        const int st = launch_p2_gather(ctx, eg, ei, mu, rhs, drop_tol);
        if (st != FEMB_ERR_UNSUPPORTED) return st;
    }
    FEMB_TRY(femb_hmap_set_canon(ctx, ctx->hm, false));
    const FembHMap& H = ctx->hm;
    if (!H.map || H.space != EG.space || EG.nqp > 10 || s.ncomp != 1) return FEMB_ERR_UNSUPPORTED;
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
    a.coef = ctx->cellcoef; a.hslice_ptr = H.slice_ptr; a.hslice_zero = H.slice_zero; a.hmap = H.map; a.hstride = H.stride; a.colptr = ctx->colptr; a.sub0 = a.sublen = nullptr; a.ndofs = s.ndofs;
    a.nzval = ctx->nzval; a.touched = ctx->touched; a.bvec = ctx->bvec; a.errflag = ctx->errflag; a.drop_tol = drop_tol;
    FEMB_TRY(femb_settle_partial_zero(ctx));
    a.acc_A = ctx->zero_pending_A ? 0 : 1;
    a.acc_b = ctx->zero_pending_b ? 0 : 1;
    a.has_rhs = rhs.kind != FEMB_RHS_NONE;
    a.slice_begin = a.slice_end = 0;
    ctx->zero_pending_A = false;
    if (a.has_rhs) ctx->zero_pending_b = false;
    if (key == 20603) return launch_h1_gather_t<2, 6, 3>(ctx, H, EG, EI, a);
    return launch_h1_gather_t<3, 10, 4>(ctx, H, EG, EI, a);
}

// mu (grad u, grad v) of a component-blocked vector-valued P2 space in block (0,0) of a multi-space layout (the Laplacian of
// Example210:284-290,368): every component is the scalar P2 stiffness, the cross-component entries are exact zeros.  Owner-
// computes by column like the scalar case (write-once per launch, bit-reproducible); the columns also hold rows of other
// blocks, so the result is ADDED to nzval (the lazy zero is flushed first).  FEMB_ERR_UNSUPPORTED when not covered.
int launch_h1_gather_vec(femb_ctx* ctx, int eg, double mu, int ei, const RhsDev* rhs, bool* rhs_done) {
    if (rhs_done) *rhs_done = false;
    const FembEval& EG = ctx->evals[eg];
    const FembSpace& s = ctx->spaces[EG.space];
    if (s.ncomp < 2 || ctx->touched || EG.nqp > 10) return FEMB_ERR_UNSUPPORTED;
    {
        // P2 on triangles: closed-form kernel (femb_p2.cu)
        const int st2 = launch_p2_gather_vec(ctx, eg, mu, ei, rhs, rhs_done);
        if (st2 != FEMB_ERR_UNSUPPORTED) return st2;
    }
    if (EG.h_refderivs.size() != (size_t)EG.nqp * s.nd_all * s.ncomp * ctx->dim) return FEMB_ERR_UNSUPPORTED;
    const int nds = s.nd / s.ncomp;
    const int key = ctx->dim * 10000 + nds * 100 + EG.nqp;
    if (key != 20603 && key != 20609 && key != 31004) return FEMB_ERR_UNSUPPORTED;
    int st = femb_build_hmap_vec(ctx);
    if (st != FEMB_OK) return st;
    FEMB_TRY(femb_hmap_set_canon(ctx, ctx->hm_vec, false));
    const FembHMap& H = ctx->hm_vec;
    if (H.space != EG.space) return FEMB_ERR_UNSUPPORTED;
    if (!ctx->cellcoef) FEMB_TRY(femb_alloc(ctx, &ctx->cellcoef, (size_t)ctx->ncells * FEMB_COEF_STRIDE));
    if (ctx->ncells == 0) return FEMB_OK;
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    RhsDev norhs;
    norhs.kind = FEMB_RHS_NONE;
    norhs.ncomp = 1;
    norhs.fvals = nullptr;
    QuadDev q = quad_dev(EG, ctx->dim);
    const unsigned gridc = (unsigned)((ctx->ncells + 127) / 128);
    if (ctx->dim == 2) k_cell_coef<2><<<gridc, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, mu, norhs, q, ctx->cellcoef);
    else k_cell_coef<3><<<gridc, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, mu, norhs, q, ctx->cellcoef);
    KERNEL_CHECK(ctx);
    HArgs a;
    a.coef = ctx->cellcoef; a.hslice_ptr = H.slice_ptr; a.hslice_zero = H.slice_zero; a.hmap = H.map; a.hstride = H.stride;
    a.colptr = ctx->colptr + H.col_off; a.sub0 = H.sub0; a.sublen = H.sublen; a.ndofs = s.ndofs;
    a.nzval = ctx->nzval; a.touched = nullptr; a.bvec = nullptr; a.errflag = ctx->errflag; a.drop_tol = 0.0;
    a.acc_A = 1;
    a.acc_b = 1;
    a.has_rhs = 0;
    a.slice_begin = a.slice_end = 0;
    if (key == 20603) return launch_h1_gather_t<2, 6, 3>(ctx, H, EG, nullptr, a);
    if (key == 20609) return launch_h1_gather_t<2, 6, 9>(ctx, H, EG, nullptr, a);
    return launch_h1_gather_t<3, 10, 4>(ctx, H, EG, nullptr, a);
}
