// Numeric phase, generic part: timers / lazy zeroing, batched update_basis! probes (stages 1-2), the generic
// quadrature kernels for any pair of evaluators (stage 3-4: k_bilinear, k_bilinear_rows) and the right-hand-side gathers
// (stage 5), plus their C-ABI entry points.
//   stage 1  affine map / determinant          (L2GTransformer; feevaluator.jl:344-363)
//   stage 2  push-forward / Piola               (feevaluator_h1.jl, feevaluator_hdiv.jl)
//   stage 3  quadrature-weighted contraction    (Example200:139-146, Example210:284-301,322-355)
//   stage 4  scatter into the CSC image         (Example200:149-161, Example210:368-382)
//   stage 5  right-hand side                    (Example200:165-178, Example210:304-318,358-363)
#include "femb_kernels.cuh"

// result lengths the operators of this path produce (Length4Operator): compile-time R keeps per-thread result arrays in registers
#define FEMB_DISPATCH_R(R, CALL)                         \
    switch (R) {                                         \
        case 1: CALL(1); break;                          \
        case 2: CALL(2); break;                          \
        case 3: CALL(3); break;                          \
        case 4: CALL(4); break;                          \
        case 6: CALL(6); break;                          \
        case 9: CALL(9); break;                          \
        default: return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "operator result length %d is not instantiated", (int)(R)); \
    }

// ---------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------
int femb_timer_begin(femb_ctx* ctx) {
    ctx->last_launches = 0;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return FEMB_OK;
}
int femb_timer_end(femb_ctx* ctx) {
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return FEMB_OK;
}

SpaceDev space_dev(const FembSpace& s) {
    SpaceDev d;
    d.family = s.family; d.ncomp = s.ncomp; d.nd = s.nd; d.nd_all = s.nd_all; d.handler = s.handler; d.nsign = s.nsign;
    d.celldofs = s.celldofs; d.signs = s.signs; d.orient = s.orient;
    return d;
}

// ---- lazy zeroing (femb_matrix_zero / femb_vector_zero) -------------------------------------------------
// The rows of a layout block form one contiguous run in every column of the block (rowval ascending): [lo, hi) by two binary searches.
__device__ __forceinline__ void block_run(const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int64_t col, int64_t r0, int64_t r1, int64_t& lo,
                                          int64_t& hi) {
    const int64_t b = colptr[col] - 1, e = colptr[col + 1] - 1;
    int64_t l = b, h = e;
    while (l < h) {
        const int64_t m = (l + h) >> 1;
        if (rowval[m] <= r0) l = m + 1;  // 1-based rows of the block: r0 + 1 .. r1
        else h = m;
    }
    lo = l;
    h = e;
    while (l < h) {
        const int64_t m = (l + h) >> 1;
        if (rowval[m] <= r1) l = m + 1;
        else h = m;
    }
    hi = l;
}

// one warp per column of the block: zero the run
__global__ void k_zero_block(const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int64_t c0, int64_t c1, int64_t r0, int64_t r1,
                             double* __restrict__ nzval) {
    const int64_t col = c0 + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (col >= c1) return;
    int64_t lo, hi;
    block_run(colptr, rowval, col, r0, r1, lo, hi);
    for (int64_t p = lo + (threadIdx.x & 31); p < hi; p += 32) nzval[p] = 0.0;
}

__global__ void k_count_block(const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int64_t c0, int64_t c1, int64_t r0, int64_t r1,
                              unsigned long long* __restrict__ count) {
    const int64_t col = c0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t lo = 0, hi = 0;
    if (col < c1) block_run(colptr, rowval, col, r0, r1, lo, hi);
    unsigned long long n = (unsigned long long)(hi - lo);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, n);
}

// number of pattern entries inside layout block (brow, bcol); cached until the pattern changes
int femb_block_entries(femb_ctx* ctx, int brow, int bcol, int64_t* n) {
    if (brow < 0 || brow >= ctx->nrs || bcol < 0 || bcol >= ctx->ncs || !ctx->colptr) return femb_fail(ctx, FEMB_ERR_ARG, "block (%d,%d): no such block / no pattern", brow, bcol);
    int64_t& cached = ctx->block_nnz[brow][bcol];
    if (cached < 0) {
        const int64_t c0 = ctx->col_off[bcol], c1 = ctx->col_off[bcol + 1];
        unsigned long long* cnt = nullptr;
        unsigned long long h = 0;
        CUDA_TRY(ctx, cudaMalloc(&cnt, sizeof(*cnt)));
        cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(*cnt), ctx->stream);
        if (e == cudaSuccess && c1 > c0) {
            k_count_block<<<(unsigned)((c1 - c0 + 255) / 256), 256, 0, ctx->stream>>>(ctx->colptr, ctx->rowval, c0, c1, ctx->row_off[brow], ctx->row_off[brow + 1], cnt);
            e = cudaGetLastError();
            ctx->launches++;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(cnt);
        if (e != cudaSuccess) return femb_fail(ctx, FEMB_ERR_CUDA, "femb_block_entries: %s", cudaGetErrorString(e));
        cached = (int64_t)h;
    }
    *n = cached;
    return FEMB_OK;
}

// A kernel that overwrites every pattern entry of block (brow, bcol) takes the block out of the pending zero: true = store, do not add.
bool femb_claim_block(femb_ctx* ctx, int brow, int bcol) {
    static const bool on = [] { const char* e = getenv("FEMB_BLOCK_CLAIMS"); return !(e && e[0] == '0'); }();  // 0: always zero, then add (A/B timing)
    if (!on || !ctx->zero_pending_A) return false;
    const uint64_t bit = 1ull << (brow * FEMB_MAX_SPACES + bcol);
    if (!(ctx->zero_mask & bit)) return false;
    ctx->zero_mask &= ~bit;
    return true;
}

// apply the lazy zeroing for kernels that accumulate: the whole array, or — after claims — the runs of the blocks nobody has overwritten
int femb_flush_zero(femb_ctx* ctx, bool mat, bool vec) {
    if (mat && ctx->zero_pending_A) {
        if (ctx->nzval && ctx->zero_mask == ~0ull) {
            CUDA_TRY(ctx, cudaMemsetAsync(ctx->nzval, 0, sizeof(double) * ctx->nnz, ctx->stream));
        } else if (ctx->nzval) {
            for (int br = 0; br < ctx->nrs; br++)
                for (int bc = 0; bc < ctx->ncs; bc++) {
                    const int64_t c0 = ctx->col_off[bc], c1 = ctx->col_off[bc + 1];
                    if (!(ctx->zero_mask & (1ull << (br * FEMB_MAX_SPACES + bc))) || c1 <= c0 || ctx->row_off[br + 1] <= ctx->row_off[br]) continue;
                    int64_t n = 0;
                    FEMB_TRY(femb_block_entries(ctx, br, bc, &n));
                    if (n == 0) continue;
                    k_zero_block<<<(unsigned)((c1 - c0 + 7) / 8), 256, 0, ctx->stream>>>(ctx->colptr, ctx->rowval, c0, c1, ctx->row_off[br], ctx->row_off[br + 1], ctx->nzval);
                    KERNEL_CHECK(ctx);
                }
        }
        ctx->zero_pending_A = false;
        ctx->zero_mask = ~0ull;
    }
    if (vec && ctx->zero_pending_b) {
        if (ctx->bvec) CUDA_TRY(ctx, cudaMemsetAsync(ctx->bvec, 0, sizeof(double) * ctx->nrows, ctx->stream));
        ctx->zero_pending_b = false;
    }
    return FEMB_OK;
}

// the single-kernel paths read zero_pending_A as "every entry is still to be zeroed": settle the blocks left over by earlier claims first
int femb_settle_partial_zero(femb_ctx* ctx) {
    if (ctx->zero_pending_A && ctx->zero_mask != ~0ull) return femb_flush_zero(ctx, true, false);
    return FEMB_OK;
}

int make_rhs(femb_ctx* ctx, int kind, const double* p, int ncomp, int nqp, RhsDev& r, double** dev_fvals) {
    r.kind = kind;
    r.ncomp = ncomp;
    r.fvals = nullptr;
    *dev_fvals = nullptr;
    for (double& v : r.p) v = 0.0;
    if (kind == FEMB_RHS_NONE) return FEMB_OK;
    if (!p) return femb_fail(ctx, FEMB_ERR_ARG, "rhs data pointer is NULL");
    if (kind == FEMB_RHS_AFFINE) {
        int n = ncomp * (ctx->dim + 1);
        if (n > 16) return femb_fail(ctx, FEMB_ERR_ARG, "too many affine rhs coefficients");
        for (int i = 0; i < n; i++) r.p[i] = p[i];
    } else if (kind == FEMB_RHS_EX210) {
        if (ctx->dim != 2 || ncomp != 2) return femb_fail(ctx, FEMB_ERR_ARG, "FEMB_RHS_EX210 is the 2-D two-component f! of Example210");
        r.p[0] = p[0];
        r.p[1] = p[1];
    } else if (kind == FEMB_RHS_QPVALUES) {
        size_t n = (size_t)ctx->ncells * nqp * ncomp;
        FEMB_TRY(femb_alloc(ctx, dev_fvals, n));
        FEMB_TRY(femb_h2d(ctx, *dev_fvals, p, n));
        r.fvals = *dev_fvals;
    } else {
        return femb_fail(ctx, FEMB_ERR_ARG, "unknown rhs kind %d", kind);
    }
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// batched update_basis! and quadrature-point coordinates (parity probes for stages 1-2)
// ---------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void k_eval_basis(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol,
                             SpaceDev s, int op, int R, int nqp, const double* __restrict__ rv, const double* __restrict__ rd, int64_t c0, int64_t c1,
                             double* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t per = (int64_t)nqp * s.nd;
    if (t >= (c1 - c0) * per) return;
    int64_t cl = t / per;
    int rem = (int)(t - cl * per);
    int qp = rem / s.nd, dof = rem - qp * s.nd;
    Geo<DIM> g;
    load_geo<DIM>(coords, cellnodes, cellvol, c0 + cl, g);
    double v[9];
    for (int r = 0; r < R; r++) v[r] = 0.0;
    eval_op<DIM>(s, op, g, c0 + cl, dof, qp, rv, rd, v);
    for (int r = 0; r < R; r++) out[t * R + r] = v[r];
}


template <int DIM>
__global__ void k_qp_coords(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, int64_t ncells, QuadDev q, double* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ncells * q.nqp) return;
    int64_t cell = t / q.nqp;
    int qp = (int)(t - cell * q.nqp);
    Geo<DIM> g;
    load_geo<DIM>(coords, cellnodes, nullptr, cell, g);
    double x[DIM];
    eval_trafo<DIM>(g, q.xref + qp * DIM, x);
    for (int d = 0; d < DIM; d++) out[t * DIM + d] = x[d];
}

QuadDev quad_dev(const FembEval& e, int dim) {
    QuadDev q;
    q.nqp = e.nqp;
    for (int i = 0; i < e.nqp * dim; i++) q.xref[i] = e.xref[i];
    for (int i = 0; i < e.nqp; i++) q.w[i] = e.w[i];
    return q;
}

extern "C" int32_t femb_evaluator_eval(femb_ctx* ctx, int32_t id, int64_t c0, int64_t c1, double* out) {
    FEMB_CHECK_CTX(ctx);
    if (id < 0 || id >= FEMB_MAX_EVALS || !ctx->evals[id].set) return femb_fail(ctx, FEMB_ERR_ARG, "unknown evaluator");
    if (c0 < 0 || c1 > ctx->ncells || c0 > c1 || !out) return femb_fail(ctx, FEMB_ERR_ARG, "bad cell range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const FembEval& e = ctx->evals[id];
    const FembSpace& s = ctx->spaces[e.space];
    int64_t n = (c1 - c0) * e.nqp * s.nd;
    if (n == 0) return FEMB_OK;
    double* dout = nullptr;
    FEMB_TRY(femb_alloc(ctx, &dout, (size_t)n * e.resultdim));
    unsigned grid = (unsigned)((n + 127) / 128);
    if (ctx->dim == 2)
        k_eval_basis<2><<<grid, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, space_dev(s), e.op, e.resultdim, e.nqp, e.refvals, e.refderivs, c0, c1, dout);
    else
        k_eval_basis<3><<<grid, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, space_dev(s), e.op, e.resultdim, e.nqp, e.refvals, e.refderivs, c0, c1, dout);
    ctx->launches++;
    int st = femb_d2h(ctx, out, dout, (size_t)n * e.resultdim);
    if (st == FEMB_OK) st = femb_sync(ctx);
    cudaFree(dout);
    return st;
}

extern "C" int32_t femb_qp_coords(femb_ctx* ctx, double* out) {
    FEMB_CHECK_CTX(ctx);
    if (!out || !ctx->nqp) return femb_fail(ctx, FEMB_ERR_ARG, "femb_qp_coords: quadrature not set");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    QuadDev q;
    q.nqp = ctx->nqp;
    for (int i = 0; i < ctx->nqp * ctx->dim; i++) q.xref[i] = ctx->xref[i];
    int64_t n = ctx->ncells * ctx->nqp;
    double* dout = nullptr;
    FEMB_TRY(femb_alloc(ctx, &dout, (size_t)n * ctx->dim));
    unsigned grid = (unsigned)((n + 127) / 128);
    if (ctx->dim == 2) k_qp_coords<2><<<grid, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->ncells, q, dout);
    else k_qp_coords<3><<<grid, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->ncells, q, dout);
    ctx->launches++;
    int st = femb_d2h(ctx, out, dout, (size_t)n * ctx->dim);
    if (st == FEMB_OK) st = femb_sync(ctx);
    cudaFree(dout);
    return st;
}

// ---------------------------------------------------------------------------------------------------
// compute_error (Example210_LowLevelNavierStokes.jl:110-154): error[d, cell] = sum_qp (uh_d(x_qp) - u_d(x_qp))^p |T| w_qp with
// uh = eval_febe! of the coefficient vector through any registered evaluator (Identity: L^p error, Gradient: its seminorm)
// ---------------------------------------------------------------------------------------------------
template <int DIM, int R>  // R: result length of the operator at compile time (static indices: no local memory)
__global__ void k_eval_error(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol, int64_t ncells,
                             SpaceDev s, int op, QuadDev q, const double* __restrict__ rv, const double* __restrict__ rd,
                             const double* __restrict__ coef, const double* __restrict__ exact, int p, double* __restrict__ out) {
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cell >= ncells) return;
    Geo<DIM> g;
    load_geo<DIM>(coords, cellnodes, cellvol, cell, g);
    double err[R];
    _Pragma("unroll") for (int r = 0; r < R; r++) err[r] = 0.0;
    for (int qp = 0; qp < q.nqp; qp++) {
        double uh[R];
        _Pragma("unroll") for (int r = 0; r < R; r++) uh[r] = 0.0;
        for (int dof = 0; dof < s.nd; dof++) {
            double v[R];
            _Pragma("unroll") for (int r = 0; r < R; r++) v[r] = 0.0;
            eval_op<DIM, R>(s, op, g, cell, dof, qp, rv, rd, v);
            const double c = coef[s.celldofs[cell * s.nd + dof] - 1];
            _Pragma("unroll") for (int r = 0; r < R; r++) uh[r] += c * v[r];
        }
        _Pragma("unroll") for (int r = 0; r < R; r++) {
            const double diff = uh[r] - (exact ? exact[((size_t)cell * q.nqp + qp) * R + r] : 0.0);
            double pw = 1.0;
            for (int i = 0; i < p; i++) pw *= diff;
            err[r] += pw * g.vol * q.w[qp];
        }
    }
    _Pragma("unroll") for (int r = 0; r < R; r++) out[cell * R + r] = err[r];
}

extern "C" int32_t femb_eval_error(femb_ctx* ctx, int32_t id, const double* coefficients, const double* exact_qp, int32_t p, double* error_out) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(check_eval(ctx, id));
    const FembEval& e = ctx->evals[id];
    const FembSpace& s = ctx->spaces[e.space];
    if (!coefficients || !error_out || p < 1 || p > 16 || e.resultdim > 9) return femb_fail(ctx, FEMB_ERR_ARG, "femb_eval_error: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->ncells == 0) return FEMB_OK;
    const size_t nq = (size_t)ctx->ncells * e.nqp * e.resultdim;
    double *dcoef = nullptr, *dexact = nullptr, *dout = nullptr;
    int st = femb_alloc(ctx, &dcoef, (size_t)s.ndofs);
    if (st == FEMB_OK) st = femb_alloc(ctx, &dout, (size_t)ctx->ncells * e.resultdim);
    if (st == FEMB_OK && exact_qp) st = femb_alloc(ctx, &dexact, nq);
    if (st == FEMB_OK) st = femb_h2d(ctx, dcoef, coefficients, (size_t)s.ndofs);
    if (st == FEMB_OK && exact_qp) st = femb_h2d(ctx, dexact, exact_qp, nq);
    if (st == FEMB_OK) {
        const unsigned grid = (unsigned)((ctx->ncells + 127) / 128);
        const QuadDev q = quad_dev(e, ctx->dim);
#define FEMB_EE(RR)                                                                                                                                         \
    do {                                                                                                                                                \
        if (ctx->dim == 2)                                                                                                                              \
            k_eval_error<2, RR><<<grid, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, space_dev(s), e.op, q, e.refvals, \
                                                               e.refderivs, dcoef, dexact, p, dout);                                                    \
        else                                                                                                                                            \
            k_eval_error<3, RR><<<grid, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, space_dev(s), e.op, q, e.refvals, \
                                                               e.refderivs, dcoef, dexact, p, dout);                                                    \
    } while (0)
        switch (e.resultdim) {
            case 1: FEMB_EE(1); break;
            case 2: FEMB_EE(2); break;
            case 3: FEMB_EE(3); break;
            case 4: FEMB_EE(4); break;
            case 6: FEMB_EE(6); break;
            case 9: FEMB_EE(9); break;
            default: st = femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "operator result length %d is not instantiated", e.resultdim);
        }
#undef FEMB_EE
        ctx->launches++;
        st = femb_d2h(ctx, error_out, dout, (size_t)ctx->ncells * e.resultdim);
        if (st == FEMB_OK) st = femb_sync(ctx);
    }
    femb_free(&dcoef);
    femb_free(&dexact);
    femb_free(&dout);
    return st;
}

// ---------------------------------------------------------------------------------------------------
// generic cell-parallel bilinear form: CTA = CPB cells; cvals staged in shared memory
// ---------------------------------------------------------------------------------------------------
struct BilArgs {
    const double* coords;
    const int32_t* cellnodes;
    const double* cellvol;
    int64_t ncells;
    SpaceDev rs, cs;
    int op_r, op_c, R, nqp, same;
    const double *rv_r, *rd_r, *rv_c, *rd_c;
    double w[FEMB_MAX_QP];
    const int32_t* slots;    // [cell][ndr][ndc]
    const int32_t* slots_t;  // [cell][ndc][ndr] (transpose block) or null
    double* nzval;
    uint8_t* touched;
    int32_t* errflag;
    double factor, drop_tol;
    int flags, cpb;
    const int32_t* perm;  // cell permutation (colouring) or null
    int64_t p0, p1;       // range in perm / cells
    int atomic;
};


template <int DIM, int R>
__global__ void __launch_bounds__(128) k_bilinear(BilArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Geo<DIM>* geo = reinterpret_cast<Geo<DIM>*>(smem_raw);
    double* rowv = reinterpret_cast<double*>(geo + a.cpb);
    const int ndr = a.rs.nd, ndc = a.cs.nd;
    const int per_r = a.nqp * R * ndr, per_c = a.nqp * R * ndc;
    double* colv = a.same ? rowv : rowv + (size_t)a.cpb * per_r;
    const int64_t base = a.p0 + (int64_t)blockIdx.x * a.cpb;
    const int ncl = (int)min((int64_t)a.cpb, a.p1 - base);
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int cl = tid; cl < ncl; cl += nt) {
        int64_t cell = a.perm ? a.perm[base + cl] : base + cl;
        load_geo<DIM>(a.coords, a.cellnodes, a.cellvol, cell, geo[cl]);
    }
    __syncthreads();
    for (int i = tid; i < ncl * a.nqp * ndr; i += nt) {
        int cl = i / (a.nqp * ndr), rem = i - cl * (a.nqp * ndr);
        int qp = rem / ndr, dof = rem - qp * ndr;
        int64_t cell = a.perm ? a.perm[base + cl] : base + cl;
        double v[R];
        eval_op<DIM, R>(a.rs, a.op_r, geo[cl], cell, dof, qp, a.rv_r, a.rd_r, v);
#pragma unroll
        for (int r = 0; r < R; r++) rowv[(size_t)cl * per_r + (qp * R + r) * ndr + dof] = v[r];
    }
    if (!a.same) {
        for (int i = tid; i < ncl * a.nqp * ndc; i += nt) {
            int cl = i / (a.nqp * ndc), rem = i - cl * (a.nqp * ndc);
            int qp = rem / ndc, dof = rem - qp * ndc;
            int64_t cell = a.perm ? a.perm[base + cl] : base + cl;
            double v[R];
            eval_op<DIM, R>(a.cs, a.op_c, geo[cl], cell, dof, qp, a.rv_c, a.rd_c, v);
#pragma unroll
            for (int r = 0; r < R; r++) colv[(size_t)cl * per_c + (qp * R + r) * ndc + dof] = v[r];
        }
    }
    __syncthreads();
    const int per = ndr * ndc;
    const bool sym = a.flags & FEMB_BIL_SYMMETRIC_UPPER;
    for (int i = tid; i < ncl * per; i += nt) {
        int cl = i / per, jk = i - cl * per;
        int j = jk / ndc, k = jk - j * ndc;
        if (sym && j > k) continue;
        int64_t cell = a.perm ? a.perm[base + cl] : base + cl;
        const double* rp = rowv + (size_t)cl * per_r + j;
        const double* cp = colv + (size_t)cl * per_c + k;
        double temp = 0.0;
        for (int qp = 0; qp < a.nqp; qp++) {
            double d = 0.0;
            for (int r = 0; r < R; r++) d += rp[(qp * R + r) * ndr] * cp[(qp * R + r) * ndc];
            temp += a.w[qp] * d;
        }
        const double vol = (a.flags & FEMB_BIL_NO_VOLUME) ? 1.0 : geo[cl].vol;
        double val = sym ? temp * (a.factor * vol) : (a.factor * temp) * vol;
        if (sym && !(fabs(val) > a.drop_tol)) continue;
        scatter_add(a.nzval, a.touched, a.errflag, a.slots[cell * per + jk], val, a.atomic);
        if (sym && k > j) scatter_add(a.nzval, a.touched, a.errflag, a.slots[cell * per + k * ndc + j], val, a.atomic);
        if (a.slots_t) scatter_add(a.nzval, a.touched, a.errflag, a.slots_t[cell * per + k * ndr + j], val, a.atomic);
    }
}

// Row-blocked variant: one thread per (cell, row dof j).  The row-side operator values of (cell, j) stay in registers
// (compile-time NQ x R), the column-side values of the cell are staged once in shared memory; every thread then forms
// its row of the local matrix (one LDS per FMA, no index arithmetic per entry) and scatters it.  Same arithmetic order
// as k_bilinear / the reference loops: temp = sum_qp w_qp (sum_r row * col), value = (factor * temp) * |T|.
#define FEMB_ROWS_PAD 2
template <int DIM, int NQ, int R>
__global__ void __launch_bounds__(128) k_bilinear_rows(BilArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Geo<DIM>* geo = reinterpret_cast<Geo<DIM>*>(smem_raw);
    double* colv = reinterpret_cast<double*>(geo + a.cpb);  // [cl][qp * R + r][k], cells FEMB_ROWS_PAD doubles apart in the banks
    const int ndr = a.rs.nd, ndc = a.cs.nd;
    const int cstride = NQ * R * ndc + FEMB_ROWS_PAD;  // the lanes of different cells read different banks (broadcast per cell)
    const int64_t base = a.p0 + (int64_t)blockIdx.x * a.cpb;
    const int ncl = (int)min((int64_t)a.cpb, a.p1 - base);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int cl = tid; cl < ncl; cl += nt) load_geo<DIM>(a.coords, a.cellnodes, a.cellvol, a.perm ? a.perm[base + cl] : base + cl, geo[cl]);
    __syncthreads();
    const int cl = tid / ndr, j = tid - cl * ndr;
    const bool active = cl < ncl;
    const int64_t cell = active ? (a.perm ? a.perm[base + cl] : base + cl) : 0;
    double rv[NQ * R];
    if (active) {
#pragma unroll
        for (int qp = 0; qp < NQ; qp++) {
            double v[R];
            eval_op<DIM, R>(a.rs, a.op_r, geo[cl], cell, j, qp, a.rv_r, a.rd_r, v);
#pragma unroll
            for (int r = 0; r < R; r++) rv[qp * R + r] = v[r];
        }
        if (a.same) {
#pragma unroll
            for (int i = 0; i < NQ * R; i++) colv[(size_t)cl * cstride + i * ndc + j] = rv[i];
        }
    }
    if (!a.same) {
        for (int i = tid; i < ncl * NQ * ndc; i += nt) {
            const int c2 = i / (NQ * ndc), rem = i - c2 * (NQ * ndc);
            const int qp = rem / ndc, k = rem - qp * ndc;
            double v[R];
            eval_op<DIM, R>(a.cs, a.op_c, geo[c2], a.perm ? a.perm[base + c2] : base + c2, k, qp, a.rv_c, a.rd_c, v);
#pragma unroll
            for (int r = 0; r < R; r++) colv[(size_t)c2 * cstride + (qp * R + r) * ndc + k] = v[r];
        }
    }
    __syncthreads();
    if (!active) return;
    const bool sym = a.flags & FEMB_BIL_SYMMETRIC_UPPER;
    const double vol = (a.flags & FEMB_BIL_NO_VOLUME) ? 1.0 : geo[cl].vol;
    const int per = ndr * ndc;
    const double* cv = colv + (size_t)cl * cstride;
    const int32_t* sl = a.slots + cell * per + j * ndc;
    const int kfirst = sym ? j : 0;
    for (int k0 = kfirst; k0 < ndc; k0 += 4) {
        int32_t sk[4];  // slots of the next four entries in flight together (static indices: registers)
#pragma unroll
        for (int i = 0; i < 4; i++) sk[i] = (k0 + i < ndc) ? __ldg(sl + k0 + i) : 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = k0 + i;
            if (k < ndc) {
                double temp = 0.0;
#pragma unroll
                for (int qp = 0; qp < NQ; qp++) {
                    double d = 0.0;
#pragma unroll
                    for (int r = 0; r < R; r++) d += rv[qp * R + r] * cv[(qp * R + r) * ndc + k];
                    temp += a.w[qp] * d;
                }
                const double val = sym ? temp * (a.factor * vol) : (a.factor * temp) * vol;
                if (!sym || fabs(val) > a.drop_tol) {
                    scatter_add(a.nzval, a.touched, a.errflag, sk[i], val, a.atomic);
                    if (sym && k > j) scatter_add(a.nzval, a.touched, a.errflag, __ldg(a.slots + cell * per + k * ndc + j), val, a.atomic);
                    if (a.slots_t) scatter_add(a.nzval, a.touched, a.errflag, __ldg(a.slots_t + cell * per + k * ndr + j), val, a.atomic);
                }
            }
        }
    }
}

const FembBlock* find_block(const femb_ctx* ctx, int brow, int bcol) {
    for (auto& b : ctx->blocks)
        if (b.brow == brow && b.bcol == bcol) return &b;
    return nullptr;
}

int launch_bilinear(femb_ctx* ctx, int er, int ec, int op_r, int op_c, int brow, int bcol, double factor, int flags, double drop_tol) {
    const FembEval& ER = ctx->evals[er];
    const FembEval& EC = ctx->evals[ec];
    const FembSpace& rs = ctx->spaces[ER.space];
    const FembSpace& cs = ctx->spaces[EC.space];
    if (brow < 0 || brow >= ctx->nrs || bcol < 0 || bcol >= ctx->ncs || ctx->row_spaces[brow] != ER.space || ctx->col_spaces[bcol] != EC.space)
        return femb_fail(ctx, FEMB_ERR_ARG, "evaluator spaces do not match block (%d,%d)", brow, bcol);
    FEMB_TRY(femb_build_slotmaps(ctx));
    const FembBlock* blk = find_block(ctx, brow, bcol);
    if (!blk || !blk->slots) return femb_fail(ctx, FEMB_ERR_ARG, "block (%d,%d) is not part of the pattern", brow, bcol);
    const FembBlock* blk_t = nullptr;
    if (flags & FEMB_BIL_TRANSPOSE_COPY) {
        // transposed block: row space list index of the column space and vice versa
        int tr = -1, tc = -1;
        for (int i = 0; i < ctx->nrs; i++) if (ctx->row_spaces[i] == EC.space) tr = i;
        for (int i = 0; i < ctx->ncs; i++) if (ctx->col_spaces[i] == ER.space) tc = i;
        blk_t = (tr >= 0 && tc >= 0) ? find_block(ctx, tr, tc) : nullptr;
        if (!blk_t) return femb_fail(ctx, FEMB_ERR_ARG, "transposed block of (%d,%d) is not part of the pattern", brow, bcol);
    }
    int Rr = result_dim(op_r, ctx->dim, rs.ncomp), Rc = result_dim(op_c, ctx->dim, cs.ncomp);
    if (Rr != Rc) return femb_fail(ctx, FEMB_ERR_ARG, "operator result lengths differ (%d vs %d)", Rr, Rc);
    if (ER.nqp != EC.nqp) return femb_fail(ctx, FEMB_ERR_ARG, "evaluators use different quadrature rules");
    if ((flags & FEMB_BIL_SYMMETRIC_UPPER) && (er != ec)) return femb_fail(ctx, FEMB_ERR_ARG, "SYMMETRIC_UPPER needs identical evaluators");
    BilArgs a;
    a.coords = ctx->coords; a.cellnodes = ctx->cellnodes; a.cellvol = ctx->cellvol; a.ncells = ctx->ncells;
    a.rs = space_dev(rs); a.cs = space_dev(cs);
    a.op_r = op_r; a.op_c = op_c; a.R = Rr; a.nqp = ER.nqp; a.same = (er == ec && op_r == op_c);
    a.rv_r = ER.refvals; a.rd_r = ER.refderivs; a.rv_c = EC.refvals; a.rd_c = EC.refderivs;
    for (int i = 0; i < ER.nqp; i++) a.w[i] = ER.w[i];
    a.slots = blk->slots; a.slots_t = blk_t ? blk_t->slots : nullptr;
    a.nzval = ctx->nzval; a.touched = ctx->touched; a.errflag = ctx->errflag;
    a.factor = factor; a.drop_tol = drop_tol; a.flags = flags;
    size_t per_cell = ((size_t)a.nqp * a.R * rs.nd + (a.same ? 0 : (size_t)a.nqp * a.R * cs.nd)) * sizeof(double);
    size_t geo_bytes = ctx->dim == 2 ? sizeof(Geo<2>) : sizeof(Geo<3>);
    int cpb = (int)std::min<size_t>(32, (96 * 1024) / (per_cell + geo_bytes));
    if (cpb < 1) return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "local tables too large for shared memory");
    a.cpb = cpb;
    size_t smem = cpb * (per_cell + geo_bytes);
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    void (*kern)(BilArgs) = nullptr;
#define FEMB_BK(RR) kern = ctx->dim == 2 ? k_bilinear<2, RR> : k_bilinear<3, RR>
    FEMB_DISPATCH_R(a.R, FEMB_BK)
#undef FEMB_BK
    // row-blocked kernel for the common (dim, nqp, R) combinations: one thread per (cell, row dof)
    void (*rows)(BilArgs) = nullptr;
    const int key = ctx->dim * 10000 + a.nqp * 100 + a.R;
    switch (key) {
        case 20301: rows = k_bilinear_rows<2, 3, 1>; break;   // 2-D order-2 rule: divergences, P0
        case 20302: rows = k_bilinear_rows<2, 3, 2>; break;   // 2-D Hdiv mass, P2 gradients
        case 20901: rows = k_bilinear_rows<2, 9, 1>; break;   // Example210 div-pressure (9-point Stroud rule)
        case 20902: rows = k_bilinear_rows<2, 9, 2>; break;   // P3 / Pk gradients in 2-D
        case 30401: rows = k_bilinear_rows<3, 4, 1>; break;   // 3-D order-2 rule: divergences, P0
        case 30403: rows = k_bilinear_rows<3, 4, 3>; break;   // 3-D Hdiv mass, P2 gradients
        case 31103: rows = k_bilinear_rows<3, 11, 3>; break;  // P3 gradients on tetrahedra (11-point rule)
        default: break;
    }
    if (rows && rs.nd <= 64) {
        kern = rows;
        cpb = std::max(1, 128 / rs.nd);
        per_cell = ((size_t)a.nqp * a.R * cs.nd + FEMB_ROWS_PAD) * sizeof(double);
        cpb = (int)std::min<size_t>(cpb, (96 * 1024) / (per_cell + geo_bytes));
        a.cpb = cpb;
        smem = cpb * (per_cell + geo_bytes);
    }
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (ctx->scatter == FEMB_SCATTER_COLORED) {
        FEMB_TRY(femb_build_coloring(ctx));
        a.perm = ctx->color_perm;
        a.atomic = 0;
        for (int c = 0; c < ctx->ncolors; c++) {
            a.p0 = ctx->color_ptr[c];
            a.p1 = ctx->color_ptr[c + 1];
            if (a.p1 == a.p0) continue;
            kern<<<(unsigned)((a.p1 - a.p0 + cpb - 1) / cpb), 128, smem, ctx->stream>>>(a);
            KERNEL_CHECK(ctx);
        }
    } else {
        a.perm = nullptr;
        a.atomic = 1;
        a.p0 = 0;
        a.p1 = ctx->ncells;
        if (ctx->ncells) {
            kern<<<(unsigned)((ctx->ncells + cpb - 1) / cpb), 128, smem, ctx->stream>>>(a);
            KERNEL_CHECK(ctx);
        }
    }
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// linear form: owner-computes gather over the dof -> cell adjacency (write-once, fixed order)
// ---------------------------------------------------------------------------------------------------
// LIGHT: H1 / L2 Identity needs no inverse map — only A (for x_qp) and |T|: skips the divisions of mapderiv!
template <int DIM, bool LIGHT, int R>
__global__ void __launch_bounds__(128) k_linear_gather(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes,
                                                       const double* __restrict__ cellvol, SpaceDev s, int op, QuadDev q,
                                                       const double* __restrict__ rv, const double* __restrict__ rd,
                                                       const int32_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj, int64_t ndofs,
                                                       RhsDev rhs, double factor, double* __restrict__ b) {
    int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= ndofs) return;
    double sum = 0.0;
    for (int32_t e = adj_ptr[d]; e < adj_ptr[d + 1]; e++) {
        uint32_t av = adj[e];
        int64_t cell = av / (uint32_t)s.nd;
        int k = (int)(av - cell * s.nd);
        Geo<DIM> g;
        if (LIGHT) {
            double xn[DIM + 1][DIM];
            load_nodes<DIM>(coords, cellnodes, cell, xn);
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                g.x0[d] = xn[0][d];
#pragma unroll
                for (int i = 0; i < DIM; i++) g.A[d][i] = xn[i + 1][d] - xn[0][d];
            }
            if (cellvol) {
                g.vol = __ldg(cellvol + cell);
            } else if (DIM == 2) {
                g.vol = fabs(g.A[0][0] * g.A[1][1] - g.A[0][1] * g.A[1][0]) * 0.5;
            } else {
                g.vol = fabs(g.A[0][0] * (g.A[1][1] * g.A[2][2] - g.A[1][2] * g.A[2][1]) + g.A[0][1] * (g.A[1][2] * g.A[2][0] - g.A[1][0] * g.A[2][2]) +
                             g.A[0][2] * (g.A[1][0] * g.A[2][1] - g.A[1][1] * g.A[2][0])) * (1.0 / 6.0);
            }
        } else {
            load_geo<DIM>(coords, cellnodes, cellvol, cell, g);
        }
        double temp = 0.0;
        for (int qp = 0; qp < q.nqp; qp++) {
            double x[DIM], f[R], v[R];
            eval_trafo<DIM>(g, q.xref + qp * DIM, x);
            eval_rhs<DIM, R>(rhs, x, cell, qp, q.nqp, f);
            eval_op<DIM, R>(s, op, g, cell, k, qp, rv, rd, v);
            double dt = 0.0;
#pragma unroll
            for (int r = 0; r < R; r++) dt += v[r] * f[r];
            temp += q.w[qp] * dt;
        }
        sum += temp * g.vol;
    }
    b[d] += factor * sum;
}

// Identity rhs of H1 / L2 spaces in two passes: the (possibly expensive) data f(x_q) is evaluated ONCE per cell and
// quadrature point instead of once per (dof, cell) pair, then gathered per dof in a fixed order (write-once).
template <int DIM, int NC>
__global__ void __launch_bounds__(128) k_cell_rhs(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol,
                                                  int64_t ncells, QuadDev q, RhsDev rhs, double factor, double* __restrict__ out) {
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cell >= ncells) return;
    double xn[DIM + 1][DIM], A[DIM][DIM];
    load_nodes<DIM>(coords, cellnodes, cell, xn);
#pragma unroll
    for (int d = 0; d < DIM; d++)
#pragma unroll
        for (int i = 0; i < DIM; i++) A[d][i] = xn[i + 1][d] - xn[0][d];
    double vol;
    if (cellvol) {
        vol = __ldg(cellvol + cell);
    } else if (DIM == 2) {
        vol = fabs(A[0][0] * A[1][1] - A[0][1] * A[1][0]) * 0.5;
    } else {
        vol = fabs(A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) + A[0][1] * (A[1][2] * A[2][0] - A[1][0] * A[2][2]) +
                   A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0])) * (1.0 / 6.0);
    }
    for (int qp = 0; qp < q.nqp; qp++) {
        double x[DIM], f[NC];
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            double sx = xn[0][d];
#pragma unroll
            for (int i = 0; i < DIM; i++) sx += A[d][i] * q.xref[qp * DIM + i];
            x[d] = sx;
        }
        eval_rhs<DIM, NC>(rhs, x, cell, qp, q.nqp, f);
#pragma unroll
        for (int c = 0; c < NC; c++) out[((size_t)cell * q.nqp + qp) * NC + c] = f[c] * (q.w[qp] * vol * factor);
    }
}

__global__ void __launch_bounds__(128) k_linear_gather_id(SpaceDev s, int nqp, int ncomp, const double* __restrict__ rv, const double* __restrict__ F,
                                                          const int32_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj, int64_t ndofs, double* __restrict__ b) {
    const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= ndofs) return;
    double sum = 0.0;
    for (int32_t e = adj_ptr[d]; e < adj_ptr[d + 1]; e++) {
        const uint32_t av = adj[e];
        const int64_t cell = av / (uint32_t)s.nd;
        const int k = (int)(av - cell * s.nd);
        const int sub = subset_of(s, cell, k);
        const double* Fc = F + (size_t)cell * nqp * ncomp;
        double t = 0.0;
        for (int qp = 0; qp < nqp; qp++)
            for (int c = 0; c < ncomp; c++) {
                const double phi = __ldg(rv + ((size_t)qp * s.nd_all + sub) * ncomp + c);
                if (phi != 0.0) t += phi * __ldg(Fc + qp * ncomp + c);  // component-blocked bases: most entries are exact zeros
            }
        sum += t;
    }
    b[d] += sum;
}

int launch_linear(femb_ctx* ctx, int ev, int brow, double factor, const RhsDev& rhs) {
    const FembEval& E = ctx->evals[ev];
    const FembSpace& s = ctx->spaces[E.space];
    if (brow < 0 || brow >= ctx->nrs || ctx->row_spaces[brow] != E.space) return femb_fail(ctx, FEMB_ERR_ARG, "evaluator space does not match block row %d", brow);
    if (E.resultdim != rhs.ncomp) return femb_fail(ctx, FEMB_ERR_ARG, "rhs has %d components, operator %d", rhs.ncomp, E.resultdim);
    FEMB_TRY(femb_build_adjacency(ctx, E.space));
    FEMB_TRY(femb_flush_zero(ctx, false, true));
    unsigned grid = (unsigned)((s.ndofs + 127) / 128);
    QuadDev q = quad_dev(E, ctx->dim);
    const bool light = s.family != FEMB_FAMILY_HDIV && s.family != FEMB_FAMILY_HCURL && E.op == FEMB_OP_IDENTITY;
    if (light && E.refvals && ctx->ncells) {
        const size_t need = (size_t)ctx->ncells * E.nqp * rhs.ncomp;
        if (ctx->cellrhs_n < need) {
            FEMB_TRY(femb_alloc(ctx, &ctx->cellrhs, need));
            ctx->cellrhs_n = need;
        }
        const unsigned gc = (unsigned)((ctx->ncells + 127) / 128);
#define FEMB_CR(NC)                                                                                                                                   \
    do {                                                                                                                                          \
        if (ctx->dim == 2) k_cell_rhs<2, NC><<<gc, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, q, rhs, factor, ctx->cellrhs); \
        else k_cell_rhs<3, NC><<<gc, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, q, rhs, factor, ctx->cellrhs);               \
    } while (0)
        if (rhs.ncomp == 1) FEMB_CR(1);
        else if (rhs.ncomp == 2) FEMB_CR(2);
        else if (rhs.ncomp == 3) FEMB_CR(3);
        else return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "rhs with %d components", rhs.ncomp);
#undef FEMB_CR
        KERNEL_CHECK(ctx);
        k_linear_gather_id<<<grid, 128, 0, ctx->stream>>>(space_dev(s), E.nqp, rhs.ncomp, E.refvals, ctx->cellrhs, s.adj_ptr, s.adj, s.ndofs,
                                                          ctx->bvec + ctx->row_off[brow]);
        KERNEL_CHECK(ctx);
        return FEMB_OK;
    }
#define FEMB_LIN(D, L, RR)                                                                                                                        \
    k_linear_gather<D, L, RR><<<grid, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, space_dev(s), E.op, q, E.refvals, E.refderivs, \
                                                             s.adj_ptr, s.adj, s.ndofs, rhs, factor, ctx->bvec + ctx->row_off[brow])
#define FEMB_LIN_R(RR)                       \
    do {                                     \
        if (ctx->dim == 2) {                 \
            if (light) FEMB_LIN(2, true, RR); \
            else FEMB_LIN(2, false, RR);     \
        } else {                             \
            if (light) FEMB_LIN(3, true, RR); \
            else FEMB_LIN(3, false, RR);     \
        }                                    \
    } while (0)
    if (E.resultdim == 1) FEMB_LIN_R(1);
    else if (E.resultdim == 2) FEMB_LIN_R(2);
    else if (E.resultdim == 3) FEMB_LIN_R(3);
    else return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "rhs with %d components", E.resultdim);
#undef FEMB_LIN_R
#undef FEMB_LIN
    KERNEL_CHECK(ctx);
    return FEMB_OK;
}


// ---------------------------------------------------------------------------------------------------
// C ABI entry points
// ---------------------------------------------------------------------------------------------------
int check_eval(femb_ctx* ctx, int id) {
    if (id < 0 || id >= FEMB_MAX_EVALS || !ctx->evals[id].set) return femb_fail(ctx, FEMB_ERR_ARG, "unknown evaluator %d", id);
    return FEMB_OK;
}

int require_pattern(femb_ctx* ctx) {
    if (!ctx->colptr || ctx->blocks.empty()) return femb_fail(ctx, FEMB_ERR_ARG, "no pattern: call femb_pattern_build or femb_pattern_adopt first");
    return FEMB_OK;
}

extern "C" int32_t femb_assemble_bilinear(femb_ctx* ctx, int32_t er, int32_t ec, int32_t brow, int32_t bcol, double factor, int32_t flags,
                                          double drop_tol) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(check_eval(ctx, er));
    FEMB_TRY(check_eval(ctx, ec));
    FEMB_TRY(require_pattern(ctx));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->dirty_A = true;
    FEMB_TRY(femb_timer_begin(ctx));
    int st = FEMB_ERR_UNSUPPORTED;
    if (ctx->scatter == FEMB_SCATTER_GATHER && !ctx->touched && !(drop_tol > 0.0)) {
        // HDIVRT0 on tetrahedra: closed-form local matrices, owner-computes mass gather, conflict-free divergence blocks
        if (er == ec && flags == 0 && ctx->evals[er].op == FEMB_OP_IDENTITY) {
            st = launch_rt0_mass_gather(ctx, er, brow, bcol, factor);
            if (st == FEMB_ERR_UNSUPPORTED) st = launch_bdm1_mass_gather(ctx, er, brow, bcol, factor);
        }
        else if (ctx->evals[er].op == FEMB_OP_DIVERGENCE && ctx->evals[ec].op == FEMB_OP_IDENTITY && (flags & ~FEMB_BIL_TRANSPOSE_COPY) == 0)
            st = launch_rt0_div(ctx, er, ec, brow, bcol, factor, flags);
    }
    if (st == FEMB_ERR_UNSUPPORTED) st = launch_bilinear(ctx, er, ec, ctx->evals[er].op, ctx->evals[ec].op, brow, bcol, factor, flags, drop_tol);
    if (st != FEMB_OK) return st;
    return femb_timer_end(ctx);
}

extern "C" int32_t femb_assemble_hdiv_mixed(femb_ctx* ctx, int32_t e_id, int32_t e_div, int32_t e_q, double factor_mass, double factor_div) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(check_eval(ctx, e_id));
    FEMB_TRY(check_eval(ctx, e_div));
    FEMB_TRY(check_eval(ctx, e_q));
    FEMB_TRY(require_pattern(ctx));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->scatter == FEMB_SCATTER_GATHER && !ctx->touched) {
        FEMB_TRY(femb_timer_begin(ctx));
        const int st = launch_rt0_mixed(ctx, e_id, e_div, e_q, factor_mass, factor_div);
        if (st == FEMB_OK) {
            ctx->dirty_A = true;
            return femb_timer_end(ctx);
        }
        if (st != FEMB_ERR_UNSUPPORTED) return st;
    }
    // everything else: the two assemblies one after the other (last_ms then reports the second one only)
    FEMB_TRY(femb_assemble_bilinear(ctx, e_id, e_id, 0, 0, factor_mass, 0, 0.0));
    return femb_assemble_bilinear(ctx, e_div, e_q, 0, 1, factor_div, FEMB_BIL_TRANSPOSE_COPY, 0.0);
}

extern "C" int32_t femb_assemble_linear(femb_ctx* ctx, int32_t ev, int32_t brow, double factor, int32_t rhs_kind, const double* data) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(check_eval(ctx, ev));
    if (!ctx->bvec) return femb_fail(ctx, FEMB_ERR_ARG, "femb_matrix_layout must precede femb_assemble_linear");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RhsDev rhs;
    double* dev_f = nullptr;
    FEMB_TRY(make_rhs(ctx, rhs_kind, data, ctx->evals[ev].resultdim, ctx->evals[ev].nqp, rhs, &dev_f));
    ctx->dirty_b = true;
    int st = femb_timer_begin(ctx);
    if (st == FEMB_OK && rhs_kind != FEMB_RHS_NONE) st = launch_linear(ctx, ev, brow, factor, rhs);
    if (st == FEMB_OK) st = femb_timer_end(ctx);
    cudaFree(dev_f);
    return st;
}

