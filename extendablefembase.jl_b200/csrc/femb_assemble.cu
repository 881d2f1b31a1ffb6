// Numeric phase: the five stages of the reference's cell loop, batched over all cells.
//   stage 1  affine map / determinant          (L2GTransformer; feevaluator.jl:344-363)
//   stage 2  push-forward / Piola               (feevaluator_h1.jl, feevaluator_hdiv.jl)
//   stage 3  quadrature-weighted contraction    (Example200:139-146, Example210:284-301,322-355)
//   stage 4  scatter into the CSC image         (Example200:149-161, Example210:368-382)
//   stage 5  right-hand side                    (Example200:165-178, Example210:304-318,358-363)
#include <algorithm>

#include "femb_device.cuh"
#include "femb_internal.h"

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

static SpaceDev space_dev(const FembSpace& s) {
    SpaceDev d;
    d.family = s.family; d.ncomp = s.ncomp; d.nd = s.nd; d.nd_all = s.nd_all; d.handler = s.handler; d.nsign = s.nsign;
    d.celldofs = s.celldofs; d.signs = s.signs; d.orient = s.orient;
    return d;
}

// apply the lazy zeroing requested by femb_matrix_zero / femb_vector_zero for kernels that accumulate
int femb_flush_zero(femb_ctx* ctx, bool mat, bool vec) {
    if (mat && ctx->zero_pending_A) {
        if (ctx->nzval) CUDA_TRY(ctx, cudaMemsetAsync(ctx->nzval, 0, sizeof(double) * ctx->nnz, ctx->stream));
        ctx->zero_pending_A = false;
    }
    if (vec && ctx->zero_pending_b) {
        if (ctx->bvec) CUDA_TRY(ctx, cudaMemsetAsync(ctx->bvec, 0, sizeof(double) * ctx->nrows, ctx->stream));
        ctx->zero_pending_b = false;
    }
    return FEMB_OK;
}

static int make_rhs(femb_ctx* ctx, int kind, const double* p, int ncomp, int nqp, RhsDev& r, double** dev_fvals) {
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

struct QuadDev {
    int nqp;
    double xref[FEMB_MAX_QP * 3];
    double w[FEMB_MAX_QP];
};

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

static QuadDev quad_dev(const FembEval& e, int dim) {
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

__device__ __forceinline__ void scatter_add(double* nzval, uint8_t* touched, int32_t* errflag, int32_t slot, double v, int atomic) {
    if (slot < 0) {
        if (v != 0.0) atomicAdd(errflag, 1);
        return;
    }
    if (atomic) atomicAdd(nzval + slot, v);
    else nzval[slot] += v;
    if (touched) touched[slot] = 1;
}

template <int DIM>
__global__ void __launch_bounds__(128) k_bilinear(BilArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Geo<DIM>* geo = reinterpret_cast<Geo<DIM>*>(smem_raw);
    double* rowv = reinterpret_cast<double*>(geo + a.cpb);
    const int ndr = a.rs.nd, ndc = a.cs.nd;
    const int per_r = a.nqp * a.R * ndr, per_c = a.nqp * a.R * ndc;
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
        double v[9];
        eval_op<DIM>(a.rs, a.op_r, geo[cl], cell, dof, qp, a.rv_r, a.rd_r, v);
        for (int r = 0; r < a.R; r++) rowv[(size_t)cl * per_r + (qp * a.R + r) * ndr + dof] = v[r];
    }
    if (!a.same) {
        for (int i = tid; i < ncl * a.nqp * ndc; i += nt) {
            int cl = i / (a.nqp * ndc), rem = i - cl * (a.nqp * ndc);
            int qp = rem / ndc, dof = rem - qp * ndc;
            int64_t cell = a.perm ? a.perm[base + cl] : base + cl;
            double v[9];
            eval_op<DIM>(a.cs, a.op_c, geo[cl], cell, dof, qp, a.rv_c, a.rd_c, v);
            for (int r = 0; r < a.R; r++) colv[(size_t)cl * per_c + (qp * a.R + r) * ndc + dof] = v[r];
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
            for (int r = 0; r < a.R; r++) d += rp[(qp * a.R + r) * ndr] * cp[(qp * a.R + r) * ndc];
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
template <int DIM, int NQ, int R>
__global__ void __launch_bounds__(128) k_bilinear_rows(BilArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Geo<DIM>* geo = reinterpret_cast<Geo<DIM>*>(smem_raw);
    double* colv = reinterpret_cast<double*>(geo + a.cpb);  // [cl][qp * R + r][k]
    const int ndr = a.rs.nd, ndc = a.cs.nd;
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
            double v[9];
            eval_op<DIM>(a.rs, a.op_r, geo[cl], cell, j, qp, a.rv_r, a.rd_r, v);
#pragma unroll
            for (int r = 0; r < R; r++) rv[qp * R + r] = v[r];
        }
        if (a.same) {
#pragma unroll
            for (int i = 0; i < NQ * R; i++) colv[((size_t)cl * (NQ * R) + i) * ndc + j] = rv[i];
        }
    }
    if (!a.same) {
        for (int i = tid; i < ncl * NQ * ndc; i += nt) {
            const int c2 = i / (NQ * ndc), rem = i - c2 * (NQ * ndc);
            const int qp = rem / ndc, k = rem - qp * ndc;
            double v[9];
            eval_op<DIM>(a.cs, a.op_c, geo[c2], a.perm ? a.perm[base + c2] : base + c2, k, qp, a.rv_c, a.rd_c, v);
#pragma unroll
            for (int r = 0; r < R; r++) colv[((size_t)c2 * (NQ * R) + qp * R + r) * ndc + k] = v[r];
        }
    }
    __syncthreads();
    if (!active) return;
    const bool sym = a.flags & FEMB_BIL_SYMMETRIC_UPPER;
    const double vol = (a.flags & FEMB_BIL_NO_VOLUME) ? 1.0 : geo[cl].vol;
    const int per = ndr * ndc;
    const double* cv = colv + (size_t)cl * (NQ * R) * ndc;
    const int32_t* sl = a.slots + cell * per + j * ndc;
    for (int k = sym ? j : 0; k < ndc; k++) {
        double temp = 0.0;
#pragma unroll
        for (int qp = 0; qp < NQ; qp++) {
            double d = 0.0;
#pragma unroll
            for (int r = 0; r < R; r++) d += rv[qp * R + r] * cv[(qp * R + r) * ndc + k];
            temp += a.w[qp] * d;
        }
        const double val = sym ? temp * (a.factor * vol) : (a.factor * temp) * vol;
        if (sym && !(fabs(val) > a.drop_tol)) continue;
        scatter_add(a.nzval, a.touched, a.errflag, __ldg(sl + k), val, a.atomic);
        if (sym && k > j) scatter_add(a.nzval, a.touched, a.errflag, __ldg(a.slots + cell * per + k * ndc + j), val, a.atomic);
        if (a.slots_t) scatter_add(a.nzval, a.touched, a.errflag, __ldg(a.slots_t + cell * per + k * ndr + j), val, a.atomic);
    }
}

static const FembBlock* find_block(const femb_ctx* ctx, int brow, int bcol) {
    for (auto& b : ctx->blocks)
        if (b.brow == brow && b.bcol == bcol) return &b;
    return nullptr;
}

static int launch_bilinear(femb_ctx* ctx, int er, int ec, int op_r, int op_c, int brow, int bcol, double factor, int flags, double drop_tol) {
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
    void (*kern)(BilArgs) = ctx->dim == 2 ? k_bilinear<2> : k_bilinear<3>;
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
        per_cell = (size_t)a.nqp * a.R * cs.nd * sizeof(double);
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
template <int DIM, bool LIGHT>
__global__ void __launch_bounds__(128) k_linear_gather(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes,
                                                       const double* __restrict__ cellvol, SpaceDev s, int op, int R, QuadDev q,
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
            double x[DIM], f[3], v[9];
            eval_trafo<DIM>(g, q.xref + qp * DIM, x);
            eval_rhs<DIM>(rhs, x, cell, qp, q.nqp, f);
            eval_op<DIM>(s, op, g, cell, k, qp, rv, rd, v);
            double dt = 0.0;
            for (int r = 0; r < R; r++) dt += v[r] * f[r];
            temp += q.w[qp] * dt;
        }
        sum += temp * g.vol;
    }
    b[d] += factor * sum;
}

// Identity rhs of H1 / L2 spaces in two passes: the (possibly expensive) data f(x_q) is evaluated ONCE per cell and
// quadrature point instead of once per (dof, cell) pair, then gathered per dof in a fixed order (write-once).
template <int DIM>
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
        double x[DIM], f[3];
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            double sx = xn[0][d];
#pragma unroll
            for (int i = 0; i < DIM; i++) sx += A[d][i] * q.xref[qp * DIM + i];
            x[d] = sx;
        }
        eval_rhs<DIM>(rhs, x, cell, qp, q.nqp, f);
        for (int c = 0; c < rhs.ncomp; c++) out[((size_t)cell * q.nqp + qp) * rhs.ncomp + c] = f[c] * (q.w[qp] * vol * factor);
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

static int launch_linear(femb_ctx* ctx, int ev, int brow, double factor, const RhsDev& rhs) {
    const FembEval& E = ctx->evals[ev];
    const FembSpace& s = ctx->spaces[E.space];
    if (brow < 0 || brow >= ctx->nrs || ctx->row_spaces[brow] != E.space) return femb_fail(ctx, FEMB_ERR_ARG, "evaluator space does not match block row %d", brow);
    if (E.resultdim != rhs.ncomp) return femb_fail(ctx, FEMB_ERR_ARG, "rhs has %d components, operator %d", rhs.ncomp, E.resultdim);
    FEMB_TRY(femb_build_adjacency(ctx, E.space));
    FEMB_TRY(femb_flush_zero(ctx, false, true));
    unsigned grid = (unsigned)((s.ndofs + 127) / 128);
    QuadDev q = quad_dev(E, ctx->dim);
    const bool light = s.family != FEMB_FAMILY_HDIV && E.op == FEMB_OP_IDENTITY;
    if (light && E.refvals && ctx->ncells) {
        const size_t need = (size_t)ctx->ncells * E.nqp * rhs.ncomp;
        if (ctx->cellrhs_n < need) {
            FEMB_TRY(femb_alloc(ctx, &ctx->cellrhs, need));
            ctx->cellrhs_n = need;
        }
        const unsigned gc = (unsigned)((ctx->ncells + 127) / 128);
        if (ctx->dim == 2) k_cell_rhs<2><<<gc, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, q, rhs, factor, ctx->cellrhs);
        else k_cell_rhs<3><<<gc, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, q, rhs, factor, ctx->cellrhs);
        KERNEL_CHECK(ctx);
        k_linear_gather_id<<<grid, 128, 0, ctx->stream>>>(space_dev(s), E.nqp, rhs.ncomp, E.refvals, ctx->cellrhs, s.adj_ptr, s.adj, s.ndofs,
                                                          ctx->bvec + ctx->row_off[brow]);
        KERNEL_CHECK(ctx);
        return FEMB_OK;
    }
#define FEMB_LIN(D, L)                                                                                                                          \
    k_linear_gather<D, L><<<grid, 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, space_dev(s), E.op, E.resultdim, q, E.refvals, \
                                                         E.refderivs, s.adj_ptr, s.adj, s.ndofs, rhs, factor, ctx->bvec + ctx->row_off[brow])
    if (ctx->dim == 2) {
        if (light) FEMB_LIN(2, true);
        else FEMB_LIN(2, false);
    } else {
        if (light) FEMB_LIN(3, true);
        else FEMB_LIN(3, false);
    }
#undef FEMB_LIN
    KERNEL_CHECK(ctx);
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// fused P1 Poisson, owner-computes by column (stages 1-5 in one kernel, write-once)
// ---------------------------------------------------------------------------------------------------
struct P1Tab {
    int nqp;
    double wsum;                  // sum_qp w_qp  (P1 gradients do not depend on qp: Example200:139-145 collapses to wsum * dot)
    double w[FEMB_MAX_QP];
    double xref[FEMB_MAX_QP * 3];
    double phi[FEMB_MAX_QP * 4];  // Identity table [qp][dof]
};

struct P1Args {
    const double* coords;
    const double* cellvol;
    const int32_t* gslice_ptr;
    const uint32_t* gmap;  // SoA planes na, nb, w, cell (, nc): femb_internal.h
    size_t gstride;
    const int64_t* colptr;
    int64_t ndofs;
    double* nzval;
    uint8_t* touched;
    double* bvec;
    int32_t* errflag;
    double mu, drop_tol;
    int acc_A, acc_b;
    int maxlen;  // longest CSC column: shared-memory row `maxlen` is the dump row of the 2-D kernel
    int64_t slice_begin, slice_end;  // slices [begin, end) handled by this launch (chunked / streamed assembly)
    int maxrows;  // longest slice of the gather map; > 0 selects the staged variant (entries copied to shared memory with cp.async)
    int uniform_rows;  // > 0: every slice has exactly this many rows (row0 = slice * uniform_rows)
};

#define FEMB_P1_MAXQP 8
#ifndef FEMB_P1_PIPE
#define FEMB_P1_PIPE 2  // software-pipeline depth of the 2-D gather kernel (1: entries only, 2: entries + coordinates).
// Measured on B200 (32M triangles): depth 2 @ 62 regs / 8 CTAs per SM = depth 1 @ 48 regs / 10 CTAs = 0.77 ms; spilling variants are slower.
// The staged variant (entries via cp.async into shared memory) runs 0.69 ms; a multi-slice, double-buffered variant of it
// (2 or 4 slices per warp, next slice staged during the current one) measured 0.89-0.91 ms and was dropped.
#endif
struct P1Tab2 {
    int nqp;
    double mu_wsum;                     // mu * sum_qp w_qp
    double t[3][FEMB_P1_MAXQP][4];      // [k][qp] = {w_qp phi_k, phi_k, phi_ja, phi_jb} at xref_qp; (ja, jb) = the other local indices
};

// 1/x to ~1 ulp: MUFU.RCP64H seed + two Newton steps (no denormal / special-case slow path: x = det^2 > 0 of a
// non-degenerate cell).  Every visitor of a cell evaluates this on identical bits, so symmetry is preserved.
__device__ __forceinline__ double rcp_newton(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// one visit of the 2-D P1 gather: column k of the cell's local stiffness matrix (+ rhs) into the accumulators
template <bool TRACK, int RHS, bool HAS_VOL, int NQP>
__device__ __forceinline__ void p1_visit(const unsigned w, const unsigned cell, const double2 self, const double2 pa, const double2 pb, const double volr,
                                         const P1Tab2& tab, const RhsDev& rhs, const double drop_tol, const double mu_wsum, const unsigned sacc,
                                         double& diag, unsigned& dpos, unsigned& err, unsigned long long& mask, double& bsum) {
    constexpr unsigned nt = 128;
    const unsigned k = w & 3u;
    const double uax = pb.x - self.x, uay = pb.y - self.y;
    const double ubx = self.x - pa.x, uby = self.y - pa.y;
    const double usx = pa.x - pb.x, usy = pa.y - pb.y;
    const double px = (k == 0u) ? uax : usx, py = (k == 0u) ? uay : usy;
    const double det = px * uby - py * ubx;
    const double vol = HAS_VOL ? volr : fabs(det) * 0.5;
    const double scale = (mu_wsum * vol) * rcp_newton(det * det);
    const double vs = (usx * usx + usy * usy) * scale;
    const double va = (uax * usx + uay * usy) * scale;
    const double vb = (ubx * usx + uby * usy) * scale;
    const unsigned ps = (w >> 8) & 0xFFu, pA = (w >> 16) & 0xFFu, pB = w >> 24;
    const bool ks = fabs(vs) > drop_tol, ka = fabs(va) > drop_tol, kb = fabs(vb) > drop_tol;
    err |= (unsigned)((ks && ps == 0xFFu) || (ka && pA == 0xFFu) || (kb && pB == 0xFFu));
    // diagonal A[c,c]: the same slot on every visit -> a register; off-diagonals: shared-memory rows.  The
    // branches are warp-uniform on structured meshes (same filtered edge in every lane) and cheap otherwise.
    if (ks && ps != 0xFFu) {
        diag += vs;
        dpos = ps;
    }
    if (ka && pA != 0xFFu) {
        const unsigned ad = sacc + pA * (nt * 8u);
        sts64(ad, lds64(ad) + va);
        if (TRACK) mask |= 1ull << (pA & 63u);
    }
    if (kb && pB != 0xFFu) {
        const unsigned ad = sacc + pB * (nt * 8u);
        sts64(ad, lds64(ad) + vb);
        if (TRACK) mask |= 1ull << (pB & 63u);
    }
    if (RHS != FEMB_RHS_NONE) {
        double temp = 0.0;
        const int nqp = NQP ? NQP : tab.nqp;
#pragma unroll
        for (int qp = 0; qp < nqp; qp++) {
            const double* t = tab.t[k][qp];
            double f;
            if (RHS == FEMB_RHS_AFFINE) {  // f(x_qp); x_qp = sum_j phi_j(xref_qp) x_j  (= A xref_qp + x_1, eval_trafo!, Example200:171)
                const double xq = t[1] * self.x + t[2] * pa.x + t[3] * pb.x;
                const double yq = t[1] * self.y + t[2] * pa.y + t[3] * pb.y;
                f = rhs.p[0] + rhs.p[1] * xq + rhs.p[2] * yq;
            } else {
                f = __ldg(rhs.fvals + (size_t)cell * nqp + qp);
            }
            temp += t[0] * f;
        }
        bsum += temp * vol;
    }
    }

// ---- 2-D scalar P1, the headline kernel ------------------------------------------------------------------------
// Fused stages 1-5 (Example200:133-179), owner-computes by matrix column: one thread per column c (= dof = node),
// one warp per slice of the sliced-ELL gather map (femb_pattern.cu), streamed with fully coalesced loads.  Every
// entry names one cell adjacent to c by its two other nodes; the thread rebuilds the cell's affine map from
// Coordinates alone (its own node is loaded once), forms column k of the local stiffness matrix and adds it into
// the <= 255 slots of CSC column c, accumulated in shared memory and written ONCE in a fixed order:
// bit-reproducible, no atomics, no memset of nzval.
//  * with u_self = x_a - x_b, u_a = x_b - x_self, u_b = x_self - x_a (edges opposite to each node, taken in the
//    cell's cyclic local order up to one global sign that cancels in every product) the cofactor vectors are the
//    rotated edges, so  Aloc[j,k] = mu |T| wsum (u_j . u_k) / det^2 : one reciprocal per visit instead of the
//    reference's four divisions (mapderiv!), no local re-ordering;
//  * det is the cross product of the same two edges the reference's A = [x2-x1 | x3-x1] uses (chosen by k), so
//    every visitor of a cell gets bit-identical scale factors and the assembled matrix is exactly symmetric;
//  * software pipeline: the gather-map entry two visits ahead and the coordinates / volume of the next visit are
//    in flight while the current visit is computed;
//  * accumulation is branch-free: filtered (|v| <= drop_tol, Example200:153) contributions go to a dump row.
//  * STAGE: the slice's gather-map entries (all visits) are copied to shared memory with cp.async (LDGSTS) at the very
//    start, so the streamed part of the input arrives in ONE memory latency instead of one per visit and costs no
//    registers; only the coordinate / volume loads of the next visit stay in a register pipeline.
#ifndef FEMB_P1_MINB
#define FEMB_P1_MINB 8  // resident CTAs per SM the 2-D gather kernel is compiled for (register cap 65536 / (128 * MINB))
#endif
template <bool TRACK, int RHS, bool HAS_VOL, int NQP, bool STAGE, int MINB = FEMB_P1_MINB>  // NQP: compile-time number of quadrature points, 0 = tab.nqp
__global__ void __launch_bounds__(128, MINB) k_p1_poisson_gather2d(const P1Args a, const RhsDev rhs, const P1Tab2 tab) {
    constexpr unsigned nt = 128;
    constexpr bool NEED_CELL = HAS_VOL || RHS == FEMB_RHS_QPVALUES;
    extern __shared__ double acc[];  // [maxlen + 1][128], then (STAGE) u32 [4 warps][4 planes][maxrows][32]
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const int64_t slice = a.slice_begin + (int64_t)blockIdx.x * 4 + (tid >> 5);
    if (slice >= a.slice_end) return;  // whole warp out of range
    const int64_t c = slice * 32 + lane;
    const bool valid = c < a.ndofs;
    int row0, nrows;
    if (a.uniform_rows > 0) {
        row0 = (int)slice * a.uniform_rows;
        nrows = a.uniform_rows;
    } else {
        row0 = __ldg(a.gslice_ptr + slice);
        nrows = __ldg(a.gslice_ptr + slice + 1) - row0;
    }
    const double2* __restrict__ xy = reinterpret_cast<const double2*>(a.coords);
    int64_t base = 0;
    int len = 0;
    double2 self = make_double2(0.0, 0.0);
    if (valid) {
        base = a.colptr[c] - 1;
        len = (int)(a.colptr[c + 1] - 1 - base);
        self = __ldg(xy + c);
    }
    const unsigned sacc = (unsigned)__cvta_generic_to_shared(acc) + tid * 8u;
    for (int i = 0; i < len; i++) sts64(sacc + i * (nt * 8u), 0.0);
    unsigned long long mask = 0ull;
    double bsum = 0.0, diag = 0.0;
    unsigned err = 0u, dpos = 0xFFu;
    const double drop_tol = a.drop_tol, mu_wsum = tab.mu_wsum;

    const uint32_t* __restrict__ ga = a.gmap + ((size_t)row0 << 5) + lane;
    const size_t gs = a.gstride;
    if (STAGE) {
        const unsigned sst = (unsigned)__cvta_generic_to_shared(acc + (size_t)(a.maxlen + 1) * nt) + ((tid >> 5) * 4u * (unsigned)a.maxrows * 32u + lane) * 4u;
        const unsigned pst = (unsigned)a.maxrows * 128u;  // plane stride in bytes
        for (int r = 0; r < nrows; r++) {
            const uint32_t* g = ga + ((size_t)r << 5);
            const unsigned d = sst + (unsigned)r * 128u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(g) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + pst), "l"(g + gs) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 2u * pst), "l"(g + 2 * gs) : "memory");
            if (NEED_CELL) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 3u * pst), "l"(g + 3 * gs) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        auto lds32 = [](unsigned addr) {
            unsigned v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
            return v;
        };
        double2 pa1 = make_double2(0.0, 0.0), pb1 = pa1;
        double vol1 = 0.0;
        if (nrows > 0) {
            pa1 = __ldg(xy + lds32(sst));
            pb1 = __ldg(xy + lds32(sst + pst));
            if (HAS_VOL) vol1 = __ldg(a.cellvol + lds32(sst + 3u * pst));
        }
        for (int r = 0; r < nrows; r++) {
            const unsigned d = sst + (unsigned)r * 128u;
            const unsigned w = lds32(d + 2u * pst);
            const unsigned cell = NEED_CELL ? lds32(d + 3u * pst) : 0u;
            const double2 pa = pa1, pb = pb1;
            const double volr = vol1;
            if (r + 1 < nrows) {  // coordinates / volume of the next visit
                pa1 = __ldg(xy + lds32(d + 128u));
                pb1 = __ldg(xy + lds32(d + 128u + pst));
                if (HAS_VOL) vol1 = __ldg(a.cellvol + lds32(d + 128u + 3u * pst));
            }
            if (w == 0xFFFFFFFFu) continue;
            p1_visit<TRACK, RHS, HAS_VOL, NQP>(w, cell, self, pa, pb, volr, tab, rhs, drop_tol, mu_wsum, sacc, diag, dpos, err, mask, bsum);
        }
    } else {
#if FEMB_P1_PIPE == 1
    // shallow pipeline: only the next visit's entry is in flight; coordinates are loaded at the top of the visit
    unsigned w1 = 0xFFFFFFFFu, a1 = 0u, b1 = 0u, c1 = 0u;
    if (nrows > 0) {
        w1 = __ldg(ga + 2 * gs); a1 = __ldg(ga); b1 = __ldg(ga + gs);
        if (NEED_CELL) c1 = __ldg(ga + 3 * gs);
    }
    for (int r = 0; r < nrows; r++) {
        const unsigned w = w1, cell = c1;
        const double2 pa = __ldg(xy + a1), pb = __ldg(xy + b1);
        const double volr = HAS_VOL ? __ldg(a.cellvol + c1) : 0.0;
        if (r + 1 < nrows) {
            const uint32_t* g1 = ga + ((size_t)(r + 1) << 5);
            w1 = __ldg(g1 + 2 * gs); a1 = __ldg(g1); b1 = __ldg(g1 + gs);
            if (NEED_CELL) c1 = __ldg(g1 + 3 * gs);
        }
#else
    const uint32_t* __restrict__ gb = ga + gs;
    const uint32_t* __restrict__ gw = gb + gs;
    const uint32_t* __restrict__ gc = gw + gs;
    // pipeline registers: (w1, a1, b1, c1) = entry of visit r, (w2, ...) = visit r + 1; pa1/pb1/vol1 = data of visit r
    unsigned w1 = 0xFFFFFFFFu, a1 = 0u, b1 = 0u, c1 = 0u, w2 = 0xFFFFFFFFu, a2 = 0u, b2 = 0u, c2 = 0u;
    if (nrows > 0) {
        w1 = __ldg(gw); a1 = __ldg(ga); b1 = __ldg(gb);
        if (NEED_CELL) c1 = __ldg(gc);
    }
    if (nrows > 1) {
        w2 = __ldg(gw + 32); a2 = __ldg(ga + 32); b2 = __ldg(gb + 32);
        if (NEED_CELL) c2 = __ldg(gc + 32);
    }
    double2 pa1 = __ldg(xy + a1), pb1 = __ldg(xy + b1);
    double vol1 = HAS_VOL ? __ldg(a.cellvol + c1) : 0.0;
    for (int r = 0; r < nrows; r++) {
        const unsigned w = w1, cell = c1;
        const double2 pa = pa1, pb = pb1;
        const double volr = vol1;
        w1 = w2; a1 = a2; b1 = b2; c1 = c2;
        if (r + 2 < nrows) {
            const size_t o = (size_t)(r + 2) << 5;
            w2 = __ldg(gw + o); a2 = __ldg(ga + o); b2 = __ldg(gb + o);
            if (NEED_CELL) c2 = __ldg(gc + o);
        } else {
            w2 = 0xFFFFFFFFu; a2 = 0u; b2 = 0u; c2 = 0u;
        }
        pa1 = __ldg(xy + a1);
        pb1 = __ldg(xy + b1);
        if (HAS_VOL) vol1 = __ldg(a.cellvol + c1);
#endif
        if (w == 0xFFFFFFFFu) continue;
        p1_visit<TRACK, RHS, HAS_VOL, NQP>(w, cell, self, pa, pb, volr, tab, rhs, drop_tol, mu_wsum, sacc, diag, dpos, err, mask, bsum);
    }
    }
    if (!valid) return;
    if (err) atomicAdd(a.errflag, 1);
    if (dpos != 0xFFu) {
        sts64(sacc + dpos * (nt * 8u), diag);  // the diagonal row received no shared-memory update
        if (TRACK) mask |= 1ull << (dpos & 63u);
    }
    double* out = a.nzval + base;
    if (a.acc_A) {
        for (int i = 0; i < len; i++) out[i] += lds64(sacc + i * (nt * 8u));
    } else {
        for (int i = 0; i < len; i++) out[i] = lds64(sacc + i * (nt * 8u));
    }
    if (TRACK) {
        for (int i = 0; i < len; i++)
            if ((mask >> i) & 1ull) a.touched[base + i] = 1;
    }
    if (RHS != FEMB_RHS_NONE) {
        if (a.acc_b) a.bvec[c] += bsum;
        else a.bvec[c] = bsum;
    }
}

// ---- generic scalar P1 (instantiated for tetrahedra) ------------------------------------------------------------
// Same owner-computes scheme; the cell's nodes are put back into local order (selects on k) and
//   Aloc[j,k] = mu |T| wsum (C_j . C_k) / det^2,   C = cofactor matrix of the Jacobian A (grad phi_j = C_j / det).
template <int DIM, bool TRACK, int RHS>  // RHS: FEMB_RHS_NONE / _AFFINE / _QPVALUES
__global__ void __launch_bounds__(128) k_p1_poisson_gather(const P1Args a, const RhsDev rhs, const P1Tab tab) {
    static_assert(DIM == 3, "the 2-D case has its own kernel");
    constexpr int ND = DIM + 1;
    constexpr unsigned nt = 128;
    constexpr unsigned MISSING = 0x3Fu;
    extern __shared__ double acc[];  // [maxlen][128]
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const int64_t slice = a.slice_begin + (int64_t)blockIdx.x * 4 + (tid >> 5);
    const int64_t c = slice * 32 + lane;
    if (slice >= a.slice_end) return;  // whole warp out of range
    const bool valid = c < a.ndofs;
    const int row0 = __ldg(a.gslice_ptr + slice), row1 = __ldg(a.gslice_ptr + slice + 1);
    int64_t base = 0;
    int len = 0;
    double self[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) self[d] = 0.0;
    if (valid) {
        base = a.colptr[c] - 1;
        len = (int)(a.colptr[c + 1] - 1 - base);
#pragma unroll
        for (int d = 0; d < 3; d++) self[d] = __ldg(a.coords + c * 3 + d);
    }
    for (int i = 0; i < len; i++) acc[i * nt + tid] = 0.0;
    unsigned long long mask = 0ull;
    double bsum = 0.0;
    unsigned err = 0u;
    const uint32_t* __restrict__ ga = a.gmap + ((size_t)row0 << 5) + lane;
    for (int r = row0; r < row1; r++, ga += 32) {
        const unsigned w = __ldg(ga + 2 * a.gstride);
        if (w == 0xFFFFFFFFu) continue;
        const unsigned k = w & 3u;
        const unsigned cell = __ldg(ga + 3 * a.gstride);
        const unsigned on[3] = {__ldg(ga), __ldg(ga + a.gstride), __ldg(ga + 4 * a.gstride)};
        double x[ND][DIM], o[3][3];  // x: the cell's nodes in local order
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int d = 0; d < 3; d++) o[i][d] = __ldg(a.coords + (size_t)on[i] * 3 + d);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            x[0][d] = (k == 0) ? self[d] : o[0][d];
            x[1][d] = (k == 0) ? o[0][d] : ((k == 1) ? self[d] : o[1][d]);
            x[2][d] = (k <= 1) ? o[1][d] : ((k == 2) ? self[d] : o[2][d]);
            x[3][d] = (k == 3) ? self[d] : o[2][d];
        }
        double A[DIM][DIM], cf[ND][DIM], det;  // cf[j] = cofactor vector: grad phi_j = cf[j] / det
#pragma unroll
        for (int d = 0; d < DIM; d++)
#pragma unroll
            for (int i = 0; i < DIM; i++) A[d][i] = x[i + 1][d] - x[0][d];
#pragma unroll
        for (int r2 = 0; r2 < 3; r2++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int ra = (r2 + 1) % 3, rb = (r2 + 2) % 3, ja = (j + 1) % 3, jb = (j + 2) % 3;
                cf[j + 1][r2] = A[ra][ja] * A[rb][jb] - A[ra][jb] * A[rb][ja];
            }
        det = A[0][0] * cf[1][0] + A[0][1] * cf[2][0] + A[0][2] * cf[3][0];
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            double s0 = 0.0;
#pragma unroll
            for (int j = 1; j < ND; j++) s0 -= cf[j][d];
            cf[0][d] = s0;
        }
        const double vol = a.cellvol ? __ldg(a.cellvol + cell) : fabs(det) * (1.0 / 6.0);
        const double scale = (a.mu * vol) * tab.wsum / (det * det);
        double ck[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            double v = cf[0][d];
#pragma unroll
            for (int m = 1; m < ND; m++) v = (k == (unsigned)m) ? cf[m][d] : v;
            ck[d] = v;
        }
#pragma unroll
        for (int j = 0; j < ND; j++) {
            double d = 0.0;
#pragma unroll
            for (int kk = 0; kk < DIM; kk++) d += cf[j][kk] * ck[kk];
            const double val = d * scale;
            const unsigned pos = (w >> (2 + 6 * j)) & 0x3Fu;
            if (fabs(val) > a.drop_tol) {
                if (pos == MISSING) {
                    err = 1u;
                } else {
                    acc[pos * nt + tid] += val;
                    if (TRACK) mask |= 1ull << (pos & 63u);
                }
            }
        }
        if (RHS != FEMB_RHS_NONE) {
            double temp = 0.0;
            for (int qp = 0; qp < tab.nqp; qp++) {
                double f;
                if (RHS == FEMB_RHS_AFFINE) {  // f(x_qp), x_qp = A xref_qp + x_1  (eval_trafo!, Example200:171)
                    f = rhs.p[0];
#pragma unroll
                    for (int d = 0; d < DIM; d++) {
                        double sx = x[0][d];
#pragma unroll
                        for (int i = 0; i < DIM; i++) sx += A[d][i] * tab.xref[qp * DIM + i];
                        f += rhs.p[1 + d] * sx;
                    }
                } else {
                    f = __ldg(rhs.fvals + (size_t)cell * tab.nqp + qp);
                }
                temp += tab.w[qp] * tab.phi[qp * ND + k] * f;
            }
            bsum += temp * vol;
        }
    }
    if (!valid) return;
    if (err) atomicAdd(a.errflag, 1);
    double* out = a.nzval + base;
    if (a.acc_A) {
        for (int i = 0; i < len; i++) out[i] += acc[i * nt + tid];
    } else {
        for (int i = 0; i < len; i++) out[i] = acc[i * nt + tid];
    }
    if (TRACK) {
        for (int i = 0; i < len; i++)
            if ((mask >> i) & 1ull) a.touched[base + i] = 1;
    }
    if (RHS != FEMB_RHS_NONE) {
        if (a.acc_b) a.bvec[c] += bsum;
        else a.bvec[c] = bsum;
    }
}

template <int DIM>
static int launch_p1_gather_dim(femb_ctx* ctx, const P1Args& a, const RhsDev& rhs, const P1Tab& tab, unsigned grid, size_t smem, cudaStream_t stream) {
    const bool track = a.touched != nullptr;
#define FEMB_P1_LAUNCH(T, R)                                                                                                   \
    do {                                                                                                                       \
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_p1_poisson_gather<DIM, T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_p1_poisson_gather<DIM, T, R><<<grid, 128, smem, stream>>>(a, rhs, tab);                                           \
    } while (0)
    if (rhs.kind == FEMB_RHS_AFFINE) {
        if (track) FEMB_P1_LAUNCH(true, FEMB_RHS_AFFINE);
        else FEMB_P1_LAUNCH(false, FEMB_RHS_AFFINE);
    } else if (rhs.kind == FEMB_RHS_QPVALUES) {
        if (track) FEMB_P1_LAUNCH(true, FEMB_RHS_QPVALUES);
        else FEMB_P1_LAUNCH(false, FEMB_RHS_QPVALUES);
    } else {
        if (track) FEMB_P1_LAUNCH(true, FEMB_RHS_NONE);
        else FEMB_P1_LAUNCH(false, FEMB_RHS_NONE);
    }
#undef FEMB_P1_LAUNCH
    KERNEL_CHECK(ctx);
    return FEMB_OK;
}

struct P1Launch {
    P1Args a;
    P1Tab tab;
    P1Tab2 t2;
    RhsDev rhs;
    size_t smem;
    int64_t nslices;
};

static int prepare_p1_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol, P1Launch& L) {
    const FembEval& EG = ctx->evals[eg];
    const FembSpace& s = ctx->spaces[EG.space];
    P1Tab& tab = L.tab;
    tab.nqp = EG.nqp;
    const int ND = ctx->dim + 1;
    std::vector<double> phi((size_t)EG.nqp * ND, 0.0);
    if (rhs.kind != FEMB_RHS_NONE) {
        const FembEval& EI = ctx->evals[ei];
        if (EI.h_refvals.size() != phi.size()) return femb_fail(ctx, FEMB_ERR_ARG, "Identity table has the wrong size for P1");
        phi = EI.h_refvals;
    }
    tab.wsum = 0.0;
    for (int i = 0; i < EG.nqp; i++) {
        tab.w[i] = EG.w[i];
        tab.wsum += EG.w[i];
    }
    for (int i = 0; i < EG.nqp * ctx->dim; i++) tab.xref[i] = EG.xref[i];
    for (int i = 0; i < EG.nqp * ND; i++) tab.phi[i] = phi[i];
    P1Args& a = L.a;
    a.coords = ctx->coords; a.cellvol = ctx->cellvol; a.gslice_ptr = ctx->gslice_ptr; a.gmap = ctx->gmap; a.gstride = ctx->gmap_stride;
    a.colptr = ctx->colptr; a.ndofs = s.ndofs; a.nzval = ctx->nzval; a.touched = ctx->touched; a.bvec = ctx->bvec; a.errflag = ctx->errflag;
    a.mu = mu; a.drop_tol = drop_tol;
    a.acc_A = ctx->zero_pending_A ? 0 : 1;
    a.acc_b = ctx->zero_pending_b ? 0 : 1;
    ctx->zero_pending_A = false;
    if (rhs.kind != FEMB_RHS_NONE) ctx->zero_pending_b = false;
    a.maxlen = std::max(1, ctx->max_collen);
    a.maxrows = 0;
    a.uniform_rows = ctx->gmap_uniform ? ctx->gmap_maxrows : 0;
    L.nslices = (s.ndofs + 31) / 32;
    a.slice_begin = 0;
    a.slice_end = L.nslices;
    L.rhs = rhs;
    L.smem = (size_t)a.maxlen * 128 * sizeof(double);
    if (ctx->dim == 2) {
        P1Tab2& t2 = L.t2;
        t2.nqp = EG.nqp;
        t2.mu_wsum = mu * tab.wsum;
        static const int others[3][2] = {{1, 2}, {0, 2}, {0, 1}};
        for (int k = 0; k < 3; k++)
            for (int q = 0; q < FEMB_P1_MAXQP; q++) {
                double lam[3] = {0.0, 0.0, 0.0};  // P1 barycentric coordinates of xref_q (h1_p1.jl:64-76)
                if (q < EG.nqp) {
                    lam[1] = EG.xref[q * 2];
                    lam[2] = EG.xref[q * 2 + 1];
                    lam[0] = 1.0 - lam[1] - lam[2];
                }
                double phik = (q < EG.nqp) ? phi[(size_t)q * 3 + k] : 0.0;  // Identity table of the evaluator (= lam[k] for P1)
                t2.t[k][q][0] = (q < EG.nqp ? EG.w[q] : 0.0) * phik;
                t2.t[k][q][1] = lam[k];
                t2.t[k][q][2] = lam[others[k][0]];
                t2.t[k][q][3] = lam[others[k][1]];
            }
        L.smem = (size_t)(a.maxlen + 1) * 128 * sizeof(double);
        // staged variant: + u32 [4 warps][4 planes][maxrows][32] for the slice's gather-map entries
        static const int stage_env = getenv("FEMB_P1_STAGE") ? atoi(getenv("FEMB_P1_STAGE")) : 1;
        a.maxrows = (stage_env && ctx->gmap_maxrows > 0 && ctx->gmap_maxrows <= 12) ? ctx->gmap_maxrows : 0;
        L.smem += (size_t)a.maxrows * 4 * 4 * 32 * sizeof(uint32_t);
    }
    return FEMB_OK;
}

// launch the P1 gather kernel for slices [s0, s1) on `stream`
static int launch_p1_slices(femb_ctx* ctx, P1Launch& L, int64_t s0, int64_t s1, cudaStream_t stream) {
    if (s1 <= s0) return FEMB_OK;
    P1Args& a = L.a;
    a.slice_begin = s0;
    a.slice_end = s1;
    const unsigned grid = (unsigned)((s1 - s0 + 3) / 4);
    const size_t smem = L.smem;
    if (ctx->dim == 2) {
        const P1Tab2& t2 = L.t2;
        const RhsDev& rhs = L.rhs;
        const bool track = a.touched != nullptr, hv = a.cellvol != nullptr;
#define FEMB_P1_LAUNCH2S(T, R, V, Q, S)                                                                                                  \
    do {                                                                                                                                 \
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_p1_poisson_gather2d<T, R, V, Q, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_p1_poisson_gather2d<T, R, V, Q, S><<<grid, 128, smem, stream>>>(a, rhs, t2);                                                    \
    } while (0)
#define FEMB_P1_LAUNCH2Q(T, R, V, Q)                     \
    do {                                                 \
        if (a.maxrows > 0) FEMB_P1_LAUNCH2S(T, R, V, Q, true); \
        else FEMB_P1_LAUNCH2S(T, R, V, Q, false);        \
    } while (0)
#define FEMB_P1_LAUNCH2(T, R, V)                 \
    do {                                         \
        if (t2.nqp == 1) FEMB_P1_LAUNCH2Q(T, R, V, 1); \
        else FEMB_P1_LAUNCH2Q(T, R, V, 0);       \
    } while (0)
#define FEMB_P1_LAUNCH2_R(R)                                    \
    do {                                                        \
        if (track && hv) FEMB_P1_LAUNCH2(true, R, true);        \
        else if (track) FEMB_P1_LAUNCH2(true, R, false);        \
        else if (hv) FEMB_P1_LAUNCH2(false, R, true);           \
        else FEMB_P1_LAUNCH2(false, R, false);                  \
    } while (0)
        if (rhs.kind == FEMB_RHS_AFFINE) FEMB_P1_LAUNCH2_R(FEMB_RHS_AFFINE);
        else if (rhs.kind == FEMB_RHS_QPVALUES) FEMB_P1_LAUNCH2_R(FEMB_RHS_QPVALUES);
        else FEMB_P1_LAUNCH2_R(FEMB_RHS_NONE);
#undef FEMB_P1_LAUNCH2_R
#undef FEMB_P1_LAUNCH2
#undef FEMB_P1_LAUNCH2Q
#undef FEMB_P1_LAUNCH2S
        KERNEL_CHECK(ctx);
        return FEMB_OK;
    }
    return launch_p1_gather_dim<3>(ctx, a, L.rhs, L.tab, grid, smem, stream);
}

static int launch_p1_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol) {
    P1Launch L;
    FEMB_TRY(prepare_p1_gather(ctx, eg, ei, mu, rhs, drop_tol, L));
    if (ctx->fused_exchange && ctx->comm && !ctx->peers.empty() && ctx->stream_x) {
        // overlapped interface exchange: assemble the slices holding ghost (send) and owned-interface (recv) columns
        // first, run pack / NCCL / add on a second stream while the bulk of the columns is assembled
        int64_t r[2][2] = {{ctx->xdof_send_lo / 32, ctx->xdof_send_hi / 32 + 1}, {ctx->xdof_recv_lo / 32, ctx->xdof_recv_hi / 32 + 1}};
        if (ctx->xdof_send_hi < 0) r[0][0] = r[0][1] = 0;
        if (ctx->xdof_recv_hi < 0) r[1][0] = r[1][1] = 0;
        for (auto& q : r) {  // CTA granularity (4 slices)
            q[0] = (q[0] / 4) * 4;
            q[1] = std::min<int64_t>(((q[1] + 3) / 4) * 4, L.nslices);
        }
        if (r[1][1] > r[1][0] && (r[0][1] <= r[0][0] || r[1][0] < r[0][0])) std::swap(r[0], r[1]);  // ascending
        if (r[0][1] > r[1][0] && r[1][1] > r[1][0]) {  // overlapping: merge
            r[0][1] = std::max(r[0][1], r[1][1]);
            r[1][0] = r[1][1] = 0;
        }
        const int64_t nb = (r[0][1] - r[0][0]) + (r[1][1] - r[1][0]);
        if (nb > 0 && nb * 8 <= L.nslices) {
            // second (high-priority) stream: boundary slices -> pack / NCCL / add; main stream: all other slices, concurrently
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_x0, ctx->stream));  // everything queued so far on the main stream comes first
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_x, ctx->ev_x0, 0));
            for (auto& q : r) FEMB_TRY(launch_p1_slices(ctx, L, q[0], q[1], ctx->stream_x));
            FEMB_TRY(femb_exchange_on(ctx, ctx->stream_x));
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_x1, ctx->stream_x));
            int64_t pos = 0;  // the complement of the boundary ranges
            for (auto& q : r) {
                if (q[1] <= q[0]) continue;
                FEMB_TRY(launch_p1_slices(ctx, L, pos, q[0], ctx->stream));
                pos = q[1];
            }
            FEMB_TRY(launch_p1_slices(ctx, L, pos, L.nslices, ctx->stream));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_x1, 0));
            return FEMB_OK;
        }
        FEMB_TRY(launch_p1_slices(ctx, L, 0, L.nslices, ctx->stream));  // interface too large to be worth splitting
        return femb_exchange_on(ctx, ctx->stream);
    }
    return launch_p1_slices(ctx, L, 0, L.nslices, ctx->stream);
}

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
static int launch_stiffness_dmma(femb_ctx* ctx, int eg, double mu, double drop_tol) {
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
template <int DIM, int ND, int NQ, int SPLIT>
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
        const double* cc = a.coef + (size_t)cell1 * FEMB_COEF_STRIDE;
#pragma unroll
        for (int i = 0; i < KP; i++) kc1[i] = __ldg(cc + i);
#pragma unroll
        for (int qp = 0; qp < NQ; qp++) f1[qp] = a.has_rhs ? __ldg(cc + 6 + qp) : 0.0;
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
        {
            const double* cc = a.coef + (size_t)cell1 * FEMB_COEF_STRIDE;
            if (DIM == 2) {
                const double2 c01 = __ldg(reinterpret_cast<const double2*>(cc));
                kc1[0] = c01.x; kc1[1] = c01.y; kc1[2] = __ldg(cc + 2);
            } else {
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double2 cv = __ldg(reinterpret_cast<const double2*>(cc) + i);
                    kc1[2 * i] = cv.x; kc1[2 * i + 1] = cv.y;
                }
            }
            if (a.has_rhs) {
#pragma unroll
                for (int qp = 0; qp < NQ; qp++) f1[qp] = __ldg(cc + 6 + qp);
            }
        }
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
            if (on && (SPLIT == 1 || (int)warp == ph)) {
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
    auto kern1 = k_h1_gather<DIM, ND, NQ, 1>;
    auto kern4 = k_h1_gather<DIM, ND, NQ, 4>;
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
static int launch_h1_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol) {
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

// ---------------------------------------------------------------------------------------------------
// Example210 convection (Newton linearisation), 2-D, two components   (Example210:322-365)
// ---------------------------------------------------------------------------------------------------
struct ConvArgs {
    const double* coords;
    const int32_t* cellnodes;
    const double* cellvol;
    int64_t ncells;
    SpaceDev us;
    int nqp, cpb;
    const double *rv, *rd;  // Identity / Gradient tables of the velocity space
    double w[FEMB_MAX_QP];
    const double* sol;
    const int32_t* slots;  // (u,u) block
    double* nzval;
    uint8_t* touched;
    int32_t* errflag;
    double* b;  // velocity part of the rhs
};

__global__ void __launch_bounds__(128) k_convection2d(ConvArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Geo<2>* geo = reinterpret_cast<Geo<2>*>(smem_raw);
    const int nd = a.us.nd, nq = a.nqp;
    double* G = reinterpret_cast<double*>(geo + a.cpb);  // [cl][qp][4][nd]
    double* I = G + (size_t)a.cpb * nq * 4 * nd;         // [cl][qp][2][nd]
    double* T = I + (size_t)a.cpb * nq * 2 * nd;         // [cl][qp][2][nd]   tempV_j
    double* IN = T + (size_t)a.cpb * nq * 2 * nd;        // [cl][qp][8]       input(6), rhs tempV(2)
    const int64_t base = (int64_t)blockIdx.x * a.cpb;
    const int ncl = (int)min((int64_t)a.cpb, a.ncells - base);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int cl = tid; cl < ncl; cl += nt) load_geo<2>(a.coords, a.cellnodes, a.cellvol, base + cl, geo[cl]);
    __syncthreads();
    for (int i = tid; i < ncl * nq * nd; i += nt) {
        int cl = i / (nq * nd), rem = i - cl * (nq * nd);
        int qp = rem / nd, dof = rem - qp * nd;
        double v[4];
        eval_op<2>(a.us, FEMB_OP_GRADIENT, geo[cl], base + cl, dof, qp, a.rv, a.rd, v);
        for (int r = 0; r < 4; r++) G[((size_t)(cl * nq + qp) * 4 + r) * nd + dof] = v[r];
        eval_op<2>(a.us, FEMB_OP_IDENTITY, geo[cl], base + cl, dof, qp, a.rv, a.rd, v);
        for (int r = 0; r < 2; r++) I[((size_t)(cl * nq + qp) * 2 + r) * nd + dof] = v[r];
    }
    __syncthreads();
    // input = (u_h, grad u_h) at every (cell, qp)   (Example210:324-333)
    for (int i = tid; i < ncl * nq; i += nt) {
        int cl = i / nq, qp = i - cl * nq;
        double in[6] = {0, 0, 0, 0, 0, 0};
        for (int j = 0; j < nd; j++) {
            double sj = a.sol[a.us.celldofs[(base + cl) * nd + j] - 1];
            for (int d = 0; d < 2; d++) in[d] += sj * I[((size_t)i * 2 + d) * nd + j];
            for (int d = 0; d < 4; d++) in[2 + d] += sj * G[((size_t)i * 4 + d) * nd + j];
        }
        // value = (u.grad)u ; jac = [[g11,g12,u1,u2,0,0],[g21,g22,0,0,u1,u2]] ; tempV = jac*input - value
        double val0 = in[0] * in[2] + in[1] * in[3];
        double val1 = in[0] * in[4] + in[1] * in[5];
        double t0 = in[2] * in[0] + in[3] * in[1] + in[0] * in[2] + in[1] * in[3];
        double t1 = in[4] * in[0] + in[5] * in[1] + in[0] * in[4] + in[1] * in[5];
        double* o = IN + (size_t)i * 8;
        for (int d = 0; d < 6; d++) o[d] = in[d];
        o[6] = t0 - val0;
        o[7] = t1 - val1;
    }
    __syncthreads();
    // tempV_j = jac * [phi_j; grad phi_j]   (Example210:339-350)
    for (int i = tid; i < ncl * nq * nd; i += nt) {
        int cq = i / nd, j = i - cq * nd;
        const double* in = IN + (size_t)cq * 8;
        const double i0 = I[((size_t)cq * 2 + 0) * nd + j], i1 = I[((size_t)cq * 2 + 1) * nd + j];
        const double g0 = G[((size_t)cq * 4 + 0) * nd + j], g1 = G[((size_t)cq * 4 + 1) * nd + j];
        const double g2 = G[((size_t)cq * 4 + 2) * nd + j], g3 = G[((size_t)cq * 4 + 3) * nd + j];
        double t0 = 0.0, t1 = 0.0;
        t0 += in[2] * i0; t1 += in[4] * i0;
        t0 += in[3] * i1; t1 += in[5] * i1;
        t0 += in[0] * g0; t1 += 0.0 * g0;
        t0 += in[1] * g1; t1 += 0.0 * g1;
        t0 += 0.0 * g2;   t1 += in[0] * g2;
        t0 += 0.0 * g3;   t1 += in[1] * g3;
        T[((size_t)cq * 2 + 0) * nd + j] = t0;
        T[((size_t)cq * 2 + 1) * nd + j] = t1;
    }
    __syncthreads();
    const int per = nd * nd;
    for (int i = tid; i < ncl * per; i += nt) {
        int cl = i / per, rc = i - cl * per;
        int r = rc / nd, c = rc - r * nd;  // row = test function k, column = ansatz j
        double s = 0.0;
        for (int qp = 0; qp < nq; qp++) {
            size_t cq = (size_t)cl * nq + qp;
            double d = T[(cq * 2 + 0) * nd + c] * I[(cq * 2 + 0) * nd + r] + T[(cq * 2 + 1) * nd + c] * I[(cq * 2 + 1) * nd + r];
            s += d * a.w[qp];
        }
        scatter_add(a.nzval, a.touched, a.errflag, a.slots[(base + cl) * per + rc], s * geo[cl].vol, 1);
    }
    for (int i = tid; i < ncl * nd; i += nt) {
        int cl = i / nd, j = i - cl * nd;
        double vol = geo[cl].vol;
        double s = 0.0;
        for (int qp = 0; qp < nq; qp++) {
            size_t cq = (size_t)cl * nq + qp;
            const double* in = IN + cq * 8;
            s += (in[6] * I[(cq * 2 + 0) * nd + j] + in[7] * I[(cq * 2 + 1) * nd + j]) * a.w[qp] * vol;
        }
        atomicAdd(a.b + a.us.celldofs[(base + cl) * nd + j] - 1, s);
    }
}

// Register-blocked version for component-blocked H1 velocity spaces without per-cell handlers (H1Pk{2,2,2}, H1P2{2,2}):
// local dof (c, i) = e_c psi_i with the scalar basis psi.  With u_h, g[r][c] = d_c u_r at the quadrature point and
// adv_j = u_h . grad psi_j the Newton terms of Example210:339-355 collapse to
//     Aloc[(ck,k),(cj,j)] += w psi_k ( g[ck][cj] psi_j + delta(ck,cj) adv_j ),
//     b[(c,j)]            += w |T| psi_j ( (jac . input)_c - value_c )          (Example210:358-363)
// One thread per (cell, 6x6 block): 36 accumulators in registers, no shared-memory staging of cvals.
struct Conv2Args {
    const double* coords;
    const int32_t* cellnodes;
    const double* cellvol;
    int64_t ncells;
    const int32_t* celldofs;  // [cell][2*NDS]
    int nqp;
    const double* rv;         // Identity table [qp][2*NDS][2]
    const double* rd;         // Gradient table [qp][2*NDS][2][2]
    double w[FEMB_MAX_QP];
    const double* sol;
    const int32_t* slots;     // (u,u) block [cell][2*NDS][2*NDS]
    double* nzval;
    uint8_t* touched;
    int32_t* errflag;
    double* b;
};

template <int NDS>
__global__ void __launch_bounds__(128) k_convection2d_blocks(const Conv2Args a) {
    __shared__ double tab[FEMB_MAX_QP][NDS][3];  // psi, dpsi/dxref_0, dpsi/dxref_1
    for (int i = threadIdx.x; i < a.nqp * NDS; i += blockDim.x) {
        const int q = i / NDS, d = i - q * NDS;
        tab[q][d][0] = __ldg(a.rv + ((size_t)q * 2 * NDS + d) * 2);
        tab[q][d][1] = __ldg(a.rd + (((size_t)q * 2 * NDS + d) * 2) * 2);
        tab[q][d][2] = __ldg(a.rd + (((size_t)q * 2 * NDS + d) * 2) * 2 + 1);
    }
    __syncthreads();
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t cell = t >> 2;
    if (cell >= a.ncells) return;
    const int ck = (int)(t & 3) >> 1, cj = (int)(t & 1);
    double x[3][2];
    load_nodes<2>(a.coords, a.cellnodes, cell, x);
    const double A00 = x[1][0] - x[0][0], A01 = x[2][0] - x[0][0], A10 = x[1][1] - x[0][1], A11 = x[2][1] - x[0][1];
    const double det = A00 * A11 - A01 * A10;
    const double M00 = A11 / det, M01 = -A10 / det, M10 = -A01 / det, M11 = A00 / det;  // mapderiv!: A^{-T}
    const double vol = a.cellvol ? __ldg(a.cellvol + cell) : fabs(det) * 0.5;
    const int32_t* cd = a.celldofs + cell * (2 * NDS);
    double su[2][NDS];
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int i = 0; i < NDS; i++) su[c][i] = __ldg(a.sol + (__ldg(cd + c * NDS + i) - 1));
    double acc[NDS][NDS], rb[NDS];
#pragma unroll
    for (int k = 0; k < NDS; k++) {
        rb[k] = 0.0;
#pragma unroll
        for (int j = 0; j < NDS; j++) acc[k][j] = 0.0;
    }
    for (int q = 0; q < a.nqp; q++) {
        double psi[NDS], adv[NDS];
        double u0 = 0.0, u1 = 0.0, g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0;
#pragma unroll
        for (int i = 0; i < NDS; i++) {  // input = (u_h, grad u_h)   (Example210:324-333)
            psi[i] = tab[q][i][0];
            const double gx = M00 * tab[q][i][1] + M01 * tab[q][i][2], gy = M10 * tab[q][i][1] + M11 * tab[q][i][2];
            u0 += su[0][i] * psi[i];
            u1 += su[1][i] * psi[i];
            g00 += su[0][i] * gx; g01 += su[0][i] * gy;
            g10 += su[1][i] * gx; g11 += su[1][i] * gy;
            adv[i] = gx;  // completed below once u_h is known
        }
#pragma unroll
        for (int i = 0; i < NDS; i++) {
            const double gy = M10 * tab[q][i][1] + M11 * tab[q][i][2];
            adv[i] = u0 * adv[i] + u1 * gy;
        }
        const double gkj = ck == 0 ? (cj == 0 ? g00 : g01) : (cj == 0 ? g10 : g11);
        const double wq = a.w[q];
#pragma unroll
        for (int k = 0; k < NDS; k++) {
            const double pk = psi[k] * wq;
#pragma unroll
            for (int j = 0; j < NDS; j++) {
                const double tv = (ck == cj) ? gkj * psi[j] + adv[j] : gkj * psi[j];
                acc[k][j] += tv * pk;
            }
        }
        if (cj == 0) {  // rhs of component ck: jac . input - value
            const double gr0 = ck == 0 ? g00 : g10, gr1 = ck == 0 ? g01 : g11;
            const double value = u0 * gr0 + u1 * gr1;
            const double ji = gr0 * u0 + gr1 * u1 + u0 * gr0 + u1 * gr1;
            const double tv = ji - value;
#pragma unroll
            for (int j = 0; j < NDS; j++) rb[j] += tv * psi[j] * wq * vol;
        }
    }
    const int nd = 2 * NDS;
    const int32_t* sl = a.slots + cell * (nd * nd) + (ck * NDS) * nd + cj * NDS;
#pragma unroll
    for (int k = 0; k < NDS; k++) {
        int32_t sk[NDS];
#pragma unroll
        for (int j = 0; j < NDS; j++) sk[j] = __ldg(sl + k * nd + j);
#pragma unroll
        for (int j = 0; j < NDS; j++) scatter_add(a.nzval, a.touched, a.errflag, sk[j], acc[k][j] * vol, 1);
    }
    if (cj == 0) {
#pragma unroll
        for (int j = 0; j < NDS; j++) atomicAdd(a.b + (__ldg(cd + ck * NDS + j) - 1), rb[j]);
    }
}

int femb_build_coloring(femb_ctx* ctx) {
    return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "FEMB_SCATTER_COLORED is not built yet; use FEMB_SCATTER_ATOMIC or FEMB_SCATTER_GATHER");
}

// ---------------------------------------------------------------------------------------------------
// C ABI entry points
// ---------------------------------------------------------------------------------------------------
static int check_eval(femb_ctx* ctx, int id) {
    if (id < 0 || id >= FEMB_MAX_EVALS || !ctx->evals[id].set) return femb_fail(ctx, FEMB_ERR_ARG, "unknown evaluator %d", id);
    return FEMB_OK;
}

static int require_pattern(femb_ctx* ctx) {
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
    FEMB_TRY(femb_timer_begin(ctx));
    FEMB_TRY(launch_bilinear(ctx, er, ec, ctx->evals[er].op, ctx->evals[ec].op, brow, bcol, factor, flags, drop_tol));
    return femb_timer_end(ctx);
}

extern "C" int32_t femb_assemble_linear(femb_ctx* ctx, int32_t ev, int32_t brow, double factor, int32_t rhs_kind, const double* data) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(check_eval(ctx, ev));
    if (!ctx->bvec) return femb_fail(ctx, FEMB_ERR_ARG, "femb_matrix_layout must precede femb_assemble_linear");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RhsDev rhs;
    double* dev_f = nullptr;
    FEMB_TRY(make_rhs(ctx, rhs_kind, data, ctx->evals[ev].resultdim, ctx->evals[ev].nqp, rhs, &dev_f));
    int st = femb_timer_begin(ctx);
    if (st == FEMB_OK && rhs_kind != FEMB_RHS_NONE) st = launch_linear(ctx, ev, brow, factor, rhs);
    if (st == FEMB_OK) st = femb_timer_end(ctx);
    cudaFree(dev_f);
    return st;
}

extern "C" int32_t femb_assemble_poisson(femb_ctx* ctx, int32_t eg, int32_t ei, double mu, int32_t rhs_kind, const double* data, double drop_tol) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(check_eval(ctx, eg));
    if (rhs_kind != FEMB_RHS_NONE) FEMB_TRY(check_eval(ctx, ei));
    FEMB_TRY(require_pattern(ctx));
    const FembEval& EG = ctx->evals[eg];
    if (EG.op != FEMB_OP_GRADIENT) return femb_fail(ctx, FEMB_ERR_ARG, "eval_grad must be a Gradient evaluator");
    if (rhs_kind != FEMB_RHS_NONE && (ctx->evals[ei].op != FEMB_OP_IDENTITY || ctx->evals[ei].space != EG.space || ctx->evals[ei].nqp != EG.nqp))
        return femb_fail(ctx, FEMB_ERR_ARG, "eval_id must be the Identity evaluator of the same space and rule");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RhsDev rhs;
    double* dev_f = nullptr;
    FEMB_TRY(make_rhs(ctx, rhs_kind, data, ctx->spaces[EG.space].ncomp, EG.nqp, rhs, &dev_f));
    int st = femb_timer_begin(ctx);
    if (st == FEMB_OK) {
        if (ctx->scatter == FEMB_SCATTER_GATHER && ctx->gmap && ctx->gmap_space == EG.space && (!ctx->touched || ctx->max_collen <= 64) &&
            rhs_kind != FEMB_RHS_EX210 && (ctx->dim == 3 || EG.nqp <= FEMB_P1_MAXQP)) {
            st = launch_p1_gather(ctx, eg, ei, mu, rhs, drop_tol);
        } else if (ctx->scatter == FEMB_SCATTER_GATHER && rhs_kind != FEMB_RHS_EX210 && (st = launch_h1_gather(ctx, eg, ei, mu, rhs, drop_tol)) != FEMB_ERR_UNSUPPORTED) {
            // owner-computes gather for general scalar H1 elements (stiffness + rhs in one pass, write-once)
        } else {
            st = (ctx->scatter == FEMB_SCATTER_COLORED) ? FEMB_ERR_UNSUPPORTED : launch_stiffness_dmma(ctx, eg, mu, drop_tol);
            if (st == FEMB_ERR_UNSUPPORTED)  // elements with per-cell basis subsets, vector-valued spaces, ...: generic quadrature kernel
                st = launch_bilinear(ctx, eg, eg, FEMB_OP_GRADIENT, FEMB_OP_GRADIENT, 0, 0, mu, FEMB_BIL_SYMMETRIC_UPPER, drop_tol);
            if (st == FEMB_OK && rhs_kind != FEMB_RHS_NONE) st = launch_linear(ctx, ei, 0, 1.0, rhs);
        }
    }
    if (st == FEMB_OK) st = femb_timer_end(ctx);
    cudaFree(dev_f);
    return st;
}

// ---------------------------------------------------------------------------------------------------
// streamed host-to-host P1 assembly: upload / kernel / download overlap over column chunks
// ---------------------------------------------------------------------------------------------------
// per chunk: largest node id and cell id referenced by the gather-map rows of the chunk's slices
__global__ void k_chunk_ranges(const uint32_t* __restrict__ gmap, size_t stride, const int32_t* __restrict__ slice_ptr, const int64_t* __restrict__ chunk_slice,
                               int nchunks, int planes_nc, unsigned long long* __restrict__ nmax, unsigned long long* __restrict__ cmax) {
    const int k = blockIdx.x;
    const size_t i0 = (size_t)slice_ptr[chunk_slice[k]] << 5, i1 = (size_t)slice_ptr[chunk_slice[k + 1]] << 5;
    unsigned long long nm = 0ull, cm = 0ull;
    for (size_t i = i0 + blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < i1; i += (size_t)gridDim.y * blockDim.x) {
        if (gmap[2 * stride + i] == 0xFFFFFFFFu) continue;
        unsigned long long a = gmap[i], b = gmap[stride + i];
        nm = max(nm, max(a, b));
        if (planes_nc) nm = max(nm, (unsigned long long)gmap[4 * stride + i]);
        cm = max(cm, (unsigned long long)gmap[3 * stride + i]);
    }
    atomicMax(nmax + k, nm);
    atomicMax(cmax + k, cm);
}

static int plan_chunks(femb_ctx* ctx, int nchunks, int64_t nslices, int64_t ndofs) {
    if (ctx->chunk_n == nchunks) return FEMB_OK;
    ctx->chunk_n = 0;
    if (!ctx->stream_h2d) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_h2d, cudaStreamNonBlocking));
    if (!ctx->stream_d2h) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_d2h, cudaStreamNonBlocking));
    while ((int)ctx->chunk_ev.size() < 2 * nchunks) {
        cudaEvent_t e;
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_ev.push_back(e);
    }
    ctx->chunk_slice.assign(nchunks + 1, 0);
    for (int k = 0; k <= nchunks; k++) {
        int64_t s = nslices * k / nchunks;
        ctx->chunk_slice[k] = (k == nchunks) ? nslices : (s / 4) * 4;  // CTA = 4 slices
    }
    int64_t* d_slice = nullptr;
    unsigned long long* d_max = nullptr;
    FEMB_TRY(femb_alloc(ctx, &d_slice, (size_t)nchunks + 1));
    int st = femb_alloc(ctx, &d_max, (size_t)2 * nchunks);
    if (st != FEMB_OK) { cudaFree(d_slice); return st; }
    cudaMemcpyAsync(d_slice, ctx->chunk_slice.data(), sizeof(int64_t) * (nchunks + 1), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(d_max, 0, sizeof(unsigned long long) * 2 * nchunks, ctx->stream);
    k_chunk_ranges<<<dim3(nchunks, 64), 256, 0, ctx->stream>>>(ctx->gmap, ctx->gmap_stride, ctx->gslice_ptr, d_slice, nchunks, ctx->dim == 3, d_max, d_max + nchunks);
    ctx->launches++;
    std::vector<unsigned long long> h(2 * nchunks);
    cudaMemcpyAsync(h.data(), d_max, sizeof(unsigned long long) * 2 * nchunks, cudaMemcpyDeviceToHost, ctx->stream);
    ctx->chunk_cp.assign(nchunks + 1, 0);
    for (int k = 0; k <= nchunks; k++) {
        int64_t d = std::min<int64_t>(ctx->chunk_slice[k] * 32, ndofs);
        cudaMemcpyAsync(&ctx->chunk_cp[k], ctx->colptr + d, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_slice);
    cudaFree(d_max);
    CUDA_TRY(ctx, e);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->chunk_nmax.assign(nchunks, 0);
    ctx->chunk_cmax.assign(nchunks, 0);
    for (int k = 0; k < nchunks; k++) {
        // the chunk's own dofs are nodes too (their coordinates are read as `self`)
        int64_t dlast = std::min<int64_t>(ctx->chunk_slice[k + 1] * 32, ndofs) - 1;
        ctx->chunk_nmax[k] = std::max<int64_t>((int64_t)h[k], dlast);
        ctx->chunk_cmax[k] = (int64_t)h[nchunks + k];
        ctx->chunk_cp[k] -= 1;
    }
    ctx->chunk_cp[nchunks] -= 1;
    ctx->chunk_n = nchunks;
    return FEMB_OK;
}

extern "C" int32_t femb_assemble_poisson_host(femb_ctx* ctx, int32_t eg, int32_t ei, double mu, int32_t rhs_kind, const double* params, double drop_tol,
                                              const double* coords, const int32_t* cellnodes, const double* cellvol, double* nzval_out, double* b_out,
                                              int32_t nchunks) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(check_eval(ctx, eg));
    if (rhs_kind != FEMB_RHS_NONE) FEMB_TRY(check_eval(ctx, ei));
    FEMB_TRY(require_pattern(ctx));
    if (!coords || !nzval_out || (rhs_kind != FEMB_RHS_NONE && !b_out)) return femb_fail(ctx, FEMB_ERR_ARG, "femb_assemble_poisson_host: NULL buffer");
    if ((cellvol != nullptr) != (ctx->cellvol != nullptr)) return femb_fail(ctx, FEMB_ERR_ARG, "cellvolumes must be given iff the registered mesh has them");
    const FembEval& EG = ctx->evals[eg];
    if (EG.op != FEMB_OP_GRADIENT) return femb_fail(ctx, FEMB_ERR_ARG, "eval_grad must be a Gradient evaluator");
    if (!(ctx->gmap && ctx->gmap_space == EG.space) || ctx->touched || (rhs_kind != FEMB_RHS_NONE && rhs_kind != FEMB_RHS_AFFINE) ||
        (ctx->dim == 2 && EG.nqp > FEMB_P1_MAXQP))
        return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "the streamed path covers scalar P1 with rhs NONE/AFFINE and no pattern tracking; use femb_mesh_set + femb_assemble_poisson + femb_matrix_get");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RhsDev rhs;
    double* dev_f = nullptr;
    FEMB_TRY(make_rhs(ctx, rhs_kind, params, 1, EG.nqp, rhs, &dev_f));
    const FembSpace& s = ctx->spaces[EG.space];
    const int64_t nslices = (s.ndofs + 31) / 32;
    nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, nslices / 4 > 0 ? nslices / 4 : 1));
    FEMB_TRY(plan_chunks(ctx, nchunks, nslices, s.ndofs));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // earlier kernels may still read the mesh arrays
    ctx->zero_pending_A = true;  // "zero + assemble": the write-once kernel overwrites
    ctx->zero_pending_b = true;
    P1Launch L;
    FEMB_TRY(prepare_p1_gather(ctx, eg, ei, mu, rhs, drop_tol, L));
    FEMB_TRY(femb_timer_begin(ctx));
    const int dim = ctx->dim;
    int64_t hi_n = 0, hi_c = 0;
    for (int k = 0; k < nchunks; k++) {
        const int64_t need_n = ctx->chunk_nmax[k] + 1, need_c = ctx->chunk_cmax[k] + 1;
        if (need_n > hi_n) {
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->coords + hi_n * dim, coords + hi_n * dim, sizeof(double) * (need_n - hi_n) * dim, cudaMemcpyHostToDevice, ctx->stream_h2d));
            hi_n = need_n;
        }
        if (cellvol && need_c > hi_c) {
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->cellvol + hi_c, cellvol + hi_c, sizeof(double) * (need_c - hi_c), cudaMemcpyHostToDevice, ctx->stream_h2d));
            hi_c = need_c;
        }
        CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[2 * k], ctx->stream_h2d));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[2 * k], 0));
        FEMB_TRY(launch_p1_slices(ctx, L, ctx->chunk_slice[k], ctx->chunk_slice[k + 1], ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[2 * k + 1], ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_d2h, ctx->chunk_ev[2 * k + 1], 0));
        const int64_t cp0 = ctx->chunk_cp[k], cp1 = ctx->chunk_cp[k + 1];
        const int64_t d0 = std::min<int64_t>(ctx->chunk_slice[k] * 32, s.ndofs), d1 = std::min<int64_t>(ctx->chunk_slice[k + 1] * 32, s.ndofs);
        if (cp1 > cp0) CUDA_TRY(ctx, cudaMemcpyAsync(nzval_out + cp0, ctx->nzval + cp0, sizeof(double) * (cp1 - cp0), cudaMemcpyDeviceToHost, ctx->stream_d2h));
        if (rhs_kind != FEMB_RHS_NONE && d1 > d0)
            CUDA_TRY(ctx, cudaMemcpyAsync(b_out + d0, ctx->bvec + d0, sizeof(double) * (d1 - d0), cudaMemcpyDeviceToHost, ctx->stream_d2h));
    }
    // the rest of the mesh: nodes / volumes no chunk referenced, and the topology (not read by the numeric phase)
    if (ctx->nnodes > hi_n)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->coords + hi_n * dim, coords + hi_n * dim, sizeof(double) * (ctx->nnodes - hi_n) * dim, cudaMemcpyHostToDevice, ctx->stream_h2d));
    if (cellvol && ctx->ncells > hi_c)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->cellvol + hi_c, cellvol + hi_c, sizeof(double) * (ctx->ncells - hi_c), cudaMemcpyHostToDevice, ctx->stream_h2d));
    if (cellnodes)  // NULL: topology unchanged since femb_mesh_set (it is what the reused pattern was built from)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->cellnodes, cellnodes, sizeof(int32_t) * ctx->ncells * ctx->nnpc, cudaMemcpyHostToDevice, ctx->stream_h2d));
    int st = femb_timer_end(ctx);
    cudaError_t e1 = cudaStreamSynchronize(ctx->stream_h2d), e2 = cudaStreamSynchronize(ctx->stream_d2h);
    cudaFree(dev_f);
    if (st != FEMB_OK) return st;
    CUDA_TRY(ctx, e1);
    CUDA_TRY(ctx, e2);
    return femb_sync(ctx);
}

extern "C" int32_t femb_assemble_navierstokes(femb_ctx* ctx, int32_t egu, int32_t eiu, int32_t eip, double mu, int32_t linear, int32_t nonlinear,
                                              int32_t rhs_kind, const double* data) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(check_eval(ctx, egu));
    FEMB_TRY(check_eval(ctx, eiu));
    FEMB_TRY(check_eval(ctx, eip));
    FEMB_TRY(require_pattern(ctx));
    const FembEval& EG = ctx->evals[egu];
    const FembEval& EI = ctx->evals[eiu];
    const FembEval& EP = ctx->evals[eip];
    const FembSpace& us = ctx->spaces[EG.space];
    if (ctx->dim != 2 || us.ncomp != 2 || EG.op != FEMB_OP_GRADIENT || EI.op != FEMB_OP_IDENTITY || EP.op != FEMB_OP_IDENTITY || EI.space != EG.space)
        return femb_fail(ctx, FEMB_ERR_ARG, "femb_assemble_navierstokes expects (Gradient u, Identity u, Identity p) of a 2-D two-component velocity space");
    if (ctx->nrs != 2 || ctx->ncs != 2 || ctx->row_spaces[0] != EG.space || ctx->row_spaces[1] != EP.space || ctx->col_spaces[0] != EG.space ||
        ctx->col_spaces[1] != EP.space)
        return femb_fail(ctx, FEMB_ERR_ARG, "layout must be FEMatrix([FES_u, FES_p])");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RhsDev rhs;
    double* dev_f = nullptr;
    FEMB_TRY(make_rhs(ctx, linear ? rhs_kind : FEMB_RHS_NONE, data, 2, EG.nqp, rhs, &dev_f));
    int st = femb_timer_begin(ctx);
    if (st == FEMB_OK && linear) {
        // Aloc[k,j] = mu * sum_qp w <grad_j, grad_k>, then * |T|                 (Example210:284-290,368)
        // component-blocked basis: the 12x12 local matrix is two copies of the scalar 6x6 P2 stiffness; its
        // off-diagonal blocks are exact zeros (their slots stay in the pattern, adding 0.0 changes nothing)
        st = launch_stiffness_dmma(ctx, egu, mu, -1.0);
        if (st == FEMB_ERR_UNSUPPORTED) st = launch_bilinear(ctx, egu, egu, FEMB_OP_GRADIENT, FEMB_OP_GRADIENT, 0, 0, mu, 0, 0.0);
        // Bloc[j,k] = -sum_qp w (g[1,j]+g[4,j]) psi_k * |T| into (u,p) and (p,u) (Example210:293-301,376-380)
        if (st == FEMB_OK) st = launch_bilinear(ctx, egu, eip, FEMB_OP_GRADTRACE, FEMB_OP_IDENTITY, 0, 1, -1.0, FEMB_BIL_TRANSPOSE_COPY, 0.0);
        if (st == FEMB_OK && rhs.kind != FEMB_RHS_NONE) st = launch_linear(ctx, eiu, 0, 1.0, rhs);
    }
    if (st == FEMB_OK && nonlinear) {
        st = femb_build_slotmaps(ctx);
        const FembBlock* blk = find_block(ctx, 0, 0);
        if (st == FEMB_OK && (!blk || !blk->slots)) st = femb_fail(ctx, FEMB_ERR_ARG, "(u,u) block missing from the pattern");
        if (st == FEMB_OK) st = femb_flush_zero(ctx, true, true);
        const int nds = us.nd / 2;
        if (st == FEMB_OK && us.handler == FEMB_HANDLER_NONE && us.nd == us.nd_all && nds == 6 && ctx->ncells) {
            Conv2Args a;
            a.coords = ctx->coords; a.cellnodes = ctx->cellnodes; a.cellvol = ctx->cellvol; a.ncells = ctx->ncells; a.celldofs = us.celldofs; a.nqp = EG.nqp;
            a.rv = EI.refvals; a.rd = EG.refderivs;
            for (int i = 0; i < EG.nqp; i++) a.w[i] = EG.w[i];
            a.sol = ctx->sol; a.slots = blk->slots; a.nzval = ctx->nzval; a.touched = ctx->touched; a.errflag = ctx->errflag; a.b = ctx->bvec;
            const int64_t nthreads = ctx->ncells * 4;
            k_convection2d_blocks<6><<<(unsigned)((nthreads + 127) / 128), 128, 0, ctx->stream>>>(a);
            ctx->launches++;
            ctx->last_launches++;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) st = femb_fail(ctx, FEMB_ERR_CUDA, "k_convection2d_blocks: %s", cudaGetErrorString(e));
        } else if (st == FEMB_OK) {
            ConvArgs a;
            a.coords = ctx->coords; a.cellnodes = ctx->cellnodes; a.cellvol = ctx->cellvol; a.ncells = ctx->ncells;
            a.us = space_dev(us); a.nqp = EG.nqp; a.rv = EI.refvals; a.rd = EG.refderivs;
            for (int i = 0; i < EG.nqp; i++) a.w[i] = EG.w[i];
            a.sol = ctx->sol; a.slots = blk->slots; a.nzval = ctx->nzval; a.touched = ctx->touched; a.errflag = ctx->errflag; a.b = ctx->bvec;
            size_t per_cell = ((size_t)EG.nqp * 8 * us.nd + (size_t)EG.nqp * 8) * sizeof(double) + sizeof(Geo<2>);
            int cpb = (int)std::min<size_t>(16, (160 * 1024) / per_cell);
            if (cpb < 1) st = femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "convection tables too large for shared memory");
            if (st == FEMB_OK) {
                a.cpb = cpb;
                size_t smem = cpb * per_cell;
                cudaError_t e = cudaFuncSetAttribute(k_convection2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) st = femb_fail(ctx, FEMB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                if (st == FEMB_OK && ctx->ncells) {
                    k_convection2d<<<(unsigned)((ctx->ncells + cpb - 1) / cpb), 128, smem, ctx->stream>>>(a);
                    ctx->launches++;
                    ctx->last_launches++;
                    e = cudaGetLastError();
                    if (e != cudaSuccess) st = femb_fail(ctx, FEMB_ERR_CUDA, "k_convection2d: %s", cudaGetErrorString(e));
                }
            }
        }
    }
    if (st == FEMB_OK) st = femb_timer_end(ctx);
    cudaFree(dev_f);
    return st;
}
