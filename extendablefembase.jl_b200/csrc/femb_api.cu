// libfemb C ABI: context lifetime, mesh / space / quadrature / evaluator registration, data movement.
#include <stdarg.h>
#include <string.h>

#include "femb_device.cuh"
#include "femb_internal.h"

int femb_fail(femb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

static thread_local std::string g_err;

extern "C" {

int32_t femb_version(void) { return 100; }

const char* femb_last_error(const femb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int32_t femb_create(int32_t device, femb_ctx** out) {
    if (!out) return FEMB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_err = std::string("no CUDA device usable (libfemb has no CPU fallback): ") + cudaGetErrorString(e);
        return FEMB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        g_err = "device index out of range";
        return FEMB_ERR_ARG;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_err = cudaGetErrorString(e);
        return FEMB_ERR_CUDA;
    }
    if (prop.major < 10) {
        g_err = "libfemb is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        return FEMB_ERR_CUDA;
    }
    femb_ctx* ctx = new femb_ctx();
    femb_reset_block_counts(ctx);
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev2)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev3)) != cudaSuccess ||
        (e = cudaMalloc((void**)&ctx->errflag, 4 * sizeof(int32_t))) != cudaSuccess ||  // [0] missing slots, [1] femb_dirichlet_apply, [2] connectivity compare
        (e = cudaMemset(ctx->errflag, 0, 4 * sizeof(int32_t))) != cudaSuccess) {
        g_err = cudaGetErrorString(e);
        delete ctx;
        return FEMB_ERR_CUDA;
    }
    *out = ctx;
    return FEMB_OK;
}

extern "C++" void femb_free_space(FembSpace& s) {
    if (s.celldofs_owned) femb_free(&s.celldofs);
    s.celldofs = nullptr;
    femb_free(&s.signs);
    femb_free(&s.orient);
    femb_free(&s.adj_ptr);
    femb_free(&s.adj);
    s.set = false;
}

int32_t femb_destroy(femb_ctx* ctx) {
    FEMB_CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    femb_comm_destroy(ctx);
    femb_free_pattern(ctx);
    for (auto& s : ctx->spaces) femb_free_space(s);
    for (auto& e : ctx->evals) {
        femb_free(&e.refvals);
        femb_free(&e.refderivs);
        femb_free(&e.reftensor);
    }
    femb_free_entities(ctx);
    femb_free_system(ctx);
    femb_free(&ctx->coords);
    femb_free(&ctx->cellnodes);
    femb_free(&ctx->cellnodes_tmp);
    femb_free(&ctx->cellvol);
    femb_free(&ctx->cellrhs);
    femb_free(&ctx->bvec);
    femb_free(&ctx->sol);
    femb_free(&ctx->errflag);
    femb_free(&ctx->color_perm);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    for (auto e : ctx->chunk_ev) cudaEventDestroy(e);
    if (ctx->stream_aux) cudaStreamDestroy(ctx->stream_aux);
    if (ctx->ev_a0) cudaEventDestroy(ctx->ev_a0);
    if (ctx->ev_a1) cudaEventDestroy(ctx->ev_a1);
    if (ctx->stream_h2d) cudaStreamDestroy(ctx->stream_h2d);
    if (ctx->stream_d2h) cudaStreamDestroy(ctx->stream_d2h);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return FEMB_OK;
}

int32_t femb_sync(femb_ctx* ctx) {
    FEMB_CHECK_CTX(ctx);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int32_t flag = 0;
    CUDA_TRY(ctx, cudaMemcpy(&flag, ctx->errflag, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(ctx->errflag, 0, sizeof(flag));
        return femb_fail(ctx, FEMB_ERR_PATTERN, "%d contribution(s) above drop_tol had no slot in the sparsity pattern", flag);
    }
    return FEMB_OK;
}

int32_t femb_host_register(femb_ctx* ctx, void* ptr, int64_t bytes) {
    FEMB_CHECK_CTX(ctx);
    if (!ptr || bytes <= 0) return femb_fail(ctx, FEMB_ERR_ARG, "femb_host_register: bad range");
    CUDA_TRY(ctx, cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return FEMB_OK;
}

int32_t femb_host_unregister(femb_ctx* ctx, void* ptr) {
    FEMB_CHECK_CTX(ctx);
    CUDA_TRY(ctx, cudaHostUnregister(ptr));
    return FEMB_OK;
}

}  // extern "C" (reopened below)

// flag = 1 if the two index arrays differ anywhere (grid-stride, one atomic per CTA that saw a difference)
__global__ void k_differs_i32(const int32_t* __restrict__ a, const int32_t* __restrict__ b, size_t n, int32_t* __restrict__ flag) {
    bool d = false;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d |= a[i] != b[i];
    if (__syncthreads_or(d) && threadIdx.x == 0) atomicOr(flag, 1);
}

// uploads `cellnodes` into the scratch copy on `stream` and compares it with the registered connectivity there; the
// result is read with femb_cellnodes_differ after the stream has been synchronised
int femb_cellnodes_compare_async(femb_ctx* ctx, const int32_t* cellnodes, cudaStream_t stream) {
    const size_t n = (size_t)ctx->ncells * ctx->nnpc;
    if (!ctx->cellnodes_tmp) FEMB_TRY(femb_alloc(ctx, &ctx->cellnodes_tmp, n));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->errflag + 2, 0, sizeof(int32_t), stream));
    if (n == 0) return FEMB_OK;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->cellnodes_tmp, cellnodes, sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream));
    k_differs_i32<<<(unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16), 256, 0, stream>>>(ctx->cellnodes, ctx->cellnodes_tmp, n, ctx->errflag + 2);
    KERNEL_CHECK(ctx);
    return FEMB_OK;
}

int femb_cellnodes_differ(femb_ctx* ctx, bool* differ) {
    int32_t f = 0;
    CUDA_TRY(ctx, cudaMemcpy(&f, ctx->errflag + 2, sizeof(f), cudaMemcpyDeviceToHost));
    *differ = f != 0;
    return FEMB_OK;
}

extern "C" {

int32_t femb_mesh_set(femb_ctx* ctx, int32_t dim, int64_t nnodes, const double* coords, int32_t nnpc, int64_t ncells,
                      const int32_t* cellnodes, const double* cellvolumes) {
    FEMB_CHECK_CTX(ctx);
    if ((dim != 2 && dim != 3) || nnpc != dim + 1)
        return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "only affine simplex cells (Triangle2D, Tetrahedron3D) are supported");
    if (!coords || !cellnodes || nnodes <= 0 || ncells < 0) return femb_fail(ctx, FEMB_ERR_ARG, "femb_mesh_set: bad arguments");
    if (ncells * (int64_t)32 >= ((int64_t)1 << 32)) return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "ncells too large for 32-bit adjacency packing");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    bool same_shape = (ctx->dim == dim && ctx->nnodes == nnodes && ctx->ncells == ncells && ctx->cellnodes);
    bool same_topology = false;
    if (same_shape) {
        // equal counts do not mean equal connectivity (edge flips, renumbering): everything derived from CellNodes — the
        // pattern, the gather maps, the dof adjacency, the entities — is kept only if the new array is identical
        FEMB_TRY(femb_cellnodes_compare_async(ctx, cellnodes, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        bool differ = true;
        FEMB_TRY(femb_cellnodes_differ(ctx, &differ));
        same_topology = !differ;
    }
    if (!same_topology) {
        // a new mesh invalidates everything derived from it
        femb_free_pattern(ctx);
        for (auto& s : ctx->spaces) femb_free_space(s);
        femb_free_entities(ctx);
        if (same_shape) {  // the scratch copy already holds the new connectivity
            std::swap(ctx->cellnodes, ctx->cellnodes_tmp);
        } else {
            FEMB_TRY(femb_alloc(ctx, &ctx->coords, (size_t)nnodes * dim));
            FEMB_TRY(femb_alloc(ctx, &ctx->cellnodes, (size_t)ncells * nnpc));
            femb_free(&ctx->cellvol);
        }
    }
    femb_free(&ctx->cellnodes_tmp);
    ctx->dim = dim;
    ctx->nnpc = nnpc;
    ctx->nnodes = nnodes;
    ctx->ncells = ncells;
    FEMB_TRY(femb_h2d(ctx, ctx->coords, coords, (size_t)nnodes * dim));
    if (!same_shape) FEMB_TRY(femb_h2d(ctx, ctx->cellnodes, cellnodes, (size_t)ncells * nnpc));
    ctx->gvol_valid = false;
    if (cellvolumes) {
        if (!ctx->cellvol) FEMB_TRY(femb_alloc(ctx, &ctx->cellvol, (size_t)ncells));
        FEMB_TRY(femb_h2d(ctx, ctx->cellvol, cellvolumes, (size_t)ncells));
    } else {
        femb_free(&ctx->cellvol);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FEMB_OK;
}

int32_t femb_space_set(femb_ctx* ctx, int32_t id, int32_t family, int32_t ncomp, int64_t ndofs, int32_t nd, int32_t nd_all,
                       int32_t handler, const int32_t* celldofs, const int32_t* signs, const int32_t* orient) {
    FEMB_CHECK_CTX(ctx);
    if (id < 0 || id >= FEMB_MAX_SPACES) return femb_fail(ctx, FEMB_ERR_ARG, "space_id out of range");
    if (ctx->ncells == 0 && !ctx->cellnodes) return femb_fail(ctx, FEMB_ERR_ARG, "femb_mesh_set must precede femb_space_set");
    if (nd <= 0 || nd > 64 || nd_all < nd || ncomp <= 0 || ncomp > 3) return femb_fail(ctx, FEMB_ERR_ARG, "bad local dof counts");
    if (!celldofs && !(nd == ctx->nnpc)) return femb_fail(ctx, FEMB_ERR_ARG, "celldofs == NULL requires nd_local == nodes per cell");
    int nsign = 0;
    switch (handler) {
        case FEMB_HANDLER_NONE: break;
        case FEMB_HANDLER_H1_FACESWAP2D: nsign = 3; break;
        case FEMB_HANDLER_H1_EDGESWAP3D: nsign = 6; break;
        case FEMB_HANDLER_RT0_SIGNS: nsign = ctx->dim + 1; break;
        case FEMB_HANDLER_BDM1_2D: nsign = 3; break;
        case FEMB_HANDLER_BDM1_3D: nsign = 4; break;
        case FEMB_HANDLER_N0_EDGESIGNS3D: nsign = 6; break;
        default: return femb_fail(ctx, FEMB_ERR_ARG, "unknown handler %d", handler);
    }
    if (handler == FEMB_HANDLER_N0_EDGESIGNS3D && ctx->dim != 3) return femb_fail(ctx, FEMB_ERR_ARG, "handler %d is the tetrahedral one (CellEdgeSigns)", handler);
    if (nsign && !signs) return femb_fail(ctx, FEMB_ERR_ARG, "handler %d needs the sign array", handler);
    if (handler == FEMB_HANDLER_BDM1_3D && !orient) return femb_fail(ctx, FEMB_ERR_ARG, "HDIVBDM1{3} needs CellFaceOrientations");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FembSpace& s = ctx->spaces[id];
    bool same = s.set && s.nd == nd && s.ndofs == ndofs && s.family == family && s.ncomp == ncomp && s.nd_all == nd_all &&
                s.handler == handler && (s.celldofs_owned == (celldofs != nullptr));
    if (!same) {
        femb_free_space(s);
        femb_free_pattern(ctx);
        if (celldofs) {
            FEMB_TRY(femb_alloc(ctx, &s.celldofs, (size_t)ctx->ncells * nd));
            s.celldofs_owned = true;
        } else {
            s.celldofs = ctx->cellnodes;
            s.celldofs_owned = false;
        }
        if (nsign) FEMB_TRY(femb_alloc(ctx, &s.signs, (size_t)ctx->ncells * nsign));
        if (orient) FEMB_TRY(femb_alloc(ctx, &s.orient, (size_t)ctx->ncells * 4));
    }
    s.family = family; s.ncomp = ncomp; s.ndofs = ndofs; s.nd = nd; s.nd_all = nd_all; s.handler = handler; s.nsign = nsign;
    if (celldofs) FEMB_TRY(femb_h2d(ctx, s.celldofs, celldofs, (size_t)ctx->ncells * nd));
    if (nsign) FEMB_TRY(femb_h2d(ctx, s.signs, signs, (size_t)ctx->ncells * nsign));
    if (orient && s.orient) FEMB_TRY(femb_h2d(ctx, s.orient, orient, (size_t)ctx->ncells * 4));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    s.set = true;
    return FEMB_OK;
}

int32_t femb_quadrature_set(femb_ctx* ctx, int32_t nqp, const double* xref, const double* w) {
    FEMB_CHECK_CTX(ctx);
    if (nqp <= 0 || nqp > FEMB_MAX_QP || !xref || !w) return femb_fail(ctx, FEMB_ERR_ARG, "nqp must be in 1..%d", FEMB_MAX_QP);
    if (!ctx->dim) return femb_fail(ctx, FEMB_ERR_ARG, "femb_mesh_set must precede femb_quadrature_set");
    ctx->nqp = nqp;
    memcpy(ctx->xref, xref, sizeof(double) * nqp * ctx->dim);
    memcpy(ctx->w, w, sizeof(double) * nqp);
    return FEMB_OK;
}

int32_t femb_evaluator_set(femb_ctx* ctx, int32_t id, int32_t space_id, int32_t op, const double* refvals, const double* refderivs) {
    FEMB_CHECK_CTX(ctx);
    if (id < 0 || id >= FEMB_MAX_EVALS) return femb_fail(ctx, FEMB_ERR_ARG, "eval_id out of range");
    if (space_id < 0 || space_id >= FEMB_MAX_SPACES || !ctx->spaces[space_id].set) return femb_fail(ctx, FEMB_ERR_ARG, "unknown space");
    if (!ctx->nqp) return femb_fail(ctx, FEMB_ERR_ARG, "femb_quadrature_set must precede femb_evaluator_set");
    const FembSpace& s = ctx->spaces[space_id];
    if (op != FEMB_OP_IDENTITY && op != FEMB_OP_GRADIENT && op != FEMB_OP_DIVERGENCE)
        return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "operator %d is not on the assembly path (Identity, Gradient, Divergence)", op);
    if (op == FEMB_OP_GRADIENT && s.family == FEMB_FAMILY_HDIV) return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "Gradient of Hdiv functions is out of scope");
    if (op != FEMB_OP_IDENTITY && s.family == FEMB_FAMILY_HCURL)
        return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "Hcurl spaces: only the Identity push-forward (feevaluator_hcurl.jl:2-16) is on the path");
    if (op == FEMB_OP_IDENTITY && !refvals) return femb_fail(ctx, FEMB_ERR_ARG, "Identity needs refvals");
    if (op != FEMB_OP_IDENTITY && !refderivs) return femb_fail(ctx, FEMB_ERR_ARG, "derivative operators need refderivs");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FembEval& e = ctx->evals[id];
    e.space = space_id;
    e.op = op;
    e.nqp = ctx->nqp;
    e.resultdim = result_dim(op, ctx->dim, s.ncomp);
    memcpy(e.xref, ctx->xref, sizeof(e.xref));
    memcpy(e.w, ctx->w, sizeof(e.w));
    size_t nv = (size_t)e.nqp * s.nd_all * s.ncomp;
    femb_free(&e.refvals);
    femb_free(&e.refderivs);
    e.h_refvals.clear();
    e.h_refderivs.clear();
    femb_free(&e.reftensor);
    e.reftensor_nt = 0;
    if (refvals) {
        e.h_refvals.assign(refvals, refvals + nv);
        FEMB_TRY(femb_alloc(ctx, &e.refvals, nv));
        FEMB_TRY(femb_h2d(ctx, e.refvals, refvals, nv));
    }
    if (refderivs) {
        e.h_refderivs.assign(refderivs, refderivs + nv * ctx->dim);
        FEMB_TRY(femb_alloc(ctx, &e.refderivs, nv * ctx->dim));
        FEMB_TRY(femb_h2d(ctx, e.refderivs, refderivs, nv * ctx->dim));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    e.set = true;
    return FEMB_OK;
}

int32_t femb_matrix_layout(femb_ctx* ctx, int32_t nr, const int32_t* rows, int32_t nc, const int32_t* cols) {
    FEMB_CHECK_CTX(ctx);
    if (nr <= 0 || nc <= 0 || nr > FEMB_MAX_SPACES || nc > FEMB_MAX_SPACES || !rows || !cols) return femb_fail(ctx, FEMB_ERR_ARG, "bad layout");
    ctx->row_off[0] = ctx->col_off[0] = 0;
    for (int i = 0; i < nr; i++) {
        if (rows[i] < 0 || rows[i] >= FEMB_MAX_SPACES || !ctx->spaces[rows[i]].set) return femb_fail(ctx, FEMB_ERR_ARG, "row space %d not set", rows[i]);
        ctx->row_spaces[i] = rows[i];
        ctx->row_off[i + 1] = ctx->row_off[i] + ctx->spaces[rows[i]].ndofs;
    }
    for (int i = 0; i < nc; i++) {
        if (cols[i] < 0 || cols[i] >= FEMB_MAX_SPACES || !ctx->spaces[cols[i]].set) return femb_fail(ctx, FEMB_ERR_ARG, "col space %d not set", cols[i]);
        ctx->col_spaces[i] = cols[i];
        ctx->col_off[i + 1] = ctx->col_off[i] + ctx->spaces[cols[i]].ndofs;
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    bool resized = (ctx->nrows != ctx->row_off[nr] || ctx->ncols != ctx->col_off[nc]);
    ctx->nrs = nr;
    ctx->ncs = nc;
    ctx->nrows = ctx->row_off[nr];
    ctx->ncols = ctx->col_off[nc];
    if (resized) {
        ctx->extra_entries.clear();  // their keys were encoded against the old row count
        femb_free(&ctx->dirval);      // boundary data in the old numbering
    }
    if (resized || !ctx->bvec) {
        femb_free_pattern(ctx);
        FEMB_TRY(femb_alloc(ctx, &ctx->bvec, (size_t)ctx->nrows));
        FEMB_TRY(femb_alloc(ctx, &ctx->sol, (size_t)ctx->ncols));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->bvec, 0, sizeof(double) * ctx->nrows, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->sol, 0, sizeof(double) * ctx->ncols, ctx->stream));
    }
    return FEMB_OK;
}

int32_t femb_matrix_zero(femb_ctx* ctx) {
    FEMB_CHECK_CTX(ctx);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->dirty_A = false;
    ctx->zero_pending_A = true;  // applied lazily (femb_flush_zero) or skipped by write-once kernels
    ctx->zero_mask = ~0ull;
    if (ctx->touched) CUDA_TRY(ctx, cudaMemsetAsync(ctx->touched, 0, ctx->nnz, ctx->stream));
    return FEMB_OK;
}

int32_t femb_vector_zero(femb_ctx* ctx) {
    FEMB_CHECK_CTX(ctx);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->dirty_b = false;
    ctx->zero_pending_b = true;
    return FEMB_OK;
}

int32_t femb_matrix_get(femb_ctx* ctx, double* out) {
    FEMB_CHECK_CTX(ctx);
    if (!out || !ctx->nzval) return femb_fail(ctx, FEMB_ERR_ARG, "no matrix values");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    FEMB_TRY(femb_d2h(ctx, out, ctx->nzval, (size_t)ctx->nnz));
    return femb_sync(ctx);
}

int32_t femb_matrix_get_range(femb_ctx* ctx, int64_t first, int64_t count, double* out) {
    FEMB_CHECK_CTX(ctx);
    if (!out || !ctx->nzval || first < 0 || count < 0 || first + count > ctx->nnz) return femb_fail(ctx, FEMB_ERR_ARG, "femb_matrix_get_range: bad range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    FEMB_TRY(femb_d2h(ctx, out, ctx->nzval + first, (size_t)count));
    return femb_sync(ctx);
}

int32_t femb_vector_get_range(femb_ctx* ctx, int64_t first, int64_t count, double* out) {
    FEMB_CHECK_CTX(ctx);
    if (!out || !ctx->bvec || first < 0 || count < 0 || first + count > ctx->nrows) return femb_fail(ctx, FEMB_ERR_ARG, "femb_vector_get_range: bad range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(femb_flush_zero(ctx, false, true));
    FEMB_TRY(femb_d2h(ctx, out, ctx->bvec + first, (size_t)count));
    return femb_sync(ctx);
}

__global__ void k_axpy_inplace(double* __restrict__ y, const double* __restrict__ x, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

int32_t femb_matrix_add_to(femb_ctx* ctx, double* inout) {
    FEMB_CHECK_CTX(ctx);
    if (!inout || !ctx->nzval) return femb_fail(ctx, FEMB_ERR_ARG, "no matrix values");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    double* tmp = nullptr;
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    FEMB_TRY(femb_alloc(ctx, &tmp, (size_t)ctx->nnz));
    int st = femb_h2d(ctx, tmp, inout, (size_t)ctx->nnz);
    if (st == FEMB_OK) {
        k_axpy_inplace<<<(unsigned)((ctx->nnz + 255) / 256), 256, 0, ctx->stream>>>(tmp, ctx->nzval, ctx->nnz);
        ctx->launches++;
        st = femb_d2h(ctx, inout, tmp, (size_t)ctx->nnz);
    }
    if (st == FEMB_OK) st = femb_sync(ctx);
    cudaFree(tmp);
    return st;
}

int32_t femb_vector_get(femb_ctx* ctx, double* out) {
    FEMB_CHECK_CTX(ctx);
    if (!out || !ctx->bvec) return femb_fail(ctx, FEMB_ERR_ARG, "no vector");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(femb_flush_zero(ctx, false, true));
    FEMB_TRY(femb_d2h(ctx, out, ctx->bvec, (size_t)ctx->nrows));
    return femb_sync(ctx);
}

int32_t femb_vector_set(femb_ctx* ctx, int32_t which, const double* v) {
    FEMB_CHECK_CTX(ctx);
    if (!v || (which != 0 && which != 1) || !ctx->bvec) return femb_fail(ctx, FEMB_ERR_ARG, "femb_vector_set: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (which == 0) {
        ctx->zero_pending_b = false;
        ctx->dirty_b = true;
        FEMB_TRY(femb_h2d(ctx, ctx->bvec, v, (size_t)ctx->nrows));
    }
    else FEMB_TRY(femb_h2d(ctx, ctx->sol, v, (size_t)ctx->ncols));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FEMB_OK;
}

int32_t femb_set_scatter(femb_ctx* ctx, int32_t policy) {
    FEMB_CHECK_CTX(ctx);
    if (policy < 0 || policy > 2) return femb_fail(ctx, FEMB_ERR_ARG, "unknown scatter policy");
    ctx->scatter = policy;
    return FEMB_OK;
}

int32_t femb_last_kernel_ms(femb_ctx* ctx, float* ms, int32_t* n) {
    FEMB_CHECK_CTX(ctx);
    if (ms) *ms = ctx->last_ms;
    if (n) *n = ctx->last_launches;
    return FEMB_OK;
}

int64_t femb_launch_count(const femb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t femb_region_begin(femb_ctx* ctx) {
    FEMB_CHECK_CTX(ctx);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
    return FEMB_OK;
}

int32_t femb_region_end(femb_ctx* ctx, float* ms) {
    FEMB_CHECK_CTX(ctx);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev3, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev3));
    float t = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&t, ctx->ev2, ctx->ev3));
    if (ms) *ms = t;
    return FEMB_OK;
}

void* femb_device_ptr(femb_ctx* ctx, int32_t which) {
    if (!ctx) return nullptr;
    switch (which) {
        case 0: return ctx->nzval;
        case 1: return ctx->bvec;
        case 2: return ctx->colptr;
        case 3: return ctx->rowval;
        default: return nullptr;
    }
}

}  // extern "C"
