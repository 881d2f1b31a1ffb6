// The steps right after the assembly in the reference's drivers, kept on the device so that A and b need no host round trip
// (SURVEY.md §8(f)3-4):
//   boundarydofs(FES)                      /root/reference/src/dofmaps.jl:871-899 (used at Example200:86, Example210:172)
//   A[dof,dof] = 1e60; b[dof] = 1e60 * u   Example200_LowLevelPoisson.jl:86-92, Example210_LowLevelNavierStokes.jl:197-200
//   res = A * sol - b; res[fixed] = 0      Example210:180-184, :204-208  (norm(res) drives the Newton loop)
#include <cub/cub.cuh>

#include "femb_internal.h"

namespace {

const int8_t h_tri_faces[3][3] = {{0, 1, -1}, {1, 2, -1}, {2, 0, -1}};
const int8_t h_tet_faces[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {0, 3, 2}};
const int8_t h_tet_edges[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

#define FEMB_MAX_FACE_DOFS 64
struct FaceDofTable {
    int n[4];
    int16_t loc[4][FEMB_MAX_FACE_DOFS];  // local dof numbers (0-based, all components) attached to each local face
};

__global__ void k_mark_boundary_dofs(const FaceDofTable t, int per, int nd, const int32_t* __restrict__ facecells, const int32_t* __restrict__ cellfaces,
                                     const int32_t* __restrict__ celldofs, int64_t nfaces, uint8_t* __restrict__ flags) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= nfaces || facecells[2 * f + 1] != 0) return;  // interior face
    const int64_t cell = facecells[2 * f] - 1;
    int lf = 0;
    while (lf < per - 1 && cellfaces[cell * per + lf] != (int32_t)(f + 1)) lf++;
    for (int i = 0; i < t.n[lf]; i++) flags[celldofs[cell * nd + t.loc[lf][i]] - 1] = 1;
}

__global__ void k_fixed_list(const int32_t* __restrict__ bdofs, int64_t nb, int64_t offset, const int64_t* __restrict__ extra, int64_t nextra,
                             int64_t* __restrict__ fixed) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nb) fixed[i] = (int64_t)bdofs[i] - 1 + offset;
    else if (i < nb + nextra) fixed[i] = extra[i - nb] - 1;
}

__global__ void k_dirichlet(const int64_t* __restrict__ fixed, int64_t n, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                            double* __restrict__ nzval, double* __restrict__ b, const double* __restrict__ values, double penalty, int64_t nrows,
                            int32_t* __restrict__ errflag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t d = fixed[i];
    if (d < 0 || d >= nrows) {
        atomicAdd(errflag, 1);
        return;
    }
    int64_t lo = colptr[d] - 1, hi = colptr[d + 1] - 1;
    const int64_t end = hi;
    while (lo < hi) {  // rows ascend within the column
        const int64_t mid = (lo + hi) >> 1;
        if (rowval[mid] < d + 1) lo = mid + 1;
        else hi = mid;
    }
    if (lo < end && rowval[lo] == d + 1) nzval[lo] = penalty;
    else atomicAdd(errflag, 1);  // no diagonal entry in the pattern
    b[d] = values ? penalty * values[d] : 0.0;
}

// ---- CSR view of the CSC pattern (stable sort of the slots by row): deterministic, write-once products ----
__global__ void k_csr_keys(const int64_t* __restrict__ rowval, int64_t nnz, uint32_t* __restrict__ key, uint32_t* __restrict__ slot) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    key[p] = (uint32_t)(rowval[p] - 1);
    slot[p] = (uint32_t)p;
}

__global__ void k_csr_cols(const int64_t* __restrict__ colptr, int64_t ncols, const uint32_t* __restrict__ slot, int64_t nnz, int32_t* __restrict__ col) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    const int64_t p1 = (int64_t)slot[q] + 1;  // 1-based slot: the column j with colptr[j] <= p1 < colptr[j+1]
    int64_t lo = 0, hi = ncols;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (colptr[mid + 1] <= p1) lo = mid + 1;
        else hi = mid;
    }
    col[q] = (int32_t)lo;
}

__global__ void k_csr_ptr(const uint32_t* __restrict__ skey, int64_t nnz, int64_t nrows, int64_t* __restrict__ ptr) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > nrows) return;
    int64_t lo = 0, hi = nnz;  // first position with key >= r
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)skey[mid] < r) lo = mid + 1;
        else hi = mid;
    }
    ptr[r] = lo;
}

// res = A x - b, 8 lanes per row, fixed reduction tree
__global__ void __launch_bounds__(256) k_residual(const int64_t* __restrict__ ptr, const int32_t* __restrict__ col, const uint32_t* __restrict__ slot,
                                                  const double* __restrict__ nzval, const double* __restrict__ x, const double* __restrict__ b,
                                                  int64_t nrows, double* __restrict__ res) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t row = t >> 3;
    const int sub = (int)(t & 7);
    double acc = 0.0;
    if (row < nrows) {
        const int64_t e = ptr[row + 1];
        for (int64_t q = ptr[row] + sub; q < e; q += 8) acc += nzval[slot[q]] * x[col[q]];
    }
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
    if (row < nrows && sub == 0) res[row] = acc - b[row];
}

__global__ void k_zero_rows(const int64_t* __restrict__ fixed, int64_t n, int64_t nrows, double* __restrict__ res) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && fixed[i] >= 0 && fixed[i] < nrows) res[fixed[i]] = 0.0;
}

__global__ void k_add_inplace(double* __restrict__ y, const double* __restrict__ x, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

// a[row - r0] += factor * sum_{col in [c0, c1)} A[row, col] b[col - c0]   (row-wise view, 8 lanes per row, fixed tree)
__global__ void __launch_bounds__(256) k_block_matvec(const int64_t* __restrict__ ptr, const int32_t* __restrict__ col, const uint32_t* __restrict__ slot,
                                                      const double* __restrict__ nzval, const double* __restrict__ b, int64_t r0, int64_t r1, int64_t c0,
                                                      int64_t c1, double factor, double* __restrict__ a) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t row = r0 + (t >> 3);
    const int sub = (int)(t & 7);
    double acc = 0.0;
    if (row < r1) {
        const int64_t e = ptr[row + 1];
        for (int64_t q = ptr[row] + sub; q < e; q += 8) {
            const int64_t c = col[q];
            if (c >= c0 && c < c1) acc += nzval[slot[q]] * b[c - c0];
        }
    }
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
    if (row < r1 && sub == 0) a[row - r0] += factor * acc;
}

// transposed: a[col - c0] += factor * sum_{row in [r0, r1)} A[row, col] b[row - r0]   (CSC columns directly)
// LR = 1: out[col - c0] = factor * (b1 - b2)[col] * sum_row A[row, col] (a1 - a2)[row]  (terms of lrmatmul / ldrdmatmul)
template <int LR>
__global__ void __launch_bounds__(256) k_block_matvec_t(const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, const double* __restrict__ nzval,
                                                        const double* __restrict__ b, const double* __restrict__ b2, const double* __restrict__ c1v,
                                                        const double* __restrict__ c2v, int64_t r0, int64_t r1, int64_t c0, int64_t c1, double factor,
                                                        double* __restrict__ a) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t cl = c0 + (t >> 3);
    const int sub = (int)(t & 7);
    double acc = 0.0;
    if (cl < c1) {
        const int64_t e = colptr[cl + 1] - 1;
        for (int64_t q = colptr[cl] - 1 + sub; q < e; q += 8) {
            const int64_t r = rowval[q] - 1;
            if (r >= r0 && r < r1) acc += nzval[q] * (b[r - r0] - (LR && b2 ? b2[r - r0] : 0.0));
        }
    }
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
    if (cl < c1 && sub == 0) {
        if (LR) a[cl - c0] = factor * acc * (c1v[cl - c0] - (c2v ? c2v[cl - c0] : 0.0));
        else a[cl - c0] += factor * acc;
    }
}

struct Square {
    __host__ __device__ double operator()(double v) const { return v * v; }
};

int bits_for(uint64_t v) {
    int b = 1;
    while (b < 64 && (v >> b)) b++;
    return b;
}

int build_csr(femb_ctx* ctx) {
    if (ctx->csr_ptr) return FEMB_OK;
    const int64_t nnz = ctx->nnz;
    if (nnz >= ((int64_t)1 << 32) || ctx->nrows >= ((int64_t)1 << 31) || ctx->ncols >= ((int64_t)1 << 31))
        return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "the row-wise view needs nnz < 2^32");
    uint32_t *key = nullptr, *key_alt = nullptr, *slot_alt = nullptr;
    void* tmp = nullptr;
    int rc = FEMB_OK;
    cudaError_t e = cudaSuccess;
    cudaStream_t st = ctx->stream;
    const unsigned grid = (unsigned)((std::max<int64_t>(nnz, 1) + 255) / 256);
    if ((rc = femb_alloc(ctx, &key, (size_t)nnz)) == FEMB_OK && (rc = femb_alloc(ctx, &key_alt, (size_t)nnz)) == FEMB_OK &&
        (rc = femb_alloc(ctx, &slot_alt, (size_t)nnz)) == FEMB_OK && (rc = femb_alloc(ctx, &ctx->csr_slot, (size_t)nnz)) == FEMB_OK &&
        (rc = femb_alloc(ctx, &ctx->csr_col, (size_t)nnz)) == FEMB_OK && (rc = femb_alloc(ctx, &ctx->csr_ptr, (size_t)ctx->nrows + 1)) == FEMB_OK) {
        if (nnz) k_csr_keys<<<grid, 256, 0, st>>>(ctx->rowval, nnz, key, slot_alt);
        cub::DoubleBuffer<uint32_t> dk(key, key_alt), dv(slot_alt, ctx->csr_slot);
        size_t tb = 0;
        e = cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, nnz, 0, bits_for((uint64_t)ctx->nrows), st);
        if (e == cudaSuccess) e = cudaMalloc(&tmp, std::max<size_t>(tb, 1));
        if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, nnz, 0, bits_for((uint64_t)ctx->nrows), st);
        if (e == cudaSuccess && dv.Current() != ctx->csr_slot)
            e = cudaMemcpyAsync(ctx->csr_slot, dv.Current(), sizeof(uint32_t) * nnz, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) {
            if (nnz) k_csr_cols<<<grid, 256, 0, st>>>(ctx->colptr, ctx->ncols, ctx->csr_slot, nnz, ctx->csr_col);
            k_csr_ptr<<<(unsigned)((ctx->nrows + 256) / 256), 256, 0, st>>>(dk.Current(), nnz, ctx->nrows, ctx->csr_ptr);
            e = cudaStreamSynchronize(st);
            if (e == cudaSuccess) e = cudaGetLastError();
        }
        ctx->launches += 4;
    }
    femb_free(&key);
    femb_free(&key_alt);
    femb_free(&slot_alt);
    if (tmp) cudaFree(tmp);
    if (rc == FEMB_OK && e != cudaSuccess) rc = femb_fail(ctx, FEMB_ERR_CUDA, "row-wise view of the pattern: %s", cudaGetErrorString(e));
    if (rc != FEMB_OK) femb_free_csr(ctx);
    return rc;
}

}  // namespace

void femb_free_csr(femb_ctx* ctx) {
    femb_free(&ctx->csr_ptr);
    femb_free(&ctx->csr_col);
    femb_free(&ctx->csr_slot);
}

void femb_free_saved(femb_ctx* ctx) {
    femb_free(&ctx->saved_nzval);
    femb_free(&ctx->saved_b);
}

void femb_free_system(femb_ctx* ctx) {
    femb_free_csr(ctx);
    femb_free_saved(ctx);
    femb_free(&ctx->bdofs);
    femb_free(&ctx->fixed);
    femb_free(&ctx->resvec);
    femb_free(&ctx->dirval);
    ctx->nbdofs = ctx->nfixed = 0;
}

// ---- interpolate!(target, FES, u) for nodal H1 spaces with node dofs and at most one dof per edge (H1P1, H1P2, H1Pk order <= 2) ----
// node dofs: point evaluation (NodalInterpolator, interpolators.jl:93-127); edge dofs: the edge mean is preserved
// (MomentInterpolator of order 0, h1_p2.jl:47-76 / interpolators.jl:330-460): with mean(phi_vertex) = 1/6 and
// mean(phi_edge) = 2/3 on the reference edge, value = (mean_e(u) - (u(a) + u(b)) / 6) * 3/2, mean_e(u) by the caller's
// rule on [0,1] (the reference: Gauss rule of order 2 * order_FE).  nq = 0: the value at the edge midpoint instead.
__constant__ int8_t d_tri_faces[3][2] = {{0, 1}, {1, 2}, {2, 0}};
__constant__ int8_t d_tet_edges[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

struct InterpFunc {
    int kind, ncomp;
    double p[16];  // FEMB_FUNC_AFFINE: per component c0, g_1..g_dim;  FEMB_FUNC_EX210_U: {mu, t}
};
struct InterpArgs {
    const double* coords;
    const int32_t* cellnodes;
    const int32_t* celldofs;
    int64_t ncells;
    int nd, nds, nnpc, nedge;      // local dofs (all components), per component, nodes per cell, edge dofs per component
    int nq;
    double qx[16], qw[16];
    const uint8_t* only;           // per dof of the space: 1 = write (NULL: all)
    double* target;                // + dof_offset already applied
};

template <int DIM>
__device__ __forceinline__ void interp_eval(const InterpFunc& f, const double* x, double* u) {
    if (f.kind == FEMB_FUNC_EX210_U) {  // Example210:51-57
        const double pi = 3.14159265358979323846;
        const double e = exp(-8.0 * pi * pi * f.p[0] * f.p[1]);
        u[0] = e * sin(2.0 * pi * x[0]) * sin(2.0 * pi * x[1]);
        u[1] = e * cos(2.0 * pi * x[0]) * cos(2.0 * pi * x[1]);
    } else {
        for (int c = 0; c < f.ncomp; c++) {
            double s = f.p[c * (DIM + 1)];
#pragma unroll
            for (int d = 0; d < DIM; d++) s += f.p[c * (DIM + 1) + 1 + d] * x[d];
            u[c] = s;
        }
    }
}

template <int DIM>
__global__ void k_interpolate(const InterpArgs a, const InterpFunc f) {
    const int per = a.nnpc + a.nedge;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= a.ncells * per) return;
    const int64_t cell = t / per;
    const int l = (int)(t - cell * per);
    // skip early when none of the component dofs is selected
    if (a.only) {
        bool any = false;
        for (int c = 0; c < f.ncomp; c++) any = any || a.only[a.celldofs[cell * a.nd + c * a.nds + l] - 1];
        if (!any) return;
    }
    double u[3] = {0.0, 0.0, 0.0};
    if (l < a.nnpc) {
        const int64_t n = a.cellnodes[cell * a.nnpc + l] - 1;
        double x[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) x[d] = a.coords[n * DIM + d];
        interp_eval<DIM>(f, x, u);
    } else {
        const int e = l - a.nnpc;
        const int i0 = DIM == 2 ? d_tri_faces[e][0] : d_tet_edges[e][0], i1 = DIM == 2 ? d_tri_faces[e][1] : d_tet_edges[e][1];
        int64_t na = a.cellnodes[cell * a.nnpc + i0] - 1, nb = a.cellnodes[cell * a.nnpc + i1] - 1;
        if (na > nb) {  // orient by the global node ids: every cell sharing the edge computes bit-identical values
            const int64_t tmp = na; na = nb; nb = tmp;
        }
        double xa[DIM], xb[DIM], x[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            xa[d] = a.coords[na * DIM + d];
            xb[d] = a.coords[nb * DIM + d];
        }
        if (a.nq == 0) {
#pragma unroll
            for (int d = 0; d < DIM; d++) x[d] = 0.5 * (xa[d] + xb[d]);
            interp_eval<DIM>(f, x, u);
        } else {
            double ua[3], ub[3], m[3] = {0.0, 0.0, 0.0}, v[3];
            interp_eval<DIM>(f, xa, ua);
            interp_eval<DIM>(f, xb, ub);
            for (int q = 0; q < a.nq; q++) {
#pragma unroll
                for (int d = 0; d < DIM; d++) x[d] = xa[d] + a.qx[q] * (xb[d] - xa[d]);
                interp_eval<DIM>(f, x, v);
                for (int c = 0; c < f.ncomp; c++) m[c] += a.qw[q] * v[c];
            }
            for (int c = 0; c < f.ncomp; c++) u[c] = (m[c] - (ua[c] + ub[c]) * (1.0 / 6.0)) * 1.5;
        }
    }
    for (int c = 0; c < f.ncomp; c++) {
        const int64_t dof = (int64_t)a.celldofs[cell * a.nd + c * a.nds + l] - 1;
        if (!a.only || a.only[dof]) a.target[dof] = u[c];
    }
}

__global__ void k_flag_list(const int32_t* __restrict__ list, int64_t n, uint8_t* __restrict__ flags) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) flags[list[i] - 1] = 1;
}

extern "C" {

int32_t femb_boundary_dofs(femb_ctx* ctx, int32_t id, int32_t nseg, const int32_t* segs, int64_t* n_out) {
    FEMB_CHECK_CTX(ctx);
    if (id < 0 || id >= FEMB_MAX_SPACES || !ctx->spaces[id].set) return femb_fail(ctx, FEMB_ERR_ARG, "unknown space");
    if (nseg <= 0 || nseg > FEMB_MAX_PATTERN_SEGS || !segs) return femb_fail(ctx, FEMB_ERR_ARG, "1..%d pattern segments expected", FEMB_MAX_PATTERN_SEGS);
    const FembSpace& s = ctx->spaces[id];
    FEMB_TRY(femb_build_entities(ctx, 0));
    const FembEntities& F = ctx->ent[0];
    const int dim = ctx->dim, nfl = dim + 1, nfn = dim, nel = dim == 2 ? 3 : 6;
    const int per[4] = {dim + 1, dim + 1, nel, 1};
    // local dofs attached to each local face, component-major like CellDofs (dofmaps.jl:523-559)
    int nd_each = 0;
    for (int g = 0; g < nseg; g++)
        if (segs[3 * g + 1]) nd_each += per[segs[3 * g]] * segs[3 * g + 2];
    FaceDofTable t;
    memset(&t, 0, sizeof(t));
    auto face_node = [&](int f, int i) { return dim == 2 ? h_tri_faces[f][i] : h_tet_faces[f][i]; };
    auto edge_node = [&](int e, int i) { return dim == 2 ? h_tri_faces[e][i] : h_tet_edges[e][i]; };
    auto on_face = [&](int f, int node) {
        for (int i = 0; i < nfn; i++)
            if (face_node(f, i) == node) return true;
        return false;
    };
    for (int pass = 0; pass <= s.ncomp; pass++) {
        const int each = pass < s.ncomp ? 1 : 0;
        int pos = pass * nd_each;
        for (int g = 0; g < nseg; g++) {
            const int type = segs[3 * g], q = segs[3 * g + 2];
            if ((segs[3 * g + 1] ? 1 : 0) != each) continue;
            if (type < 0 || type > 3 || q <= 0) return femb_fail(ctx, FEMB_ERR_ARG, "bad pattern segment %d", g);
            for (int f = 0; f < nfl; f++) {
                auto push = [&](int local) -> bool {
                    if (t.n[f] >= FEMB_MAX_FACE_DOFS) return false;
                    t.loc[f][t.n[f]++] = (int16_t)local;
                    return true;
                };
                bool ok = true;
                if (type == 0) {
                    for (int i = 0; i < nfn; i++)
                        for (int m = 0; m < q; m++) ok = ok && push(pos + face_node(f, i) * q + m);
                } else if (type == 1) {
                    for (int m = 0; m < q; m++) ok = ok && push(pos + f * q + m);
                } else if (type == 2) {
                    for (int e = 0; e < nel; e++)
                        if (on_face(f, edge_node(e, 0)) && on_face(f, edge_node(e, 1)))
                            for (int m = 0; m < q; m++) ok = ok && push(pos + e * q + m);
                }
                if (!ok) return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "more than %d dofs on a face", FEMB_MAX_FACE_DOFS);
            }
            pos += per[type] * q;
        }
    }
    for (int f = 0; f < nfl; f++)
        for (int i = 0; i < t.n[f]; i++)
            if (t.loc[f][i] >= s.nd) return femb_fail(ctx, FEMB_ERR_ARG, "the pattern does not match the space (local dof %d of %d)", t.loc[f][i], s.nd);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    femb_free(&ctx->bdofs);
    ctx->nbdofs = 0;
    uint8_t* flags = nullptr;
    int64_t* d_n = nullptr;
    void* tmp = nullptr;
    FEMB_TRY(femb_alloc(ctx, &flags, (size_t)s.ndofs));
    int rc = femb_alloc(ctx, &ctx->bdofs, (size_t)s.ndofs);
    if (rc == FEMB_OK) rc = femb_alloc(ctx, &d_n, (size_t)1);
    cudaError_t e = cudaSuccess;
    if (rc == FEMB_OK) {
        e = cudaMemsetAsync(flags, 0, (size_t)s.ndofs, ctx->stream);
        if (F.n) k_mark_boundary_dofs<<<(unsigned)((F.n + 255) / 256), 256, 0, ctx->stream>>>(t, nfl, s.nd, F.facecells, F.cell_ent, s.celldofs, F.n, flags);
        cub::CountingInputIterator<int32_t> it(1);
        size_t tb = 0;
        if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(nullptr, tb, it, flags, ctx->bdofs, d_n, s.ndofs, ctx->stream);
        if (e == cudaSuccess) e = cudaMalloc(&tmp, std::max<size_t>(tb, 1));
        if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(tmp, tb, it, flags, ctx->bdofs, d_n, s.ndofs, ctx->stream);
        int64_t n = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&n, d_n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        ctx->nbdofs = n;
        ctx->launches += 2;
    }
    femb_free(&flags);
    femb_free(&d_n);
    if (tmp) cudaFree(tmp);
    if (rc != FEMB_OK) return rc;
    if (e != cudaSuccess) return femb_fail(ctx, FEMB_ERR_CUDA, "femb_boundary_dofs: %s", cudaGetErrorString(e));
    if (n_out) *n_out = ctx->nbdofs;
    return FEMB_OK;
}

int32_t femb_boundary_dofs_get(femb_ctx* ctx, int32_t* dofs) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->bdofs || !dofs) return femb_fail(ctx, FEMB_ERR_ARG, "femb_boundary_dofs must precede femb_boundary_dofs_get");
    if (ctx->nbdofs) FEMB_TRY(femb_d2h(ctx, dofs, ctx->bdofs, (size_t)ctx->nbdofs));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FEMB_OK;
}

int32_t femb_interpolate(femb_ctx* ctx, int32_t id, int32_t nseg, const int32_t* segs, int64_t dof_offset, int32_t func_kind, const double* params,
                         int32_t nq, const double* qx, const double* qw, int32_t only_boundary, double* target_or_null) {
    FEMB_CHECK_CTX(ctx);
    if (id < 0 || id >= FEMB_MAX_SPACES || !ctx->spaces[id].set) return femb_fail(ctx, FEMB_ERR_ARG, "unknown space");
    const FembSpace& s = ctx->spaces[id];
    if (!ctx->bvec) return femb_fail(ctx, FEMB_ERR_ARG, "femb_matrix_layout must precede femb_interpolate");
    if (s.family != FEMB_FAMILY_H1 || s.handler != FEMB_HANDLER_NONE || s.nd != s.nd_all || s.nd % s.ncomp)
        return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "femb_interpolate covers nodal H1 spaces without per-cell handlers");
    if (nq < 0 || nq > 16 || (nq && (!qx || !qw)) || !params) return femb_fail(ctx, FEMB_ERR_ARG, "femb_interpolate: bad rule / parameters");
    if (dof_offset < 0 || dof_offset + s.ndofs > ctx->nrows) return femb_fail(ctx, FEMB_ERR_ARG, "femb_interpolate: dof_offset outside the layout");
    const int dim = ctx->dim, nnpc = dim + 1, nel = dim == 2 ? 3 : 6, nds = s.nd / s.ncomp;
    // the dofmap pattern must be: one dof per node, optionally one dof per edge (2-D: faces), every component the same
    bool nodes = false, edges = false;
    for (int g = 0; g < nseg; g++) {
        const int type = segs[3 * g], each = segs[3 * g + 1], q = segs[3 * g + 2];
        if (!each || q != 1) return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "femb_interpolate: pattern segment %d is not one dof per entity and component", g);
        if (type == 0 && !nodes && !edges) nodes = true;
        else if (((type == 1 && dim == 2) || (type == 2 && dim == 3)) && nodes && !edges) edges = true;
        else return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "femb_interpolate: only node dofs followed by one dof per edge are supported");
    }
    if (!nodes || nds != nnpc + (edges ? nel : 0)) return femb_fail(ctx, FEMB_ERR_ARG, "femb_interpolate: pattern does not match the registered space");
    InterpFunc f;
    f.kind = func_kind;
    f.ncomp = s.ncomp;
    for (double& v : f.p) v = 0.0;
    if (func_kind == FEMB_FUNC_AFFINE) {
        for (int i = 0; i < s.ncomp * (dim + 1); i++) f.p[i] = params[i];
    } else if (func_kind == FEMB_FUNC_EX210_U) {
        if (dim != 2 || s.ncomp != 2) return femb_fail(ctx, FEMB_ERR_ARG, "FEMB_FUNC_EX210_U is the 2-D two-component u! of Example210");
        f.p[0] = params[0];
        f.p[1] = params[1];
    } else {
        return femb_fail(ctx, FEMB_ERR_ARG, "unknown function kind %d", func_kind);
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->dirval) {
        FEMB_TRY(femb_alloc(ctx, &ctx->dirval, (size_t)ctx->nrows));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->dirval, 0, sizeof(double) * ctx->nrows, ctx->stream));
    }
    uint8_t* only = nullptr;
    if (only_boundary) {
        if (!ctx->bdofs) return femb_fail(ctx, FEMB_ERR_ARG, "femb_boundary_dofs must precede femb_interpolate(only_boundary)");
        FEMB_TRY(femb_alloc(ctx, &only, (size_t)s.ndofs));
        cudaMemsetAsync(only, 0, (size_t)s.ndofs, ctx->stream);
        if (ctx->nbdofs) k_flag_list<<<(unsigned)((ctx->nbdofs + 255) / 256), 256, 0, ctx->stream>>>(ctx->bdofs, ctx->nbdofs, only);
        ctx->launches++;
    }
    InterpArgs a;
    a.coords = ctx->coords; a.cellnodes = ctx->cellnodes; a.celldofs = s.celldofs; a.ncells = ctx->ncells;
    a.nd = s.nd; a.nds = nds; a.nnpc = nnpc; a.nedge = edges ? nel : 0; a.nq = nq;
    for (int q = 0; q < nq; q++) { a.qx[q] = qx[q]; a.qw[q] = qw[q]; }
    a.only = only;
    a.target = ctx->dirval + dof_offset;
    const int64_t nt = ctx->ncells * (nnpc + a.nedge);
    cudaError_t e = cudaSuccess;
    if (nt) {
        if (dim == 2) k_interpolate<2><<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(a, f);
        else k_interpolate<3><<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(a, f);
        ctx->launches++;
        e = cudaGetLastError();
    }
    int st = e == cudaSuccess ? FEMB_OK : femb_fail(ctx, FEMB_ERR_CUDA, "femb_interpolate: %s", cudaGetErrorString(e));
    if (st == FEMB_OK && target_or_null) st = femb_d2h(ctx, target_or_null, ctx->dirval + dof_offset, (size_t)s.ndofs);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && st == FEMB_OK) st = femb_fail(ctx, FEMB_ERR_CUDA, "femb_interpolate: synchronize failed");
    femb_free(&only);
    return st;
}

int32_t femb_dirichlet_apply(femb_ctx* ctx, int32_t use_boundary_list, int64_t dof_offset, int64_t n_extra, const int64_t* extra_dofs,
                             const double* values, double penalty) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->nzval || !ctx->bvec) return femb_fail(ctx, FEMB_ERR_ARG, "no pattern: femb_pattern_build / femb_pattern_adopt first");
    if (use_boundary_list && !ctx->bdofs) return femb_fail(ctx, FEMB_ERR_ARG, "femb_boundary_dofs must precede femb_dirichlet_apply");
    if (n_extra < 0 || (n_extra > 0 && !extra_dofs)) return femb_fail(ctx, FEMB_ERR_ARG, "bad extra dof list");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t nb = use_boundary_list ? ctx->nbdofs : 0, n = nb + n_extra;
    femb_free(&ctx->fixed);
    ctx->nfixed = 0;
    if (n == 0) return FEMB_OK;
    FEMB_TRY(femb_flush_zero(ctx, true, true));
    FEMB_TRY(femb_alloc(ctx, &ctx->fixed, (size_t)n));
    int64_t* d_extra = nullptr;
    double* d_values = nullptr;
    int rc = FEMB_OK;
    if (n_extra) {
        rc = femb_alloc(ctx, &d_extra, (size_t)n_extra);
        if (rc == FEMB_OK) rc = femb_h2d(ctx, d_extra, extra_dofs, (size_t)n_extra);
    }
    if (rc == FEMB_OK && values) {
        rc = femb_alloc(ctx, &d_values, (size_t)ctx->nrows);
        if (rc == FEMB_OK) rc = femb_h2d(ctx, d_values, values, (size_t)ctx->nrows);
    }
    cudaError_t e = cudaSuccess;
    int32_t nerr = 0;
    if (rc == FEMB_OK) {
        const unsigned grid = (unsigned)((n + 255) / 256);
        // its own flag word: pending "no slot" errors of the assemblies ([0]) stay for femb_sync to report
        e = cudaMemsetAsync(ctx->errflag + 1, 0, sizeof(int32_t), ctx->stream);
        k_fixed_list<<<grid, 256, 0, ctx->stream>>>(ctx->bdofs, nb, dof_offset, d_extra, n_extra, ctx->fixed);
        // values == NULL: the device-resident boundary data of femb_interpolate (zeros if none was interpolated)
        k_dirichlet<<<grid, 256, 0, ctx->stream>>>(ctx->fixed, n, ctx->colptr, ctx->rowval, ctx->nzval, ctx->bvec, d_values ? d_values : ctx->dirval, penalty, ctx->nrows,
                                                  ctx->errflag + 1);
        ctx->launches += 2;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&nerr, ctx->errflag + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    femb_free(&d_extra);
    femb_free(&d_values);
    if (rc != FEMB_OK) return rc;
    if (e != cudaSuccess) return femb_fail(ctx, FEMB_ERR_CUDA, "femb_dirichlet_apply: %s", cudaGetErrorString(e));
    if (nerr) return femb_fail(ctx, FEMB_ERR_PATTERN, "%d fixed dofs are out of range or have no diagonal entry in the pattern", nerr);
    ctx->nfixed = n;
    return FEMB_OK;
}

// block ranges of (brow, bcol) in the matrix layout; brow = bcol = -1: the whole matrix
static int block_range(femb_ctx* ctx, int brow, int bcol, int64_t& r0, int64_t& r1, int64_t& c0, int64_t& c1) {
    if (!ctx->nzval) return femb_fail(ctx, FEMB_ERR_ARG, "no pattern: femb_pattern_build / femb_pattern_adopt first");
    if (brow == -1 && bcol == -1) {
        r0 = c0 = 0; r1 = ctx->nrows; c1 = ctx->ncols;
        return FEMB_OK;
    }
    if (brow < 0 || brow >= ctx->nrs || bcol < 0 || bcol >= ctx->ncs) return femb_fail(ctx, FEMB_ERR_ARG, "block (%d, %d) is outside the matrix layout", brow, bcol);
    r0 = ctx->row_off[brow]; r1 = ctx->row_off[brow + 1];
    c0 = ctx->col_off[bcol]; c1 = ctx->col_off[bcol + 1];
    return FEMB_OK;
}

int32_t femb_block_matmul(femb_ctx* ctx, int32_t brow, int32_t bcol, const double* b, double factor, int32_t transposed, double* a) {
    FEMB_CHECK_CTX(ctx);
    int64_t r0, r1, c0, c1;
    FEMB_TRY(block_range(ctx, brow, bcol, r0, r1, c0, c1));
    if (!a || !b) return femb_fail(ctx, FEMB_ERR_ARG, "femb_block_matmul: NULL vector");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    const int64_t na = transposed ? c1 - c0 : r1 - r0, nb = transposed ? r1 - r0 : c1 - c0;
    if (na == 0) return FEMB_OK;
    if (!transposed) FEMB_TRY(build_csr(ctx));
    double *da = nullptr, *db = nullptr;
    int st = femb_alloc(ctx, &da, (size_t)na);
    if (st == FEMB_OK) st = femb_alloc(ctx, &db, (size_t)std::max<int64_t>(nb, 1));
    if (st == FEMB_OK) st = femb_h2d(ctx, da, a, (size_t)na);
    if (st == FEMB_OK && nb) st = femb_h2d(ctx, db, b, (size_t)nb);
    if (st == FEMB_OK) {
        const unsigned grid = (unsigned)((na * 8 + 255) / 256);
        if (transposed)
            k_block_matvec_t<0><<<grid, 256, 0, ctx->stream>>>(ctx->colptr, ctx->rowval, ctx->nzval, db, nullptr, nullptr, nullptr, r0, r1, c0, c1, factor, da);
        else
            k_block_matvec<<<grid, 256, 0, ctx->stream>>>(ctx->csr_ptr, ctx->csr_col, ctx->csr_slot, ctx->nzval, db, r0, r1, c0, c1, factor, da);
        ctx->launches++;
        st = femb_d2h(ctx, a, da, (size_t)na);
        if (st == FEMB_OK) st = femb_sync(ctx);
    }
    femb_free(&da);
    femb_free(&db);
    return st;
}

int32_t femb_lrmatmul(femb_ctx* ctx, const double* a1, const double* a2, const double* b1, const double* b2, double factor, double* out) {
    FEMB_CHECK_CTX(ctx);
    int64_t r0, r1, c0, c1;
    FEMB_TRY(block_range(ctx, -1, -1, r0, r1, c0, c1));
    if (!a1 || !b1 || !out) return femb_fail(ctx, FEMB_ERR_ARG, "femb_lrmatmul: NULL argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    *out = 0.0;
    if (c1 == 0 || r1 == 0) return FEMB_OK;
    double *da1 = nullptr, *da2 = nullptr, *db1 = nullptr, *db2 = nullptr, *terms = nullptr;
    void* tmp = nullptr;
    int st = femb_alloc(ctx, &da1, (size_t)r1);
    if (st == FEMB_OK) st = femb_alloc(ctx, &db1, (size_t)c1);
    if (st == FEMB_OK) st = femb_alloc(ctx, &terms, (size_t)c1 + 1);
    if (st == FEMB_OK && a2) st = femb_alloc(ctx, &da2, (size_t)r1);
    if (st == FEMB_OK && b2) st = femb_alloc(ctx, &db2, (size_t)c1);
    if (st == FEMB_OK) st = femb_h2d(ctx, da1, a1, (size_t)r1);
    if (st == FEMB_OK) st = femb_h2d(ctx, db1, b1, (size_t)c1);
    if (st == FEMB_OK && a2) st = femb_h2d(ctx, da2, a2, (size_t)r1);
    if (st == FEMB_OK && b2) st = femb_h2d(ctx, db2, b2, (size_t)c1);
    if (st == FEMB_OK) {
        k_block_matvec_t<1><<<(unsigned)((c1 * 8 + 255) / 256), 256, 0, ctx->stream>>>(ctx->colptr, ctx->rowval, ctx->nzval, da1, da2, db1, db2, 0, r1, 0, c1, factor, terms);
        ctx->launches += 2;
        size_t tb = 0;
        cudaError_t e = cub::DeviceReduce::Sum(nullptr, tb, terms, terms + c1, c1, ctx->stream);
        if (e == cudaSuccess) e = cudaMalloc(&tmp, std::max<size_t>(tb, 1));
        if (e == cudaSuccess) e = cub::DeviceReduce::Sum(tmp, tb, terms, terms + c1, c1, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, terms + c1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) st = femb_fail(ctx, FEMB_ERR_CUDA, "femb_lrmatmul: %s", cudaGetErrorString(e));
    }
    femb_free(&da1); femb_free(&da2); femb_free(&db1); femb_free(&db2); femb_free(&terms);
    if (tmp) cudaFree(tmp);
    return st;
}

int32_t femb_system_save(femb_ctx* ctx) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->nzval || !ctx->bvec) return femb_fail(ctx, FEMB_ERR_ARG, "no pattern: femb_pattern_build / femb_pattern_adopt first");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(femb_flush_zero(ctx, true, true));
    if (!ctx->saved_nzval) FEMB_TRY(femb_alloc(ctx, &ctx->saved_nzval, (size_t)ctx->nnz));
    if (!ctx->saved_b) FEMB_TRY(femb_alloc(ctx, &ctx->saved_b, (size_t)ctx->nrows));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->saved_nzval, ctx->nzval, sizeof(double) * ctx->nnz, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->saved_b, ctx->bvec, sizeof(double) * ctx->nrows, cudaMemcpyDeviceToDevice, ctx->stream));
    return FEMB_OK;
}

int32_t femb_system_add_saved(femb_ctx* ctx) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->saved_nzval || !ctx->saved_b) return femb_fail(ctx, FEMB_ERR_ARG, "femb_system_save must precede femb_system_add_saved");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(femb_flush_zero(ctx, true, true));
    if (ctx->nnz) k_add_inplace<<<(unsigned)((ctx->nnz + 255) / 256), 256, 0, ctx->stream>>>(ctx->nzval, ctx->saved_nzval, ctx->nnz);
    if (ctx->nrows) k_add_inplace<<<(unsigned)((ctx->nrows + 255) / 256), 256, 0, ctx->stream>>>(ctx->bvec, ctx->saved_b, ctx->nrows);
    ctx->launches += 2;
    KERNEL_CHECK(ctx);
    return FEMB_OK;
}

int32_t femb_residual(femb_ctx* ctx, const double* x, int32_t zero_fixed, double* res, double* norm_out) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->nzval || !ctx->bvec || !ctx->sol) return femb_fail(ctx, FEMB_ERR_ARG, "no pattern: femb_pattern_build / femb_pattern_adopt first");
    if (ctx->nrows != ctx->ncols) return femb_fail(ctx, FEMB_ERR_ARG, "A x - b needs a square layout");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(build_csr(ctx));
    FEMB_TRY(femb_flush_zero(ctx, true, true));
    if (x) FEMB_TRY(femb_h2d(ctx, ctx->sol, x, (size_t)ctx->ncols));
    if (!ctx->resvec) FEMB_TRY(femb_alloc(ctx, &ctx->resvec, (size_t)ctx->nrows + 1));
    cudaStream_t st = ctx->stream;
    if (ctx->nrows) {
        k_residual<<<(unsigned)((ctx->nrows * 8 + 255) / 256), 256, 0, st>>>(ctx->csr_ptr, ctx->csr_col, ctx->csr_slot, ctx->nzval, ctx->sol, ctx->bvec, ctx->nrows,
                                                                            ctx->resvec);
        ctx->launches++;
        if (zero_fixed && ctx->nfixed) {
            k_zero_rows<<<(unsigned)((ctx->nfixed + 255) / 256), 256, 0, st>>>(ctx->fixed, ctx->nfixed, ctx->nrows, ctx->resvec);
            ctx->launches++;
        }
    }
    KERNEL_CHECK(ctx);
    if (norm_out) {
        cub::TransformInputIterator<double, Square, const double*> sq(ctx->resvec, Square());
        double* d_sum = ctx->resvec + ctx->nrows;
        size_t tb = 0;
        void* tmp = nullptr;
        cudaError_t e = cub::DeviceReduce::Sum(nullptr, tb, sq, d_sum, ctx->nrows, st);
        if (e == cudaSuccess) e = cudaMalloc(&tmp, std::max<size_t>(tb, 1));
        if (e == cudaSuccess) e = cub::DeviceReduce::Sum(tmp, tb, sq, d_sum, ctx->nrows, st);
        double sum = 0.0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&sum, d_sum, sizeof(double), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (tmp) cudaFree(tmp);
        ctx->launches++;
        if (e != cudaSuccess) return femb_fail(ctx, FEMB_ERR_CUDA, "femb_residual: %s", cudaGetErrorString(e));
        *norm_out = sqrt(sum);
    }
    if (res && ctx->nrows) FEMB_TRY(femb_d2h(ctx, res, ctx->resvec, (size_t)ctx->nrows));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return FEMB_OK;
}

}  // extern "C"
