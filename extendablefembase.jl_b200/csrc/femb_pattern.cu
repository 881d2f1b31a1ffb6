// Symbolic phase: dof->cell adjacency, CSC pattern of the FEMatrix, cell->slot maps, P1 gather map.
// Replaces what the reference gets from ExtendableSparse's linked-list insert + flush!
// (call sites Example200:155-157,182; Example210:373-379,391; fematrix.jl:108-118).
// Sort / scan / select primitives come from CUB (shipped with the CUDA toolkit); this phase runs
// once per mesh, the numeric kernels in femb_assemble.cu run every assembly.
#include <cub/cub.cuh>

#include <algorithm>

#include "femb_internal.h"

void femb_free_pattern(femb_ctx* ctx) {
    femb_free_csr(ctx);
    femb_free_saved(ctx);
    femb_free_exchange(ctx);
    femb_free(&ctx->resvec);
    femb_free(&ctx->fixed);
    ctx->nfixed = 0;
    femb_free(&ctx->colptr);
    femb_free(&ctx->rowval);
    femb_free(&ctx->nzval);
    femb_free(&ctx->touched);
    femb_free(&ctx->gmap);
    femb_free(&ctx->gslice_ptr);
    ctx->gmap_rows = 0;
    ctx->gmap_stride = 0;
    ctx->gmap_space = -1;
    femb_free(&ctx->gvol);
    ctx->gvol_valid = false;
    ctx->chunk_n = 0;
    femb_free_hmap(ctx->hm);
    femb_free_hmap(ctx->hm_vec);
    femb_free(&ctx->cellcoef);
    femb_free(&ctx->p2rec);
    femb_free(&ctx->rt0rec);
    femb_free(&ctx->threc);
    femb_free(&ctx->p2vrhs);
    for (auto& b : ctx->blocks) femb_free(&b.slots);
    ctx->blocks.clear();
    femb_reset_block_counts(ctx);
    ctx->nnz = 0;
    ctx->max_collen = 0;
    femb_free(&ctx->color_perm);
    ctx->ncolors = 0;
    ctx->color_ptr.clear();
}

// ---------------------------------------------------------------------------------------------------
// dof -> (cell, local index) adjacency
// ---------------------------------------------------------------------------------------------------
__global__ void k_adj_keys(const int32_t* __restrict__ celldofs, int64_t n, int32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = celldofs[i] - 1;
        vals[i] = (uint32_t)i;
    }
}

__global__ void k_adj_ptr(const int32_t* __restrict__ sorted_keys, int64_t n, int64_t ndofs, int32_t* __restrict__ adj_ptr, int* __restrict__ maxdeg) {
    int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d > ndofs) return;
    // lower_bound(sorted_keys, d)
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < (int32_t)d) lo = mid + 1;
        else hi = mid;
    }
    adj_ptr[d] = (int32_t)lo;
    if (d < ndofs) {
        int64_t lo2 = lo, hi2 = n;
        while (lo2 < hi2) {
            int64_t mid = (lo2 + hi2) >> 1;
            if (sorted_keys[mid] <= (int32_t)d) lo2 = mid + 1;
            else hi2 = mid;
        }
        atomicMax(maxdeg, (int)(lo2 - lo));
    }
}

int femb_build_adjacency(femb_ctx* ctx, int sid) {
    FembSpace& s = ctx->spaces[sid];
    if (s.adj_ptr) return FEMB_OK;
    int64_t n = ctx->ncells * s.nd;
    if (n >= ((int64_t)1 << 31)) return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "ncells*nd_local >= 2^31: adjacency needs 64-bit packing");
    int32_t *k0 = nullptr, *k1 = nullptr;
    uint32_t* v0 = nullptr;
    void* tmp = nullptr;
    int* maxdeg = nullptr;
    int st = FEMB_OK;
    auto cleanup = [&]() {
        cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(tmp); cudaFree(maxdeg);
    };
#define TRY_A(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return femb_fail(ctx, FEMB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); } } while (0)
    TRY_A(cudaMalloc(&k0, n * sizeof(int32_t)));
    TRY_A(cudaMalloc(&k1, n * sizeof(int32_t)));
    TRY_A(cudaMalloc(&v0, n * sizeof(uint32_t)));
    if ((st = femb_alloc(ctx, &s.adj, (size_t)n)) != FEMB_OK) { cleanup(); return st; }
    if ((st = femb_alloc(ctx, &s.adj_ptr, (size_t)s.ndofs + 1)) != FEMB_OK) { cleanup(); return st; }
    TRY_A(cudaMalloc(&maxdeg, sizeof(int)));
    TRY_A(cudaMemsetAsync(maxdeg, 0, sizeof(int), ctx->stream));
    k_adj_keys<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(s.celldofs, n, k0, v0);
    ctx->launches++;
    int end_bit = 1;
    while (((int64_t)1 << end_bit) < s.ndofs + 1 && end_bit < 31) end_bit++;
    size_t tb = 0;
    TRY_A(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, s.adj, (int)n, 0, end_bit, ctx->stream));
    TRY_A(cudaMalloc(&tmp, tb));
    TRY_A(cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, s.adj, (int)n, 0, end_bit, ctx->stream));  // stable: cells ascending per dof
    k_adj_ptr<<<(unsigned)((s.ndofs + 1 + 255) / 256), 256, 0, ctx->stream>>>(k1, n, s.ndofs, s.adj_ptr, maxdeg);
    ctx->launches += 2;
    TRY_A(cudaMemcpyAsync(&s.max_deg, maxdeg, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TRY_A(cudaStreamSynchronize(ctx->stream));
    TRY_A(cudaGetLastError());
#undef TRY_A
    cleanup();
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// structural pattern
// ---------------------------------------------------------------------------------------------------
struct RowSrc {
    const int32_t* celldofs;
    int nd;
    int64_t off;
};
struct RowSrcList {
    RowSrc s[FEMB_MAX_SPACES];
    int n;
    int total;  // sum of nd
};

// one thread per (adjacency entry of the chunk, candidate row): key = (dof - d0) * nrows + row
__global__ void k_pattern_keys(const uint32_t* __restrict__ adj, const int32_t* __restrict__ adj_ptr, int nd_c, int64_t e0, int64_t ne,
                               int64_t d0, int64_t d1, RowSrcList rows, int64_t nrows, int64_t* __restrict__ keys) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ne * rows.total) return;
    int64_t e = t / rows.total;
    int r = (int)(t - e * rows.total);
    // dof of entry e0+e: upper_bound in adj_ptr[d0..d1]
    int64_t lo = d0, hi = d1;
    int32_t target = (int32_t)(e0 + e);
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) >> 1;
        if (adj_ptr[mid] <= target) lo = mid;
        else hi = mid - 1;
    }
    int64_t cell = adj[e0 + e] / (uint32_t)nd_c;
    int i = 0;
    while (r >= rows.s[i].nd) {
        r -= rows.s[i].nd;
        i++;
    }
    int64_t row = rows.s[i].off + rows.s[i].celldofs[cell * rows.s[i].nd + r] - 1;
    keys[t] = (lo - d0) * nrows + row;
}

__global__ void k_pattern_emit(const int64_t* __restrict__ ukeys, int64_t nu, int64_t nrows, int64_t col0, int64_t* __restrict__ rowval,
                               unsigned long long* __restrict__ colcount) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nu) return;
    int64_t k = ukeys[i];
    int64_t c = k / nrows;
    rowval[i] = k - c * nrows + 1;
    atomicAdd(colcount + col0 + c, 1ULL);
}

__global__ void k_colptr_finish(const unsigned long long* __restrict__ scanned, int64_t ncols, int64_t nnz, int64_t* __restrict__ colptr,
                                int* __restrict__ maxlen) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c > ncols) return;
    int64_t v = (c == ncols) ? nnz : (int64_t)scanned[c];
    colptr[c] = v + 1;
    if (c < ncols) {
        int64_t nx = (c + 1 == ncols) ? nnz : (int64_t)scanned[c + 1];
        atomicMax(maxlen, (int)(nx - v));
    }
}

__global__ void k_max_collen(const int64_t* __restrict__ colptr, int64_t ncols, int* __restrict__ maxlen) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c < ncols) atomicMax(maxlen, (int)(colptr[c + 1] - colptr[c]));
}

static int set_blocks(femb_ctx* ctx, int nblocks, const int32_t* brows, const int32_t* bcols) {
    if (nblocks <= 0 || nblocks > FEMB_MAX_BLOCKS || !brows || !bcols) return femb_fail(ctx, FEMB_ERR_ARG, "bad block list");
    for (int b = 0; b < nblocks; b++) {
        if (brows[b] < 0 || brows[b] >= ctx->nrs || bcols[b] < 0 || bcols[b] >= ctx->ncs) return femb_fail(ctx, FEMB_ERR_ARG, "block (%d,%d) outside the layout", brows[b], bcols[b]);
        FembBlock blk;
        blk.brow = brows[b];
        blk.bcol = bcols[b];
        ctx->blocks.push_back(blk);
    }
    return FEMB_OK;
}

static int finish_pattern(femb_ctx* ctx) {
    FEMB_TRY(femb_alloc(ctx, &ctx->nzval, (size_t)ctx->nnz));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->nzval, 0, sizeof(double) * ctx->nnz, ctx->stream));
    ctx->zero_mask = ~0ull;  // (claims refer to the blocks of the previous array)
    if (ctx->track) {
        FEMB_TRY(femb_alloc(ctx, &ctx->touched, (size_t)ctx->nnz));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->touched, 0, ctx->nnz, ctx->stream));
    }
    for (int i = 0; i < ctx->ncs; i++) FEMB_TRY(femb_build_adjacency(ctx, ctx->col_spaces[i]));
    for (int i = 0; i < ctx->nrs; i++) FEMB_TRY(femb_build_adjacency(ctx, ctx->row_spaces[i]));
    // the cell->slot maps of the cell-parallel kernels are built lazily (femb_build_slotmaps): the owner-computes
    // paths below do not need them, and they are what limits nnz to 2^31
    FEMB_TRY(femb_build_gmap(ctx));
    FEMB_TRY(femb_build_hmap(ctx));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FEMB_OK;
}

extern "C" int32_t femb_pattern_track(femb_ctx* ctx, int32_t enable) {
    FEMB_CHECK_CTX(ctx);
    ctx->track = enable != 0;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->track && ctx->nnz && !ctx->touched) {
        FEMB_TRY(femb_alloc(ctx, &ctx->touched, (size_t)ctx->nnz));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->touched, 0, ctx->nnz, ctx->stream));
    }
    if (!ctx->track) femb_free(&ctx->touched);
    return FEMB_OK;
}

extern "C" int32_t femb_pattern_extra(femb_ctx* ctx, int64_t n, const int64_t* rows, const int64_t* cols) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->nrs) return femb_fail(ctx, FEMB_ERR_ARG, "femb_matrix_layout must precede femb_pattern_extra");
    ctx->extra_entries.clear();
    if (n < 0 || (n && (!rows || !cols))) return femb_fail(ctx, FEMB_ERR_ARG, "femb_pattern_extra: bad arguments");
    for (int64_t i = 0; i < n; i++) {
        if (rows[i] < 1 || rows[i] > ctx->nrows || cols[i] < 1 || cols[i] > ctx->ncols)
            return femb_fail(ctx, FEMB_ERR_ARG, "femb_pattern_extra: entry (%lld,%lld) outside the layout", (long long)rows[i], (long long)cols[i]);
        ctx->extra_entries.push_back((cols[i] - 1) * ctx->nrows + (rows[i] - 1));
    }
    return FEMB_OK;
}

extern "C" int32_t femb_pattern_build(femb_ctx* ctx, int32_t nblocks, const int32_t* brows, const int32_t* bcols, int64_t nextra,
                                      const int64_t* extra, int64_t* nnz_out) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->nrs) return femb_fail(ctx, FEMB_ERR_ARG, "femb_matrix_layout must precede femb_pattern_build");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    femb_free_pattern(ctx);
    FEMB_TRY(set_blocks(ctx, nblocks, brows, bcols));
    ctx->extra_diag.assign(extra ? extra : nullptr, extra ? extra + nextra : nullptr);
    std::sort(ctx->extra_diag.begin(), ctx->extra_diag.end());
    const int64_t nrows = ctx->nrows, ncols = ctx->ncols;

    unsigned long long* colcount = nullptr;
    FEMB_TRY(femb_alloc(ctx, &colcount, (size_t)ncols + 1));
    CUDA_TRY(ctx, cudaMemsetAsync(colcount, 0, sizeof(unsigned long long) * (ncols + 1), ctx->stream));

    int64_t cap = 0, nnz = 0;
    int64_t* rowval = nullptr;
    int64_t *kin = nullptr, *kout = nullptr, *kuniq = nullptr, *nsel = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    int64_t kcap = 0;
    int st = FEMB_OK;
    auto cleanup = [&]() {
        cudaFree(colcount); cudaFree(kin); cudaFree(kout); cudaFree(kuniq); cudaFree(nsel); cudaFree(tmp);
    };
#define TRY_P(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); cudaFree(rowval); return femb_fail(ctx, FEMB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); } } while (0)
    TRY_P(cudaMalloc(&nsel, sizeof(int64_t)));
    const int64_t budget = (int64_t)1 << 27;  // keys per chunk

    for (int bc = 0; bc < ctx->ncs; bc++) {
        int sid = ctx->col_spaces[bc];
        FembSpace& cs = ctx->spaces[sid];
        RowSrcList rows;
        rows.n = 0;
        rows.total = 0;
        for (auto& b : ctx->blocks)
            if (b.bcol == bc) {
                FembSpace& rs = ctx->spaces[ctx->row_spaces[b.brow]];
                rows.s[rows.n++] = RowSrc{rs.celldofs, rs.nd, ctx->row_off[b.brow]};
                rows.total += rs.nd;
            }
        // extra entries falling into this column block, as (local column, global 0-based row), sorted by column:
        // forced diagonals (Dirichlet penalties) and halo rows registered with femb_pattern_extra (multi-GPU)
        std::vector<std::pair<int64_t, int64_t>> ex;
        for (int64_t g : ctx->extra_diag)
            if (g > ctx->col_off[bc] && g <= ctx->col_off[bc + 1]) ex.push_back({g - 1 - ctx->col_off[bc], g - 1});
        for (int64_t key : ctx->extra_entries) {
            int64_t col = key / nrows, row = key - col * nrows;
            if (col >= ctx->col_off[bc] && col < ctx->col_off[bc + 1]) ex.push_back({col - ctx->col_off[bc], row});
        }
        std::sort(ex.begin(), ex.end());
        if (rows.n == 0 && ex.empty()) continue;
        if ((st = femb_build_adjacency(ctx, sid)) != FEMB_OK) { cleanup(); cudaFree(rowval); return st; }
        int64_t total_keys = ctx->ncells * cs.nd * (int64_t)rows.total;
        int nchunks = (int)std::max<int64_t>(1, (total_keys + budget - 1) / budget);
        size_t exi = 0;
        for (int ch = 0; ch < nchunks; ch++) {
            int64_t d0 = cs.ndofs * ch / nchunks, d1 = cs.ndofs * (ch + 1) / nchunks;
            if (d1 == d0) continue;
            int32_t e01[2];
            TRY_P(cudaMemcpyAsync(&e01[0], cs.adj_ptr + d0, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            TRY_P(cudaMemcpyAsync(&e01[1], cs.adj_ptr + d1, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            TRY_P(cudaStreamSynchronize(ctx->stream));
            int64_t e0 = e01[0], ne = e01[1] - e01[0];
            std::vector<int64_t> exkeys;
            while (exi < ex.size() && ex[exi].first < d1) {
                exkeys.push_back((ex[exi].first - d0) * nrows + ex[exi].second);
                exi++;
            }
            int64_t nk = ne * rows.total + (int64_t)exkeys.size();
            if (nk == 0) continue;
            if (nk > kcap) {
                cudaFree(kin); cudaFree(kout); cudaFree(kuniq); cudaFree(tmp);
                kin = kout = kuniq = nullptr; tmp = nullptr;
                kcap = nk + nk / 8;
                TRY_P(cudaMalloc(&kin, kcap * sizeof(int64_t)));
                TRY_P(cudaMalloc(&kout, kcap * sizeof(int64_t)));
                TRY_P(cudaMalloc(&kuniq, kcap * sizeof(int64_t)));
                size_t t1 = 0, t2 = 0;
                TRY_P(cub::DeviceRadixSort::SortKeys(nullptr, t1, kin, kout, kcap, 0, 64, ctx->stream));
                TRY_P(cub::DeviceSelect::Unique(nullptr, t2, kout, kuniq, nsel, kcap, ctx->stream));
                tmp_bytes = std::max(t1, t2);
                TRY_P(cudaMalloc(&tmp, tmp_bytes));
            }
            if (ne * rows.total > 0) {
                int64_t nt = ne * rows.total;
                k_pattern_keys<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(cs.adj, cs.adj_ptr, cs.nd, e0, ne, d0, d1, rows, nrows, kin);
                ctx->launches++;
            }
            if (!exkeys.empty())
                TRY_P(cudaMemcpyAsync(kin + ne * rows.total, exkeys.data(), exkeys.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
            int end_bit = 1;
            {
                double maxkey = (double)(d1 - d0) * (double)nrows;
                while (end_bit < 64 && (double)((uint64_t)1 << end_bit) <= maxkey) end_bit++;
            }
            size_t tb = tmp_bytes;
            TRY_P(cub::DeviceRadixSort::SortKeys(tmp, tb, kin, kout, nk, 0, end_bit, ctx->stream));
            tb = tmp_bytes;
            TRY_P(cub::DeviceSelect::Unique(tmp, tb, kout, kuniq, nsel, nk, ctx->stream));
            int64_t nu = 0;
            TRY_P(cudaMemcpyAsync(&nu, nsel, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            TRY_P(cudaStreamSynchronize(ctx->stream));
            ctx->launches += 4;
            if (nnz + nu > cap) {
                int64_t ncap = std::max<int64_t>(nnz + nu, cap + cap / 2);
                if (ch + 1 < nchunks || bc + 1 < ctx->ncs) ncap = std::max<int64_t>(ncap, (int64_t)((double)(nnz + nu) * 1.05 * nchunks / (ch + 1)));
                int64_t* nr = nullptr;
                TRY_P(cudaMalloc(&nr, ncap * sizeof(int64_t)));
                if (nnz) TRY_P(cudaMemcpyAsync(nr, rowval, nnz * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
                TRY_P(cudaStreamSynchronize(ctx->stream));
                cudaFree(rowval);
                rowval = nr;
                cap = ncap;
            }
            k_pattern_emit<<<(unsigned)((nu + 255) / 256), 256, 0, ctx->stream>>>(kuniq, nu, nrows, ctx->col_off[bc] + d0, rowval + nnz, colcount);
            ctx->launches++;
            nnz += nu;
        }
    }
    // colptr = 1 + exclusive scan of the column counts
    {
        unsigned long long* scanned = nullptr;
        TRY_P(cudaMalloc(&scanned, sizeof(unsigned long long) * (ncols + 1)));
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, colcount, scanned, (int64_t)ncols + 1, ctx->stream);
        void* t2 = nullptr;
        cudaMalloc(&t2, tb);
        cub::DeviceScan::ExclusiveSum(t2, tb, colcount, scanned, (int64_t)ncols + 1, ctx->stream);
        st = femb_alloc(ctx, &ctx->colptr, (size_t)ncols + 1);
        int* maxlen = nullptr;
        cudaMalloc(&maxlen, sizeof(int));
        cudaMemsetAsync(maxlen, 0, sizeof(int), ctx->stream);
        if (st == FEMB_OK) {
            k_colptr_finish<<<(unsigned)((ncols + 1 + 255) / 256), 256, 0, ctx->stream>>>(scanned, ncols, nnz, ctx->colptr, maxlen);
            ctx->launches += 2;
            cudaMemcpyAsync(&ctx->max_collen, maxlen, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        }
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        cudaFree(scanned); cudaFree(t2); cudaFree(maxlen);
        if (st != FEMB_OK) { cleanup(); cudaFree(rowval); return st; }
        TRY_P(e);
    }
#undef TRY_P
    cleanup();
    ctx->rowval = rowval;
    ctx->nnz = nnz;
    if (nnz_out) *nnz_out = nnz;
    return finish_pattern(ctx);
}

extern "C" int32_t femb_pattern_adopt(femb_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* colptr, const int64_t* rowval,
                                      int32_t nblocks, const int32_t* brows, const int32_t* bcols) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->nrs) return femb_fail(ctx, FEMB_ERR_ARG, "femb_matrix_layout must precede femb_pattern_adopt");
    if (nrows != ctx->nrows || ncols != ctx->ncols) return femb_fail(ctx, FEMB_ERR_ARG, "pattern is %lld x %lld but the layout is %lld x %lld", (long long)nrows, (long long)ncols, (long long)ctx->nrows, (long long)ctx->ncols);
    if (!colptr || (!rowval && nnz) || colptr[0] != 1 || colptr[ncols] != nnz + 1) return femb_fail(ctx, FEMB_ERR_ARG, "colptr must be 1-based with colptr[end] == nnz+1");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    femb_free_pattern(ctx);
    FEMB_TRY(set_blocks(ctx, nblocks, brows, bcols));
    ctx->extra_diag.clear();
    ctx->nnz = nnz;
    FEMB_TRY(femb_alloc(ctx, &ctx->colptr, (size_t)ncols + 1));
    FEMB_TRY(femb_alloc(ctx, &ctx->rowval, (size_t)nnz));
    FEMB_TRY(femb_h2d(ctx, ctx->colptr, colptr, (size_t)ncols + 1));
    FEMB_TRY(femb_h2d(ctx, ctx->rowval, rowval, (size_t)nnz));
    int* maxlen = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&maxlen, sizeof(int)));
    cudaMemsetAsync(maxlen, 0, sizeof(int), ctx->stream);
    k_max_collen<<<(unsigned)((ncols + 255) / 256), 256, 0, ctx->stream>>>(ctx->colptr, ncols, maxlen);
    ctx->launches++;
    cudaMemcpyAsync(&ctx->max_collen, maxlen, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(maxlen);
    CUDA_TRY(ctx, e);
    return finish_pattern(ctx);
}

extern "C" int32_t femb_pattern_get(femb_ctx* ctx, int64_t* colptr_out, int64_t* rowval_out) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->colptr) return femb_fail(ctx, FEMB_ERR_ARG, "no pattern");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (colptr_out) FEMB_TRY(femb_d2h(ctx, colptr_out, ctx->colptr, (size_t)ctx->ncols + 1));
    if (rowval_out) FEMB_TRY(femb_d2h(ctx, rowval_out, ctx->rowval, (size_t)ctx->nnz));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FEMB_OK;
}

__global__ void k_gather_i64(const int64_t* __restrict__ src, const int64_t* __restrict__ idx, int64_t n, int64_t* __restrict__ dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

extern "C" int32_t femb_pattern_get_slots(femb_ctx* ctx, int64_t n, const int64_t* slots, int64_t* out) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->rowval) return femb_fail(ctx, FEMB_ERR_ARG, "no pattern");
    if (n < 0 || (n && (!slots || !out))) return femb_fail(ctx, FEMB_ERR_ARG, "femb_pattern_get_slots: bad arguments");
    if (n == 0) return FEMB_OK;
    for (int64_t i = 0; i < n; i++)
        if (slots[i] < 0 || slots[i] >= ctx->nnz) return femb_fail(ctx, FEMB_ERR_ARG, "slot out of range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int64_t *d_idx = nullptr, *d_out = nullptr;
    int st = femb_alloc(ctx, &d_idx, (size_t)n);
    if (st == FEMB_OK) st = femb_alloc(ctx, &d_out, (size_t)n);
    if (st == FEMB_OK) st = femb_h2d(ctx, d_idx, slots, (size_t)n);
    if (st == FEMB_OK) {
        k_gather_i64<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->rowval, d_idx, n, d_out);
        ctx->launches++;
        st = femb_d2h(ctx, out, d_out, (size_t)n);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_idx);
    cudaFree(d_out);
    if (st != FEMB_OK) return st;
    CUDA_TRY(ctx, e);
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// compress: Example200's value filter decides the pattern (Example200:153)
// ---------------------------------------------------------------------------------------------------
__global__ void k_keep_flags(const uint8_t* __restrict__ touched, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int64_t ncols,
                             const int64_t* __restrict__ extra, int64_t nextra, uint8_t* __restrict__ keep, unsigned long long* __restrict__ colcount) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    bool diag_forced = false;
    {
        int64_t lo = 0, hi = nextra;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (extra[mid] < c + 1) lo = mid + 1;
            else hi = mid;
        }
        diag_forced = lo < nextra && extra[lo] == c + 1;
    }
    unsigned long long cnt = 0;
    for (int64_t p = colptr[c] - 1; p < colptr[c + 1] - 1; p++) {
        uint8_t k = touched[p] || (diag_forced && rowval[p] == c + 1);
        keep[p] = k;
        cnt += k;
    }
    colcount[c] = cnt;
}

extern "C" int32_t femb_pattern_compress(femb_ctx* ctx, int64_t* nnz_out) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->touched || !ctx->nnz) return femb_fail(ctx, FEMB_ERR_ARG, "femb_pattern_compress needs femb_pattern_track(ctx,1) before the assembly");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t ncols = ctx->ncols, nnz = ctx->nnz;
    uint8_t* keep = nullptr;
    unsigned long long *colcount = nullptr, *scanned = nullptr;
    int64_t *extra = nullptr, *rv2 = nullptr, *nsel = nullptr;
    double* nz2 = nullptr;
    void* tmp = nullptr;
    int* maxlen = nullptr;
    auto cleanup = [&]() { cudaFree(keep); cudaFree(colcount); cudaFree(scanned); cudaFree(extra); cudaFree(nsel); cudaFree(tmp); cudaFree(maxlen); };
#define TRY_C(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); cudaFree(rv2); cudaFree(nz2); return femb_fail(ctx, FEMB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); } } while (0)
    TRY_C(cudaMalloc(&keep, nnz));
    TRY_C(cudaMalloc(&colcount, sizeof(unsigned long long) * (ncols + 1)));
    TRY_C(cudaMalloc(&scanned, sizeof(unsigned long long) * (ncols + 1)));
    TRY_C(cudaMemsetAsync(colcount, 0, sizeof(unsigned long long) * (ncols + 1), ctx->stream));
    int64_t nextra = (int64_t)ctx->extra_diag.size();
    TRY_C(cudaMalloc(&extra, sizeof(int64_t) * std::max<int64_t>(1, nextra)));
    if (nextra) TRY_C(cudaMemcpyAsync(extra, ctx->extra_diag.data(), sizeof(int64_t) * nextra, cudaMemcpyHostToDevice, ctx->stream));
    k_keep_flags<<<(unsigned)((ncols + 255) / 256), 256, 0, ctx->stream>>>(ctx->touched, ctx->colptr, ctx->rowval, ncols, extra, nextra, keep, colcount);
    TRY_C(cudaMalloc(&rv2, sizeof(int64_t) * nnz));
    TRY_C(cudaMalloc(&nz2, sizeof(double) * nnz));
    TRY_C(cudaMalloc(&nsel, sizeof(int64_t)));
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceSelect::Flagged(nullptr, t1, ctx->rowval, keep, rv2, nsel, nnz, ctx->stream);
    cub::DeviceSelect::Flagged(nullptr, t2, ctx->nzval, keep, nz2, nsel, nnz, ctx->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, t3, colcount, scanned, ncols + 1, ctx->stream);
    size_t tb = std::max(t1, std::max(t2, t3));
    TRY_C(cudaMalloc(&tmp, tb));
    size_t t = tb;
    TRY_C(cub::DeviceSelect::Flagged(tmp, t, ctx->rowval, keep, rv2, nsel, nnz, ctx->stream));
    t = tb;
    TRY_C(cub::DeviceSelect::Flagged(tmp, t, ctx->nzval, keep, nz2, nsel, nnz, ctx->stream));
    t = tb;
    TRY_C(cub::DeviceScan::ExclusiveSum(tmp, t, colcount, scanned, ncols + 1, ctx->stream));
    int64_t nnz2 = 0;
    TRY_C(cudaMemcpyAsync(&nnz2, nsel, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    TRY_C(cudaStreamSynchronize(ctx->stream));
    TRY_C(cudaMalloc(&maxlen, sizeof(int)));
    TRY_C(cudaMemsetAsync(maxlen, 0, sizeof(int), ctx->stream));
    k_colptr_finish<<<(unsigned)((ncols + 1 + 255) / 256), 256, 0, ctx->stream>>>(scanned, ncols, nnz2, ctx->colptr, maxlen);
    TRY_C(cudaMemcpyAsync(&ctx->max_collen, maxlen, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TRY_C(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 6;
#undef TRY_C
    cleanup();
    // swap in the compressed arrays (keep exact-size copies)
    std::vector<FembBlock> blocks = ctx->blocks;
    for (auto& b : ctx->blocks) femb_free(&b.slots);
    femb_reset_block_counts(ctx);
    femb_free(&ctx->rowval);
    femb_free(&ctx->nzval);
    femb_free(&ctx->touched);
    ctx->rowval = rv2;
    ctx->nzval = nz2;
    ctx->nnz = nnz2;
    // everything keyed on the old slot numbering goes: the row-wise view, the saved linear part, the fixed rows, the
    // exchange plan (the caller re-installs it for the new pattern) and the colouring-independent chunk table
    femb_free_csr(ctx);
    femb_free_saved(ctx);
    femb_free(&ctx->fixed);
    ctx->nfixed = 0;
    femb_free_exchange(ctx);
    ctx->chunk_n = 0;
    FEMB_TRY(femb_alloc(ctx, &ctx->touched, (size_t)nnz2));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->touched, 1, nnz2, ctx->stream));
    FEMB_TRY(femb_build_gmap(ctx));
    FEMB_TRY(femb_build_hmap(ctx));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (nnz_out) *nnz_out = nnz2;
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// cell -> slot maps and the P1 gather map
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t find_slot(const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int64_t col, int64_t row1) {
    int64_t lo = colptr[col] - 1, hi = colptr[col + 1] - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rowval[mid] < row1) lo = mid + 1;
        else hi = mid;
    }
    return (lo < colptr[col + 1] - 1 && rowval[lo] == row1) ? lo : -1;
}

__global__ void k_slotmap(const int32_t* __restrict__ rdofs, int ndr, int64_t roff, const int32_t* __restrict__ cdofs, int ndc, int64_t coff,
                          int64_t ncells, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int32_t* __restrict__ slots) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t per = (int64_t)ndr * ndc;
    if (t >= ncells * per) return;
    int64_t cell = t / per;
    int jk = (int)(t - cell * per);
    int j = jk / ndc, k = jk - j * ndc;
    int64_t row1 = roff + rdofs[cell * ndr + j];
    int64_t col = coff + cdofs[cell * ndc + k] - 1;
    slots[t] = (int32_t)find_slot(colptr, rowval, col, row1);
}

int femb_build_slotmaps(femb_ctx* ctx, int only_brow, int only_bcol) {
    if (ctx->nnz >= ((int64_t)1 << 31))
        return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "nnz >= 2^31: the cell-parallel kernels use 32-bit slot maps (the owner-computes gather paths do not)");
    for (auto& b : ctx->blocks) {
        if (b.slots) continue;  // built on first use
        if (only_brow >= 0 && (b.brow != only_brow || b.bcol != only_bcol)) continue;
        const FembSpace& rs = ctx->spaces[ctx->row_spaces[b.brow]];
        const FembSpace& cs = ctx->spaces[ctx->col_spaces[b.bcol]];
        int64_t n = ctx->ncells * rs.nd * cs.nd;
        FEMB_TRY(femb_alloc(ctx, &b.slots, (size_t)n));
        if (n == 0) continue;
        k_slotmap<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(rs.celldofs, rs.nd, ctx->row_off[b.brow], cs.celldofs, cs.nd, ctx->col_off[b.bcol],
                                                                       ctx->ncells, ctx->colptr, ctx->rowval, b.slots);
        KERNEL_CHECK(ctx);
    }
    return FEMB_OK;
}

// Gather map for scalar P1 (dof == node), sliced ELL with one slice per warp of 32 consecutive dofs.
// Structure of arrays, planes na, nb, w, cell (, nc): element [row][lane] at plane + (row << 5) + lane.
//   2-D: w = k | pos_self << 8 | pos_a << 16 | pos_b << 24     (pos 0xFF = not in pattern)
//   3-D: w = k | pos0 << 2 | ... | pos3 << 20                  (pos 0x3F = not in pattern; local order)
// na.. are the cell's nodes other than the dof itself, in local order, 0-based; k is the dof's local index;
// pos_j = position of global row celldofs[j] inside CSC column c.  Padding entries have w = 0xFFFFFFFF and node /
// cell ids 0 (so that unconditional prefetches stay in bounds).
#define FEMB_GMAP_INVALID 0xFFFFFFFFu

__global__ void k_iota_i32(int32_t* __restrict__ p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

__global__ void k_fill_i32(int32_t* __restrict__ p, int64_t n, int32_t v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_slice_deg(const int32_t* __restrict__ adj_ptr, int64_t ndofs, int64_t nslices, int32_t* __restrict__ deg) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nslices) return;
    int m = 0;
    for (int l = 0; l < 32; l++) {
        int64_t c = s * 32 + l;
        if (c < ndofs) m = max(m, adj_ptr[c + 1] - adj_ptr[c]);
    }
    deg[s] = m;
}

template <int ND>
__global__ void k_gmap(const uint32_t* __restrict__ adj, const int32_t* __restrict__ adj_ptr, int64_t ndofs, int64_t nslices,
                       const int32_t* __restrict__ celldofs, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                       const int32_t* __restrict__ slice_ptr, uint32_t* __restrict__ gmap, size_t stride) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nslices * 32) return;
    const int64_t slice = c >> 5;
    const int lane = (int)(c & 31);
    const int row0 = slice_ptr[slice], nrows = slice_ptr[slice + 1] - row0;
    const int e0 = c < ndofs ? adj_ptr[c] : 0;
    const int deg = c < ndofs ? adj_ptr[c + 1] - e0 : 0;
    const int64_t base = c < ndofs ? colptr[c] - 1 : 0;
    for (int r = 0; r < nrows; r++) {
        uint32_t cell = 0, ww = FEMB_GMAP_INVALID, oo[3] = {0u, 0u, 0u};
        if (r < deg) {
            uint32_t a = adj[e0 + r];
            cell = a / ND;
            uint32_t k = a - cell * ND;
            uint32_t w = k, others[3] = {0u, 0u, 0u}, opos[3] = {0u, 0u, 0u}, spos = 0u;
            int no = 0;
#pragma unroll
            for (int j = 0; j < ND; j++) {
                int32_t dof = celldofs[(int64_t)cell * ND + j];
                int64_t p = find_slot(colptr, rowval, c, (int64_t)dof);
                uint32_t pos = (p < 0) ? ((ND == 3) ? 0xFFu : 0x3Fu) : (uint32_t)(p - base);
                if (ND == 4) w |= pos << (2 + 6 * j);
                if ((uint32_t)j != k) {
                    opos[no] = pos;
                    others[no++] = (uint32_t)(dof - 1);
                } else {
                    spos = pos;
                }
            }
            if (ND == 3) w |= (spos << 8) | (opos[0] << 16) | (opos[1] << 24);  // 2-D: positions in (self, a, b) order
            ww = w;
            oo[0] = others[0]; oo[1] = others[1]; oo[2] = others[2];
        }
        size_t idx = ((size_t)(row0 + r) << 5) + lane;
        gmap[idx] = oo[0];
        gmap[stride + idx] = oo[1];
        gmap[2 * stride + idx] = ww;
        gmap[3 * stride + idx] = cell;
        if (ND == 4) gmap[4 * stride + idx] = oo[2];
    }
}

int femb_build_gmap(femb_ctx* ctx) {
    femb_free(&ctx->gmap);
    femb_free(&ctx->gslice_ptr);
    ctx->gmap_rows = 0;
    ctx->gmap_stride = 0;
    ctx->gmap_space = -1;
    femb_free(&ctx->gvol);
    ctx->gvol_valid = false;
    ctx->chunk_n = 0;  // chunk plan of the streamed assembly depends on the gather map
    if (ctx->nrs != 1 || ctx->ncs != 1 || ctx->row_spaces[0] != ctx->col_spaces[0] || ctx->blocks.size() != 1) return FEMB_OK;
    int sid = ctx->row_spaces[0];
    const FembSpace& s = ctx->spaces[sid];
    if (s.family != FEMB_FAMILY_H1 || s.ncomp != 1 || s.nd != ctx->nnpc || s.handler != FEMB_HANDLER_NONE || s.ndofs != ctx->nnodes) return FEMB_OK;
    int maxpos = (s.nd == 3) ? 254 : 62;
    if (ctx->max_collen > maxpos) return FEMB_OK;  // fall back to the cell-parallel kernels
    const int64_t nslices = (s.ndofs + 31) / 32;
    int32_t* deg = nullptr;
    void* tmp = nullptr;
    {
        int st = femb_alloc(ctx, &deg, (size_t)nslices + 1);
        if (st == FEMB_OK) st = femb_alloc(ctx, &ctx->gslice_ptr, (size_t)nslices + 1);
        if (st == FEMB_OK && cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nslices + 1), ctx->stream) != cudaSuccess)
            st = femb_fail(ctx, FEMB_ERR_CUDA, "femb_build_gmap: memset failed");
        if (st != FEMB_OK) {
            cudaFree(deg);
            return st;
        }
    }
    k_slice_deg<<<(unsigned)((nslices + 127) / 128), 128, 0, ctx->stream>>>(s.adj_ptr, s.ndofs, nslices, deg);
    KERNEL_CHECK(ctx);
    // longest slice and total rows; if padding every slice to the longest costs < 10 % more rows, use that uniform
    // layout: row0 = slice * maxrows needs no load, which removes one dependent memory round trip from the kernel prologue
    int32_t rows = 0, maxrows = 0;
    {
        int32_t* d_red = nullptr;
        void* tmp2 = nullptr;
        size_t tb2 = 0, tb3 = 0;
        cudaError_t e2 = cudaMalloc(&d_red, 2 * sizeof(int32_t));
        if (e2 == cudaSuccess) e2 = cub::DeviceReduce::Max(nullptr, tb2, deg, d_red, nslices, ctx->stream);
        if (e2 == cudaSuccess) e2 = cub::DeviceReduce::Sum(nullptr, tb3, deg, d_red + 1, nslices, ctx->stream);
        if (e2 == cudaSuccess) e2 = cudaMalloc(&tmp2, std::max(tb2, tb3));
        if (e2 == cudaSuccess) e2 = cub::DeviceReduce::Max(tmp2, tb2, deg, d_red, nslices, ctx->stream);
        if (e2 == cudaSuccess) e2 = cub::DeviceReduce::Sum(tmp2, tb3, deg, d_red + 1, nslices, ctx->stream);
        int32_t h[2] = {0, 0};
        if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(h, d_red, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
        cudaFree(tmp2);
        cudaFree(d_red);
        ctx->launches += 2;
        if (e2 != cudaSuccess) {
            cudaFree(deg);
            return femb_fail(ctx, FEMB_ERR_CUDA, "gather-map reduction: %s", cudaGetErrorString(e2));
        }
        maxrows = h[0];
        rows = h[1];
    }
    ctx->gmap_uniform = (double)maxrows * (double)nslices <= 1.10 * (double)rows && (int64_t)maxrows * nslices < ((int64_t)1 << 31);
    if (ctx->gmap_uniform) {
        k_fill_i32<<<(unsigned)((nslices + 255) / 256), 256, 0, ctx->stream>>>(deg, nslices, maxrows);
        ctx->launches++;
        rows = (int32_t)((int64_t)maxrows * nslices);
    }
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, deg, ctx->gslice_ptr, nslices + 1, ctx->stream);
    cudaError_t e = cudaMalloc(&tmp, tb);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb, deg, ctx->gslice_ptr, nslices + 1, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    cudaFree(deg);
    ctx->launches += 1;
    CUDA_TRY(ctx, e);
    ctx->gmap_rows = rows;
    ctx->gmap_maxrows = maxrows;
    ctx->gmap_stride = (size_t)rows * 32;
    FEMB_TRY(femb_alloc(ctx, &ctx->gmap, ctx->gmap_stride * (s.nd == 4 ? 5 : 4)));
    unsigned grid = (unsigned)((nslices * 32 + 127) / 128);
    if (s.nd == 3)
        k_gmap<3><<<grid, 128, 0, ctx->stream>>>(s.adj, s.adj_ptr, s.ndofs, nslices, s.celldofs, ctx->colptr, ctx->rowval, ctx->gslice_ptr, ctx->gmap, ctx->gmap_stride);
    else
        k_gmap<4><<<grid, 128, 0, ctx->stream>>>(s.adj, s.adj_ptr, s.ndofs, nslices, s.celldofs, ctx->colptr, ctx->rowval, ctx->gslice_ptr, ctx->gmap, ctx->gmap_stride);
    KERNEL_CHECK(ctx);
    ctx->gmap_space = sid;
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// gather map for general scalar H1 spaces without per-cell handlers (P2 on triangles / tetrahedra, ...)
// ---------------------------------------------------------------------------------------------------
__global__ void k_slice_collen(const int64_t* __restrict__ colptr, int64_t ndofs, int64_t nslices, int32_t* __restrict__ maxlen) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nslices) return;
    int m = 0;
    for (int l = 0; l < 32; l++) {
        int64_t c = s * 32 + l;
        if (c < ndofs) m = max(m, (int)(colptr[c + 1] - colptr[c]));
    }
    maxlen[s] = m;
}

// Map entry of (column c, e-th adjacent cell), planes [row][lane]: plane 0 = cell, planes 1 .. nwords = positions of the
// cell's local rows inside CSC column c, 8 bits each (0xFF = not in the pattern).  The two spare bytes of the last word hold
// the local index k of c in the cell (bits 16-19, 0xF = padding entry) and one "first touch" flag per local row (bit 20 + j:
// no earlier visit of this column wrote that position, so the gather stores instead of adding and needs no zeroed
// accumulator).  needs_zero[slice] = 1 when a column of the slice has positions no cell contributes to (halo rows of a
// partitioned mesh, extra diagonal entries, rows of other components / blocks): those slices zero their accumulators first.
// Vector-valued (component-blocked) spaces: nd = ncomp * nds; the rows of a visit are the nds dofs of the column's own
// component, k is the index within the component; row_off / col_off place the block in the matrix.
__global__ void k_hmap(const uint32_t* __restrict__ adj, const int32_t* __restrict__ adj_ptr, int64_t ndofs, int64_t nslices, int nd, int nds, int nwords,
                       const int32_t* __restrict__ celldofs, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int64_t row_off, int64_t col_off,
                       const int32_t* __restrict__ slice_ptr, uint32_t* __restrict__ hmap, size_t stride, int32_t* __restrict__ nmissing,
                       uint8_t* __restrict__ needs_zero, int32_t* __restrict__ sub0, int32_t* __restrict__ sublen, int32_t* __restrict__ slice_sublen) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nslices * 32) return;
    const int64_t slice = c >> 5;
    const int lane = (int)(c & 31);
    const int row0 = slice_ptr[slice], nrows = slice_ptr[slice + 1] - row0;
    const int e0 = c < ndofs ? adj_ptr[c] : 0;
    const int deg = c < ndofs ? adj_ptr[c + 1] - e0 : 0;
    const int64_t gc = col_off + c;
    const int64_t base = c < ndofs ? colptr[gc] - 1 : 0;
    const int len = c < ndofs ? (int)(colptr[gc + 1] - 1 - base) : 0;
    uint32_t seen[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};  // positions written by earlier visits (max_collen <= 254)
    int nseen = 0;
    uint32_t pmin = 0xFFu, pmax = 0u;  // sub-range of the column the visits write to (sub0 != NULL: positions become relative to it)
    for (int r = 0; r < nrows; r++) {
        const size_t idx = ((size_t)(row0 + r) << 5) + lane;
        uint32_t cell = 0, k = 0xFu, first = 0u;
        uint32_t words[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        if (r < deg) {
            uint32_t a = adj[e0 + r];
            cell = a / (uint32_t)nd;
            const uint32_t kf = a - cell * (uint32_t)nd;
            const uint32_t comp = kf / (uint32_t)nds;
            k = kf - comp * (uint32_t)nds;
            for (int j = 0; j < nds; j++) {
                int64_t p = find_slot(colptr, rowval, gc, row_off + (int64_t)celldofs[(int64_t)cell * nd + comp * nds + j]);
                uint32_t pos = (p < 0) ? 0xFFu : (uint32_t)(p - base);
                if (p < 0) atomicAdd(nmissing, 1);  // rows without a slot (compressed / adopted patterns): the gather must check
                else {
                    pmin = min(pmin, pos);
                    pmax = max(pmax, pos);
                    if (!((seen[pos >> 5] >> (pos & 31)) & 1u)) {
                        seen[pos >> 5] |= 1u << (pos & 31);
                        first |= 1u << j;
                        nseen++;
                    }
                }
                words[j >> 2] = (words[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | (pos << (8 * (j & 3)));
            }
        }
        words[nwords - 1] = (words[nwords - 1] & 0xFFFFu) | (k << 16) | (first << 20);
        hmap[idx] = cell;
        for (int wd = 0; wd < nwords; wd++) hmap[(1 + wd) * stride + idx] = words[wd];
    }
    if (sub0) {
        // the block's rows of a column are one contiguous run of it (component-blocked dofs, rows ascending): the gather keeps
        // accumulators for, and writes, that run only
        const int sl = pmin <= pmax ? (int)(pmax - pmin + 1) : 0;
        if (c < ndofs) {
            sub0[c] = sl ? (int32_t)pmin : 0;
            sublen[c] = sl;
        }
        if (sl) {
            atomicMax(slice_sublen + slice, sl);
            if (pmin) {
                for (int r = 0; r < min(nrows, deg); r++) {
                    const size_t idx = ((size_t)(row0 + r) << 5) + lane;
                    for (int wd = 0; wd < nwords; wd++) {
                        uint32_t w = hmap[(1 + wd) * stride + idx];
                        const int nb = wd == nwords - 1 ? nds - 4 * wd : 4;
                        for (int b = 0; b < nb; b++) {
                            const uint32_t pos = (w >> (8 * b)) & 0xFFu;
                            if (pos != 0xFFu) w = (w & ~(0xFFu << (8 * b))) | ((pos - pmin) << (8 * b));
                        }
                        hmap[(1 + wd) * stride + idx] = w;
                    }
                }
            }
        }
        if (nseen < sl) needs_zero[slice] = 1;
        return;
    }
    if (nseen < len) needs_zero[slice] = 1;
}

void femb_free_hmap(FembHMap& H) {
    femb_free(&H.map);
    femb_free(&H.slice_ptr);
    femb_free(&H.slice_zero);
    femb_free(&H.sub0);
    femb_free(&H.sublen);
    H.stride = 0;
    H.space = -1;
    H.canon = false;
    H.ranges.clear();
    H.cranges.clear();
    for (auto& x : H.xlists) {
        femb_free(&x.ifc);
        femb_free(&x.rest);
    }
    H.xlists.clear();
    H.xlists_version = -1;
    femb_free(&H.vrow);
    femb_free(&H.vinv);
    femb_free(&H.vrec);
    femb_free(&H.xsrc);
    H.xsrc_state = 0;
    H.th_state = 0;
}

// in-place permutation of the position bytes and first-touch flags of every map entry: role order <-> local dof order
struct HPerm {
    uint8_t p[12][12];
};
__global__ void k_hmap_canon(uint32_t* __restrict__ hmap, size_t stride, int nds, int nwords, HPerm perm, int inverse) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= stride) return;
    uint32_t w[3] = {0u, 0u, 0u}, o[3] = {0u, 0u, 0u};
    for (int wd = 0; wd < nwords; wd++) w[wd] = hmap[(1 + wd) * stride + idx];
    const uint32_t k = (w[nwords - 1] >> 16) & 0xFu;
    if (k == 0xFu) return;  // padding entry
    const uint32_t first = w[nwords - 1] >> 20;
    uint32_t nfirst = 0u;
    for (int wd = 0; wd < nwords; wd++) o[wd] = w[wd];
    for (int r = 0; r < nds; r++) {
        const int j = perm.p[k][r];
        const int src = inverse ? r : j, dst = inverse ? j : r;
        const uint32_t b = (w[src >> 2] >> (8 * (src & 3))) & 0xFFu;
        o[dst >> 2] = (o[dst >> 2] & ~(0xFFu << (8 * (dst & 3)))) | (b << (8 * (dst & 3)));
        nfirst |= ((first >> src) & 1u) << dst;
    }
    o[nwords - 1] = (o[nwords - 1] & 0xFFFFu) | (k << 16) | (nfirst << 20);
    for (int wd = 0; wd < nwords; wd++) hmap[(1 + wd) * stride + idx] = o[wd];
}

int femb_hmap_set_canon(femb_ctx* ctx, FembHMap& H, bool canon) {
    if (!H.map || H.canon == canon) return FEMB_OK;
    const FembSpace& s = ctx->spaces[H.space];
    const int nds = s.nd / s.ncomp;
    if (!((ctx->dim == 3 && nds == 10 && H.words == 3) || (ctx->dim == 2 && nds == 6 && H.words == 2)))
        return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "role order exists for P2 on triangles / tetrahedra only");
    HPerm perm;
    femb_p2_role_perm(perm.p, ctx->dim);
    if (H.stride) {
        k_hmap_canon<<<(unsigned)((H.stride + 255) / 256), 256, 0, ctx->stream>>>(H.map, H.stride, nds, H.words, perm, canon ? 0 : 1);
        KERNEL_CHECK(ctx);
    }
    H.canon = canon;
    return FEMB_OK;
}

// builds the map of the (component-diagonal part of the) square block of space `sid` whose rows / columns start at row_off / col_off
static int build_hmap_into(femb_ctx* ctx, FembHMap& H, int sid, int64_t row_off, int64_t col_off, bool subrange) {
    const FembSpace& s = ctx->spaces[sid];
    const int nds = (s.family == FEMB_FAMILY_HDIV || s.family == FEMB_FAMILY_HCURL) ? s.nd : s.nd / s.ncomp;  // (Hdiv / Hcurl: vector-valued functions, one block of nd dofs)
    const int64_t nslices = (s.ndofs + 31) / 32;
    const int nwords = (nds + 5) / 4;  // the position bytes + two spare bytes (k, first-touch flags) in the last word
    int32_t *deg = nullptr, *clen = nullptr;
    void* tmp = nullptr;
    {
        int st = femb_alloc(ctx, &deg, (size_t)nslices + 1);
        if (st == FEMB_OK) st = femb_alloc(ctx, &clen, (size_t)nslices);
        if (st == FEMB_OK) st = femb_alloc(ctx, &H.slice_ptr, (size_t)nslices + 1);
        if (st == FEMB_OK && cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nslices + 1), ctx->stream) != cudaSuccess)
            st = femb_fail(ctx, FEMB_ERR_CUDA, "femb_build_hmap: memset failed");
        if (st != FEMB_OK) {
            cudaFree(deg);
            cudaFree(clen);
            return st;
        }
    }
    k_slice_deg<<<(unsigned)((nslices + 127) / 128), 128, 0, ctx->stream>>>(s.adj_ptr, s.ndofs, nslices, deg);
    k_slice_collen<<<(unsigned)((nslices + 127) / 128), 128, 0, ctx->stream>>>(ctx->colptr + col_off, s.ndofs, nslices, clen);
    ctx->launches += 2;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, deg, H.slice_ptr, nslices + 1, ctx->stream);
    cudaError_t e = cudaMalloc(&tmp, tb);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb, deg, H.slice_ptr, nslices + 1, ctx->stream);
    int32_t rows = 0;
    std::vector<int32_t> h_clen((size_t)nslices);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&rows, H.slice_ptr + nslices, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_clen.data(), clen, sizeof(int32_t) * nslices, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    cudaFree(deg);
    cudaFree(clen);
    ctx->launches += 2;
    CUDA_TRY(ctx, e);
    H.stride = (size_t)rows * 32;
    H.words = nwords;
    H.col_off = col_off;
    FEMB_TRY(femb_alloc(ctx, &H.map, H.stride * (1 + nwords)));
    FEMB_TRY(femb_alloc(ctx, &H.slice_zero, (size_t)nslices));
    CUDA_TRY(ctx, cudaMemsetAsync(H.slice_zero, 0, (size_t)nslices, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->errflag, 0, sizeof(int32_t), ctx->stream));
    int32_t* slice_sublen = nullptr;
    if (subrange) {
        FEMB_TRY(femb_alloc(ctx, &H.sub0, (size_t)s.ndofs));
        FEMB_TRY(femb_alloc(ctx, &H.sublen, (size_t)s.ndofs));
        FEMB_TRY(femb_alloc(ctx, &slice_sublen, (size_t)nslices));
        CUDA_TRY(ctx, cudaMemsetAsync(slice_sublen, 0, sizeof(int32_t) * nslices, ctx->stream));
    }
    k_hmap<<<(unsigned)((nslices * 32 + 127) / 128), 128, 0, ctx->stream>>>(s.adj, s.adj_ptr, s.ndofs, nslices, s.nd, nds, nwords, s.celldofs, ctx->colptr, ctx->rowval,
                                                                            row_off, col_off, H.slice_ptr, H.map, H.stride, ctx->errflag, H.slice_zero, H.sub0, H.sublen,
                                                                            slice_sublen);
    KERNEL_CHECK(ctx);
    if (subrange) {  // shared-memory sizing follows the run the gather keeps, not the whole column
        cudaError_t e2 = cudaMemcpyAsync(h_clen.data(), slice_sublen, sizeof(int32_t) * nslices, cudaMemcpyDeviceToHost, ctx->stream);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
        femb_free(&slice_sublen);
        CUDA_TRY(ctx, e2);
    }
    {
        int32_t nmiss = 0;
        CUDA_TRY(ctx, cudaMemcpyAsync(&nmiss, ctx->errflag, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->errflag, 0, sizeof(int32_t), ctx->stream));
        H.complete = nmiss == 0;
    }
    // contiguous slice ranges of similar column length (shared memory per thread = maxlen doubles): runs of equal
    // class ceil(len / 32), short runs (< 1/64 of the slices) merged into their predecessor, at most 16 ranges
    std::vector<FembHMap::Range> rg;
    const int64_t minrun = std::max<int64_t>(512, nslices / 64);
    for (int64_t sl = 0; sl < nslices; sl++) {
        const int cls = (h_clen[sl] + 31) / 32;
        if (!rg.empty() && ((rg.back().maxlen + 31) / 32 == cls || rg.back().s1 - rg.back().s0 < minrun)) {
            rg.back().s1 = sl + 1;
            rg.back().maxlen = std::max(rg.back().maxlen, h_clen[sl]);
        } else {
            rg.push_back({sl, sl + 1, h_clen[sl], 1});
        }
    }
    while (rg.size() > 16) {  // merge the neighbouring pair with the smallest size
        size_t best = 0;
        int64_t bs = INT64_MAX;
        for (size_t i = 0; i + 1 < rg.size(); i++) {
            int64_t sz = rg[i + 1].s1 - rg[i].s0;
            if (sz < bs) { bs = sz; best = i; }
        }
        rg[best].s1 = rg[best + 1].s1;
        rg[best].maxlen = std::max(rg[best].maxlen, rg[best + 1].maxlen);
        rg.erase(rg.begin() + best + 1);
    }
    // columns with many adjacent cells (vertex dofs of P2 tetrahedra: ~24) are split over 4 threads
    {
        std::vector<int32_t> h_ptr((size_t)nslices + 1);
        CUDA_TRY(ctx, cudaMemcpy(h_ptr.data(), H.slice_ptr, sizeof(int32_t) * (nslices + 1), cudaMemcpyDeviceToHost));
        for (auto& r : rg) r.split = ((double)(h_ptr[r.s1] - h_ptr[r.s0]) / (double)(r.s1 - r.s0) >= 12.0) ? 4 : 1;
        // neighbours with the same split whose merged accumulator block stays small (<= 24 KB per CTA) become one launch
        for (size_t i = 0; i + 1 < rg.size();) {
            const int ml = std::max(rg[i].maxlen, rg[i + 1].maxlen);
            if (rg[i].split == rg[i + 1].split && (size_t)ml * (128 / rg[i].split) * 8 <= 24 * 1024) {
                rg[i].s1 = rg[i + 1].s1;
                rg[i].maxlen = ml;
                rg.erase(rg.begin() + i + 1);
            } else {
                i++;
            }
        }
    }
    H.ranges = rg;
    H.space = sid;
    return FEMB_OK;
}

static bool hmap_space_ok(const femb_ctx* ctx, const FembSpace& s) {
    if (s.family == FEMB_FAMILY_HDIV)  // RT0 / BDM1 on tetrahedra (femb_hdiv.cu)
        return ctx->max_collen <= 254 && ((s.handler == FEMB_HANDLER_RT0_SIGNS && s.nd == s.nd_all && s.nd <= 4) || (s.handler == FEMB_HANDLER_BDM1_3D && s.nd == 12));
    if (s.family != FEMB_FAMILY_H1 || s.handler != FEMB_HANDLER_NONE || s.nd != s.nd_all || s.nd % s.ncomp || ctx->max_collen > 254) return false;
    const int nds = s.nd / s.ncomp;
    return nds <= 12 && ((nds & 3) == 1 || (nds & 3) == 2);  // the last position word must have two spare bytes (k + first-touch flags)
}

int femb_build_hmap(femb_ctx* ctx) {
    femb_free_hmap(ctx->hm);
    femb_free_hmap(ctx->hm_vec);
    femb_free(&ctx->rt0rec);  // (laid out by the entries of hm_vec)
    if (ctx->gmap_space >= 0) return FEMB_OK;  // scalar P1 has its own map
    if (ctx->nrs != 1 || ctx->ncs != 1 || ctx->row_spaces[0] != ctx->col_spaces[0] || ctx->blocks.size() != 1) return FEMB_OK;
    const int sid = ctx->row_spaces[0];
    const FembSpace& s = ctx->spaces[sid];
    if (s.ncomp != 1 || !hmap_space_ok(ctx, s)) return FEMB_OK;
    return build_hmap_into(ctx, ctx->hm, sid, 0, 0, false);
}

__global__ void k_check_twins(const int32_t* __restrict__ celldofs, int64_t ncells, int nd, int ncomp, int64_t n, const int32_t* __restrict__ sublen,
                              int32_t* __restrict__ bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int nds = nd / ncomp;
    bool ok = true;
    if (i < ncells)
        for (int comp = 1; comp < ncomp; comp++)
            for (int j = 0; j < nds; j++) ok = ok && celldofs[i * nd + comp * nds + j] == celldofs[i * nd + j] + comp * n;
    if (i < n)
        for (int comp = 1; comp < ncomp; comp++) ok = ok && sublen[i + comp * n] == sublen[i];
    if (!ok) atomicAdd(bad, 1);
}

// lazy: the (u,u) block of FEMatrix([FES_u, ...]) with a component-blocked vector-valued H1 space FES_u (Example210's Laplacian)
int femb_build_hmap_vec(femb_ctx* ctx) {
    if (ctx->hm_vec.map) return FEMB_OK;
    if (ctx->nrs < 1 || ctx->ncs < 1 || ctx->row_spaces[0] != ctx->col_spaces[0]) return FEMB_ERR_UNSUPPORTED;
    bool have = false;
    for (auto& b : ctx->blocks) have = have || (b.brow == 0 && b.bcol == 0);
    const int sid = ctx->row_spaces[0];
    const FembSpace& s = ctx->spaces[sid];
    if (!have || !hmap_space_ok(ctx, s)) return FEMB_ERR_UNSUPPORTED;
    FEMB_TRY(femb_build_adjacency(ctx, sid));
    FEMB_TRY(build_hmap_into(ctx, ctx->hm_vec, sid, ctx->row_off[0], ctx->col_off[0], true));
    // component twins: dof (comp, i) = comp * n + i in every cell and equally long runs -> the component blocks of a
    // component-wise form are copies of each other, a gather may compute one and write all (femb_p2.cu)
    FembHMap& H = ctx->hm_vec;
    H.twins = 1;
    if (s.family == FEMB_FAMILY_H1 && s.ncomp > 1 && s.ndofs % s.ncomp == 0 && ctx->ncells > 0) {
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->errflag, 0, sizeof(int32_t), ctx->stream));
        const int64_t n = s.ndofs / s.ncomp;
        k_check_twins<<<(unsigned)((std::max<int64_t>(ctx->ncells, n) + 255) / 256), 256, 0, ctx->stream>>>(s.celldofs, ctx->ncells, s.nd, s.ncomp, n, H.sublen, ctx->errflag);
        KERNEL_CHECK(ctx);
        int32_t bad = 0;
        CUDA_TRY(ctx, cudaMemcpyAsync(&bad, ctx->errflag, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->errflag, 0, sizeof(int32_t), ctx->stream));
        if (!bad) H.twins = s.ncomp;
    }
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// cell colouring (FEMB_SCATTER_COLORED): no two cells of one colour share a dof of any space in the matrix layout, so a
// colour can be scattered with plain read-modify-writes; colours are processed in order => bit-reproducible sums.
// Jones-Plassmann rounds on the device: an uncoloured cell takes the smallest colour not used by its neighbours as soon
// as it has the highest priority (hash of the cell id) among its uncoloured neighbours.  Deterministic: the result
// depends only on the mesh.
// ---------------------------------------------------------------------------------------------------
struct ColorSpaces {
    int n;
    const int32_t* celldofs[2 * FEMB_MAX_SPACES];
    const int32_t* adj_ptr[2 * FEMB_MAX_SPACES];
    const uint32_t* adj[2 * FEMB_MAX_SPACES];
    int nd[2 * FEMB_MAX_SPACES];
};

__device__ __forceinline__ uint32_t cell_prio(uint32_t c) {
    uint32_t h = c * 2654435761u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    return h;
}

__global__ void k_color_round(ColorSpaces sp, int64_t ncells, int32_t* __restrict__ color, int* __restrict__ remaining, int* __restrict__ overflow) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= ncells || color[c] >= 0) return;
    const uint32_t pc = cell_prio((uint32_t)c);
    unsigned long long used[4] = {0ull, 0ull, 0ull, 0ull};  // colours 0..255
    bool blocked = false;
    for (int s = 0; s < sp.n && !blocked; s++) {
        const int nd = sp.nd[s];
        for (int j = 0; j < nd && !blocked; j++) {
            const int32_t dof = sp.celldofs[s][c * nd + j] - 1;
            for (int32_t e = sp.adj_ptr[s][dof]; e < sp.adj_ptr[s][dof + 1]; e++) {
                const uint32_t n = sp.adj[s][e] / (uint32_t)nd;
                if (n == (uint32_t)c) continue;
                const int32_t cn = color[n];
                if (cn >= 0) {
                    if (cn < 256) used[cn >> 6] |= 1ull << (cn & 63);
                } else {
                    const uint32_t pn = cell_prio(n);
                    if (pn > pc || (pn == pc && n > (uint32_t)c)) {
                        blocked = true;
                        break;
                    }
                }
            }
        }
    }
    if (blocked) {
        atomicAdd(remaining, 1);
        return;
    }
    int col = -1;
    for (int w = 0; w < 4 && col < 0; w++)
        if (~used[w]) col = w * 64 + __ffsll((long long)~used[w]) - 1;
    if (col < 0) {
        atomicExch(overflow, 1);
        col = 255;
    }
    color[c] = col;
}

int femb_build_coloring(femb_ctx* ctx) {
    if (ctx->color_perm) return FEMB_OK;
    if (ctx->ncells == 0) {
        ctx->ncolors = 0;
        ctx->color_ptr.assign(1, 0);
        return FEMB_OK;
    }
    ColorSpaces sp;
    sp.n = 0;
    bool seen[FEMB_MAX_SPACES] = {false};
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < (pass ? ctx->ncs : ctx->nrs); i++) {
            const int sid = pass ? ctx->col_spaces[i] : ctx->row_spaces[i];
            if (seen[sid]) continue;
            seen[sid] = true;
            FEMB_TRY(femb_build_adjacency(ctx, sid));
            const FembSpace& s = ctx->spaces[sid];
            sp.celldofs[sp.n] = s.celldofs; sp.adj_ptr[sp.n] = s.adj_ptr; sp.adj[sp.n] = s.adj; sp.nd[sp.n] = s.nd;
            sp.n++;
        }
    int32_t *color = nullptr, *cells = nullptr, *color_sorted = nullptr;
    int* flags = nullptr;
    void* tmp = nullptr;
    auto cleanup = [&]() { cudaFree(color); cudaFree(cells); cudaFree(color_sorted); cudaFree(flags); cudaFree(tmp); };
#define TRY_K(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return femb_fail(ctx, FEMB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); } } while (0)
    const int64_t nc = ctx->ncells;
    TRY_K(cudaMalloc(&color, nc * sizeof(int32_t)));
    TRY_K(cudaMalloc(&flags, 2 * sizeof(int)));
    TRY_K(cudaMemsetAsync(color, 0xFF, nc * sizeof(int32_t), ctx->stream));
    TRY_K(cudaMemsetAsync(flags, 0, 2 * sizeof(int), ctx->stream));
    int h[2] = {1, 0};
    for (int round = 0; round < 4096 && h[0] > 0; round++) {
        TRY_K(cudaMemsetAsync(flags, 0, sizeof(int), ctx->stream));
        k_color_round<<<(unsigned)((nc + 127) / 128), 128, 0, ctx->stream>>>(sp, nc, color, flags, flags + 1);
        ctx->launches++;
        TRY_K(cudaMemcpyAsync(h, flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        TRY_K(cudaStreamSynchronize(ctx->stream));
    }
    if (h[0] > 0 || h[1]) {
        cleanup();
        return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "cell colouring needs more than 256 colours (or did not converge)");
    }
    // perm = cells sorted by colour (stable: ascending cell id inside a colour)
    TRY_K(cudaMalloc(&cells, nc * sizeof(int32_t)));
    TRY_K(cudaMalloc(&color_sorted, nc * sizeof(int32_t)));
    FEMB_TRY(femb_alloc(ctx, &ctx->color_perm, (size_t)nc));
    k_iota_i32<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(cells, nc);
    ctx->launches++;
    size_t tb = 0;
    TRY_K(cub::DeviceRadixSort::SortPairs(nullptr, tb, color, color_sorted, cells, ctx->color_perm, (int)nc, 0, 8, ctx->stream));
    TRY_K(cudaMalloc(&tmp, tb));
    TRY_K(cub::DeviceRadixSort::SortPairs(tmp, tb, color, color_sorted, cells, ctx->color_perm, (int)nc, 0, 8, ctx->stream));
    std::vector<int32_t> hc((size_t)nc);
    TRY_K(cudaMemcpyAsync(hc.data(), color_sorted, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    TRY_K(cudaStreamSynchronize(ctx->stream));
#undef TRY_K
    cleanup();
    ctx->ncolors = hc.back() + 1;
    ctx->color_ptr.assign(ctx->ncolors + 1, 0);
    for (int64_t i = 0; i < nc; i++) ctx->color_ptr[hc[i] + 1]++;
    for (int c = 0; c < ctx->ncolors; c++) ctx->color_ptr[c + 1] += ctx->color_ptr[c];
    return FEMB_OK;
}
