// Grid-topology side of the path: face / edge enumeration and the pattern-driven CellDofs generator
// on the device.  In the reference these are the lazily instantiated grid components xgrid[CellFaces], xgrid[FaceNodes],
// xgrid[CellFaceSigns], xgrid[CellFaceOrientations], xgrid[FaceCells], xgrid[CellEdges], xgrid[EdgeNodes],
// xgrid[CellEdgeSigns] (ExtendableGrids, read at /root/reference/src/dofmaps.jl:448-458, src/fedefs/hdiv_rt0.jl:115,
// src/fedefs/hdiv_bdm1.jl:201,235; timed separately at examples/Example200_LowLevelPoisson.jl:55) and
// fe_space[CellDofs] = init_dofmap_from_pattern! (/root/reference/src/dofmaps.jl:419-627, timed at Example200:59).
//
// Enumeration = first appearance while looping cells 1..ncells and local entities 1..k (the numbering the host mirror
// grids.py produces); here: radix sort of the (sorted node tuple, item) pairs, group heads = creating items, a prefix
// sum over the creator flags IN ITEM ORDER gives the appearance rank without a second sort.
#include <cub/cub.cuh>

#include "femb_internal.h"

namespace {

// local topology (0-based local node numbers), the tables quoted at /root/reference/src/fedefs/h1_p3.jl:167-168 and
// hdiv_bdm1.jl:143,156,169,182
__constant__ int8_t c_tri_faces[3][2] = {{0, 1}, {1, 2}, {2, 0}};
__constant__ int8_t c_tet_faces[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {0, 3, 2}};
__constant__ int8_t c_tet_edges[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

struct EntDesc {
    int dim, kind, per, k;  // kind 0 = faces, 1 = edges (3-D only; 2-D edges are the faces)
};

__device__ __forceinline__ void item_nodes(const EntDesc d, const int32_t* __restrict__ cellnodes, int64_t item, int32_t (&n)[3]) {
    const int64_t cell = item / d.per;
    const int l = (int)(item - cell * d.per);
    const int32_t* cn = cellnodes + cell * (d.dim + 1);
    if (d.dim == 2) {
        n[0] = cn[c_tri_faces[l][0]]; n[1] = cn[c_tri_faces[l][1]]; n[2] = 0;
    } else if (d.kind == 0) {
        n[0] = cn[c_tet_faces[l][0]]; n[1] = cn[c_tet_faces[l][1]]; n[2] = cn[c_tet_faces[l][2]];
    } else {
        n[0] = cn[c_tet_edges[l][0]]; n[1] = cn[c_tet_edges[l][1]]; n[2] = 0;
    }
}

__global__ void k_entity_keys(const EntDesc d, const int32_t* __restrict__ cellnodes, int64_t nitems, uint64_t* __restrict__ key_hi,
                              uint32_t* __restrict__ key_lo, uint32_t* __restrict__ idx) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    int32_t n[3];
    item_nodes(d, cellnodes, i, n);
    uint32_t a = (uint32_t)n[0], b = (uint32_t)n[1], c = (uint32_t)n[2];
    if (a > b) { uint32_t t = a; a = b; b = t; }
    if (d.k == 3) {
        if (b > c) { uint32_t t = b; b = c; c = t; }
        if (a > b) { uint32_t t = a; a = b; b = t; }
        key_lo[i] = c;
    }
    key_hi[i] = ((uint64_t)a << 32) | b;
    idx[i] = (uint32_t)i;
}

__global__ void k_gather_u64(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx, int64_t n, uint64_t* __restrict__ dst) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < n) dst[p] = src[idx[p]];
}

// group heads in sorted order: creator flag of the item (in item order) and the head position for the max-scan
__global__ void k_entity_heads(int k, const uint64_t* __restrict__ skey, const uint32_t* __restrict__ key_lo, const uint32_t* __restrict__ sidx,
                               int64_t n, int32_t* __restrict__ creator, uint32_t* __restrict__ headpos) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    bool head = p == 0 || skey[p] != skey[p - 1];
    if (!head && k == 3) head = key_lo[sidx[p]] != key_lo[sidx[p - 1]];
    creator[sidx[p]] = head ? 1 : 0;
    headpos[p] = head ? (uint32_t)p : 0u;
}

__global__ void k_entity_assign(const EntDesc d, const int32_t* __restrict__ cellnodes, int64_t n, const uint32_t* __restrict__ sidx,
                                const uint32_t* __restrict__ headpos, const int32_t* __restrict__ rank, int32_t* __restrict__ cell_ent,
                                int32_t* __restrict__ ent_nodes, int32_t* __restrict__ signs, int32_t* __restrict__ orient,
                                int32_t* __restrict__ facecells, int32_t* __restrict__ errflag) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t i = sidx[p], ci = sidx[headpos[p]];
    const int32_t id = rank[ci];
    const bool creator = i == ci;
    cell_ent[i] = id + 1;
    int32_t l[3], g[3];
    item_nodes(d, cellnodes, i, l);
    item_nodes(d, cellnodes, ci, g);
    if (creator)
        for (int j = 0; j < d.k; j++) ent_nodes[(int64_t)id * d.k + j] = l[j];
    const int32_t cell1 = (int32_t)(i / (uint32_t)d.per) + 1;
    if (d.kind == 0) {
        signs[i] = creator ? 1 : -1;
        facecells[2 * (int64_t)id + (creator ? 0 : 1)] = cell1;
        int o;
        if (d.dim == 2) {
            o = creator ? 1 : 2;
        } else {
            // 1: same sequence as the global face; the neighbour sees it reversed: 2: (l3,l2,l1), 3: (l2,l1,l3), 4: (l1,l3,l2)
            o = 0;
            if (g[0] == l[0] && g[1] == l[1] && g[2] == l[2]) o = 1;
            else if (g[0] == l[2] && g[1] == l[1] && g[2] == l[0]) o = 2;
            else if (g[0] == l[1] && g[1] == l[0] && g[2] == l[2]) o = 3;
            else if (g[0] == l[0] && g[1] == l[2] && g[2] == l[1]) o = 4;
            if (o == 0) atomicAdd(errflag, 1);  // same-handed neighbour faces: not an oriented manifold mesh
        }
        orient[i] = o;
    } else {
        signs[i] = (l[0] == g[0]) ? 1 : -1;
    }
}

void free_entities(FembEntities& e) {
    femb_free(&e.cell_ent);
    femb_free(&e.ent_nodes);
    femb_free(&e.signs);
    femb_free(&e.orient);
    femb_free(&e.facecells);
    e.built = false;
    e.n = 0;
}

int bits_for(uint64_t v) {
    int b = 1;
    while (b < 64 && (v >> b)) b++;
    return b;
}

}  // namespace

void femb_free_entities(femb_ctx* ctx) {
    for (auto& e : ctx->ent) free_entities(e);
}

int femb_build_entities(femb_ctx* ctx, int kind) {
    if (kind != 0 && kind != 1) return femb_fail(ctx, FEMB_ERR_ARG, "entity kind must be FEMB_ENTITY_FACES or FEMB_ENTITY_EDGES");
    if (!ctx->cellnodes) return femb_fail(ctx, FEMB_ERR_ARG, "femb_mesh_set must precede the entity enumeration");
    if (ctx->dim == 2) kind = 0;  // the edges of a triangle are its faces
    FembEntities& E = ctx->ent[kind];
    if (E.built) return FEMB_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    EntDesc d;
    d.dim = ctx->dim;
    d.kind = kind;
    d.per = ctx->dim == 2 ? 3 : (kind == 0 ? 4 : 6);
    d.k = (ctx->dim == 3 && kind == 0) ? 3 : 2;
    const int64_t n = ctx->ncells * d.per;
    if (n >= ((int64_t)1 << 31)) return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "too many (cell, local entity) items for 32-bit indices");
    free_entities(E);
    E.per_cell = d.per;
    E.nodes_per = d.k;
    if (n == 0) {
        E.built = true;
        return FEMB_OK;
    }
    uint64_t *key_hi = nullptr, *key_alt = nullptr;
    uint32_t *key_lo = nullptr, *lo_alt = nullptr, *idx = nullptr, *idx_alt = nullptr;
    int32_t *creator = nullptr, *rank = nullptr;
    void* tmp = nullptr;
    auto cleanup = [&]() {
        femb_free(&key_hi); femb_free(&key_alt); femb_free(&key_lo); femb_free(&lo_alt); femb_free(&idx); femb_free(&idx_alt);
        femb_free(&creator); femb_free(&rank);
        if (tmp) cudaFree(tmp);
        tmp = nullptr;
    };
#define GRID_TRY(expr)                      \
    do {                                    \
        int rc__ = (expr);                  \
        if (rc__ != FEMB_OK) {              \
            cleanup();                      \
            return rc__;                    \
        }                                   \
    } while (0)
#define GRID_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            cleanup();                                                                           \
            return femb_fail(ctx, FEMB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__));      \
        }                                                                                        \
    } while (0)
    GRID_TRY(femb_alloc(ctx, &key_hi, (size_t)n));
    GRID_TRY(femb_alloc(ctx, &key_alt, (size_t)n));
    GRID_TRY(femb_alloc(ctx, &idx, (size_t)n));
    GRID_TRY(femb_alloc(ctx, &idx_alt, (size_t)n));
    if (d.k == 3) {
        GRID_TRY(femb_alloc(ctx, &key_lo, (size_t)n));
        GRID_TRY(femb_alloc(ctx, &lo_alt, (size_t)n));
    }
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t st = ctx->stream;
    k_entity_keys<<<grid, 256, 0, st>>>(d, ctx->cellnodes, n, key_hi, key_lo, idx);
    ctx->launches++;
    const int nb = bits_for((uint64_t)ctx->nnodes);
    size_t tb = 0, tb2 = 0;
    cub::DoubleBuffer<uint64_t> dk(key_hi, key_alt);
    cub::DoubleBuffer<uint32_t> dv(idx, idx_alt);
    cub::DoubleBuffer<uint32_t> dl(key_lo, lo_alt);
    GRID_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, n, 0, 32 + nb, st));
    if (d.k == 3) GRID_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dl, dv, n, 0, nb, st));
    tb = std::max(tb, tb2);
    GRID_CUDA(cudaMalloc(&tmp, tb));
    const uint32_t* lo_by_item = key_lo;  // the 3rd node of each item, in item order (kept for the head test)
    if (d.k == 3) {
        // LSD over the three nodes: stable sort by the largest node, then by (smallest, middle)
        uint32_t* lo_copy = nullptr;
        GRID_TRY(femb_alloc(ctx, &lo_copy, (size_t)n));
        cudaError_t e = cudaMemcpyAsync(lo_copy, key_lo, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb, dl, dv, n, 0, nb, st);
        // the sorted copy of the low keys is no longer needed: keep the item-ordered copy in its place
        femb_free(&lo_alt);
        femb_free(&key_lo);
        key_lo = lo_copy;
        lo_by_item = key_lo;
        if (e != cudaSuccess) {
            cleanup();
            return femb_fail(ctx, FEMB_ERR_CUDA, "entity sort (pass 1): %s", cudaGetErrorString(e));
        }
        k_gather_u64<<<grid, 256, 0, st>>>(dk.Current(), dv.Current(), n, dk.Alternate());
        dk.selector ^= 1;
        ctx->launches += 2;
    }
    GRID_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, n, 0, 32 + nb, st));
    ctx->launches++;
    const uint64_t* skey = dk.Current();
    const uint32_t* sidx = dv.Current();
    GRID_TRY(femb_alloc(ctx, &creator, (size_t)n));
    GRID_TRY(femb_alloc(ctx, &rank, (size_t)n));
    // the alternate index buffer is free now: reuse it for the head positions
    uint32_t* headpos = dv.Alternate();
    k_entity_heads<<<grid, 256, 0, st>>>(d.k, skey, lo_by_item, sidx, n, creator, headpos);
    ctx->launches++;
    cudaFree(tmp);
    tmp = nullptr;
    tb = tb2 = 0;
    GRID_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, creator, rank, n, st));
    GRID_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb2, headpos, headpos, cub::Max(), n, st));
    tb = std::max(tb, tb2);
    GRID_CUDA(cudaMalloc(&tmp, tb));
    GRID_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, creator, rank, n, st));
    GRID_CUDA(cub::DeviceScan::InclusiveScan(tmp, tb, headpos, headpos, cub::Max(), n, st));
    ctx->launches += 2;
    int32_t last_rank = 0, last_flag = 0;
    GRID_CUDA(cudaMemcpyAsync(&last_rank, rank + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GRID_CUDA(cudaMemcpyAsync(&last_flag, creator + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GRID_CUDA(cudaStreamSynchronize(st));
    E.n = (int64_t)last_rank + last_flag;
    GRID_TRY(femb_alloc(ctx, &E.cell_ent, (size_t)n));
    GRID_TRY(femb_alloc(ctx, &E.ent_nodes, (size_t)E.n * d.k));
    GRID_TRY(femb_alloc(ctx, &E.signs, (size_t)n));
    if (kind == 0) {
        GRID_TRY(femb_alloc(ctx, &E.orient, (size_t)n));
        GRID_TRY(femb_alloc(ctx, &E.facecells, (size_t)E.n * 2));
        GRID_CUDA(cudaMemsetAsync(E.facecells, 0, sizeof(int32_t) * E.n * 2, st));
    }
    GRID_CUDA(cudaMemsetAsync(ctx->errflag, 0, sizeof(int32_t), st));
    k_entity_assign<<<grid, 256, 0, st>>>(d, ctx->cellnodes, n, sidx, headpos, rank, E.cell_ent, E.ent_nodes, E.signs, E.orient, E.facecells,
                                          ctx->errflag);
    ctx->launches++;
    int32_t nerr = 0;
    GRID_CUDA(cudaMemcpyAsync(&nerr, ctx->errflag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GRID_CUDA(cudaStreamSynchronize(st));
    GRID_CUDA(cudaGetLastError());
    cleanup();
#undef GRID_TRY
#undef GRID_CUDA
    if (nerr) {
        free_entities(E);
        return femb_fail(ctx, FEMB_ERR_ARG, "inconsistent face orientation on %d (cell, face) items (non-manifold or same-handed neighbour faces)", nerr);
    }
    E.built = true;
    return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------
// CellDofs from the dofmap pattern (dofmaps.jl:263-301 parser on the host side, :419-627 generator): segments
// (type, each_component, q); local order = component-major, within a component the segments in order; the
// not-each-component segments follow all components.
// ---------------------------------------------------------------------------------------------------
namespace {

struct PatSeg {
    int type, each, q;  // type: 0 node, 1 face, 2 edge, 3 interior / cell
};
struct PatArgs {
    int nseg, ncomp, nd, nnpc;
    PatSeg seg[FEMB_MAX_PATTERN_SEGS];
    int64_t count[4];      // nnodes, nfaces, nedges, ncells
    int per[4];            // items of each type per cell
    const int32_t* items[3];  // cellnodes, cellfaces, celledges
    int64_t coffset;
};

__global__ void k_celldofs_pattern(const PatArgs a, int64_t ncells, int32_t* __restrict__ out) {
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cell >= ncells) return;
    int32_t* o = out + cell * a.nd;
    int pos = 0;
    for (int pass = 0; pass <= a.ncomp; pass++) {
        const int each = pass < a.ncomp ? 1 : 0;
        int64_t offset = (int64_t)pass * a.coffset;  // the last pass starts at ncomp * coffset
        for (int s = 0; s < a.nseg; s++) {
            const PatSeg sg = a.seg[s];
            if (sg.each != each) continue;
            const int64_t nitems = a.count[sg.type];
            if (sg.type == 3) {
                for (int m = 0; m < sg.q; m++) o[pos++] = (int32_t)(cell + 1 + offset + m * nitems);
            } else {
                const int32_t* it = a.items[sg.type] + cell * a.per[sg.type];
                for (int i = 0; i < a.per[sg.type]; i++)
                    for (int m = 0; m < sg.q; m++) o[pos++] = (int32_t)(it[i] + offset + m * nitems);
            }
            offset += nitems * sg.q;
        }
    }
}

}  // namespace

extern "C" {

int32_t femb_grid_entities(femb_ctx* ctx, int32_t kind, int64_t* n_out) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(femb_build_entities(ctx, kind));
    if (n_out) *n_out = ctx->ent[ctx->dim == 2 ? 0 : kind].n;
    return FEMB_OK;
}

int32_t femb_grid_entities_get(femb_ctx* ctx, int32_t kind, int32_t* cell_entities, int32_t* entity_nodes, int32_t* signs, int32_t* orientations,
                               int32_t* facecells) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(femb_build_entities(ctx, kind));
    const FembEntities& E = ctx->ent[ctx->dim == 2 ? 0 : kind];
    const size_t n = (size_t)ctx->ncells * E.per_cell;
    if ((orientations || facecells) && !E.orient) return femb_fail(ctx, FEMB_ERR_ARG, "orientations / facecells exist for faces only");
    if (cell_entities && n) FEMB_TRY(femb_d2h(ctx, cell_entities, E.cell_ent, n));
    if (entity_nodes && E.n) FEMB_TRY(femb_d2h(ctx, entity_nodes, E.ent_nodes, (size_t)E.n * E.nodes_per));
    if (signs && n) FEMB_TRY(femb_d2h(ctx, signs, E.signs, n));
    if (orientations && n) FEMB_TRY(femb_d2h(ctx, orientations, E.orient, n));
    if (facecells && E.n) FEMB_TRY(femb_d2h(ctx, facecells, E.facecells, (size_t)E.n * 2));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FEMB_OK;
}

int32_t femb_space_set_from_pattern(femb_ctx* ctx, int32_t id, int32_t family, int32_t ncomp, int32_t nd, int32_t nd_all, int32_t handler,
                                    int32_t nseg, const int32_t* segs, int64_t* ndofs_out, int64_t* coffset_out) {
    FEMB_CHECK_CTX(ctx);
    if (id < 0 || id >= FEMB_MAX_SPACES) return femb_fail(ctx, FEMB_ERR_ARG, "space_id out of range");
    if (!ctx->cellnodes) return femb_fail(ctx, FEMB_ERR_ARG, "femb_mesh_set must precede femb_space_set_from_pattern");
    if (nd <= 0 || nd > 64 || nd_all < nd || ncomp <= 0 || ncomp > 3) return femb_fail(ctx, FEMB_ERR_ARG, "bad local dof counts");
    if (nseg <= 0 || nseg > FEMB_MAX_PATTERN_SEGS || !segs) return femb_fail(ctx, FEMB_ERR_ARG, "1..%d pattern segments expected", FEMB_MAX_PATTERN_SEGS);
    int nsign = 0, sign_kind = 0;
    switch (handler) {
        case FEMB_HANDLER_NONE: break;
        case FEMB_HANDLER_H1_FACESWAP2D: nsign = 3; break;
        case FEMB_HANDLER_H1_EDGESWAP3D: nsign = 6; sign_kind = 1; break;
        case FEMB_HANDLER_RT0_SIGNS: nsign = ctx->dim + 1; break;
        case FEMB_HANDLER_BDM1_2D: nsign = 3; break;
        case FEMB_HANDLER_BDM1_3D: nsign = 4; break;
        case FEMB_HANDLER_N0_EDGESIGNS3D: nsign = 6; sign_kind = 1; break;
        default: return femb_fail(ctx, FEMB_ERR_ARG, "unknown handler %d", handler);
    }
    if (handler == FEMB_HANDLER_N0_EDGESIGNS3D && ctx->dim != 3) return femb_fail(ctx, FEMB_ERR_ARG, "handler %d is the tetrahedral one (CellEdgeSigns)", handler);
    PatArgs a;
    memset(&a, 0, sizeof(a));
    a.nseg = nseg; a.ncomp = ncomp; a.nd = nd; a.nnpc = ctx->nnpc;
    bool need[2] = {nsign && sign_kind == 0, nsign && sign_kind == 1};
    for (int s = 0; s < nseg; s++) {
        a.seg[s].type = segs[3 * s]; a.seg[s].each = segs[3 * s + 1] ? 1 : 0; a.seg[s].q = segs[3 * s + 2];
        if (a.seg[s].type < 0 || a.seg[s].type > 3 || a.seg[s].q <= 0) return femb_fail(ctx, FEMB_ERR_ARG, "bad pattern segment %d", s);
        if (a.seg[s].type == 1) need[0] = true;
        if (a.seg[s].type == 2) need[1] = true;
    }
    if (ctx->dim == 2 && need[1]) { need[0] = true; need[1] = false; }
    for (int k = 0; k < 2; k++)
        if (need[k]) FEMB_TRY(femb_build_entities(ctx, k));
    const FembEntities& F = ctx->ent[0];
    const FembEntities& Ed = ctx->ent[ctx->dim == 2 ? 0 : 1];
    a.count[0] = ctx->nnodes; a.count[1] = F.n; a.count[2] = Ed.n; a.count[3] = ctx->ncells;
    a.per[0] = ctx->nnpc; a.per[1] = ctx->dim + 1; a.per[2] = ctx->dim == 2 ? 3 : 6; a.per[3] = 1;
    a.items[0] = ctx->cellnodes; a.items[1] = F.cell_ent; a.items[2] = Ed.cell_ent;
    // count_ndofs (finiteelements.jl:395-449)
    int64_t total = 0, coffset = 0;
    int ndcheck = 0;
    for (int s = 0; s < nseg; s++) {
        const int64_t c = a.count[a.seg[s].type] * a.seg[s].q;
        total += c * (a.seg[s].each ? ncomp : 1);
        if (a.seg[s].each) coffset += c;
        ndcheck += a.per[a.seg[s].type] * a.seg[s].q * (a.seg[s].each ? ncomp : 1);
    }
    if (ndcheck != nd) return femb_fail(ctx, FEMB_ERR_ARG, "the pattern yields %d local dofs, nd_local = %d", ndcheck, nd);
    if (total >= ((int64_t)1 << 31)) return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "ndofs exceeds the Int32 range of CellDofs");
    a.coffset = coffset;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FembSpace& s = ctx->spaces[id];
    femb_free_space(s);
    femb_free_pattern(ctx);
    FEMB_TRY(femb_alloc(ctx, &s.celldofs, (size_t)ctx->ncells * nd));
    s.celldofs_owned = true;
    const FembEntities& S = sign_kind == 0 ? F : Ed;
    if (nsign) {
        FEMB_TRY(femb_alloc(ctx, &s.signs, (size_t)ctx->ncells * nsign));
        CUDA_TRY(ctx, cudaMemcpyAsync(s.signs, S.signs, sizeof(int32_t) * ctx->ncells * nsign, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (handler == FEMB_HANDLER_BDM1_3D) {
        FEMB_TRY(femb_alloc(ctx, &s.orient, (size_t)ctx->ncells * 4));
        CUDA_TRY(ctx, cudaMemcpyAsync(s.orient, F.orient, sizeof(int32_t) * ctx->ncells * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (ctx->ncells) {
        k_celldofs_pattern<<<(unsigned)((ctx->ncells + 127) / 128), 128, 0, ctx->stream>>>(a, ctx->ncells, s.celldofs);
        KERNEL_CHECK(ctx);
    }
    s.family = family; s.ncomp = ncomp; s.ndofs = total; s.nd = nd; s.nd_all = nd_all; s.handler = handler; s.nsign = nsign;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    s.set = true;
    if (ndofs_out) *ndofs_out = total;
    if (coffset_out) *coffset_out = coffset;
    return FEMB_OK;
}

int32_t femb_space_get_celldofs(femb_ctx* ctx, int32_t id, int32_t* celldofs) {
    FEMB_CHECK_CTX(ctx);
    if (id < 0 || id >= FEMB_MAX_SPACES || !ctx->spaces[id].set || !celldofs) return femb_fail(ctx, FEMB_ERR_ARG, "unknown space");
    const FembSpace& s = ctx->spaces[id];
    if (ctx->ncells) FEMB_TRY(femb_d2h(ctx, celldofs, s.celldofs, (size_t)ctx->ncells * s.nd));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FEMB_OK;
}

}  // extern "C"
