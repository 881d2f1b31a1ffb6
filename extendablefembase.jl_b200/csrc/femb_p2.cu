// Scalar H1P2 on tetrahedra (C3 of BASELINE.json: Example200's assemble! with H1P2{1,3}) as an owner-computes gather in
// closed form.  With constant mu on an affine cell the quadrature sum of Example200:139-146 collapses to the reference-tensor
// contraction  A_jk = sum_ab G_ab T_ab[j,k],  G_ab = mu |T| grad(lambda_a).grad(lambda_b)  (the P1 stiffness of the cell), and
// for P2 (phi_i = lambda_i (2 lambda_i - 1), phi_ij = 4 lambda_i lambda_j, src/fedefs/h1_p2.jl) T is so sparse that a column
// of the local matrix needs 3 (vertex dof) or 5 (edge dof) of the six G_ab and 15-40 flops instead of the ~220 of the
// quadrature form.  The constants are exact integrals; the launcher accepts the closed form only if it reproduces the
// evaluator's own quadrature tables for every local column (else the general kernel of femb_h1.cu runs).
//
// What this buys (ncu of the general kernel: L1 data pipe 72 %, long-scoreboard stalls at 16 warps / SM, 124 registers):
//  * a vertex-column visit needs ONE 32-byte sector (G_k. and the local rhs entry) instead of a 96-byte record; the cell pre-pass
//    writes those sectors in VISIT order, so the vertex-column kernel streams them with the bulk copies of its map,
//  * an edge-column visit reads exactly two sectors of the cell's 128-byte edge line,
//  * no quadrature tables, no h_q intermediates: 75-95 registers,
//  * the gather map is put into "role order" (self / other vertices / incident edges / opposite edges) once, so the numeric
//    kernel has no per-lane indexing except one select chain over the local edge number,
//  * the slice's map arrives by bulk copies (cp.async.bulk + mbarrier): one memory latency for all visits of a slice.
// Timing experiments (compile-time): -DP2_DBG_REC (records forced into cache), -DP2_DBG_NOACC (no
// shared-memory accumulation), -DP2_DBG_NOSTORE (no write-out), -DP2_NO_PREFETCH, -DP2_FILT_ON=0.
#include <cstdlib>

#include "femb_kernels.cuh"

#include <cstdio>
// FEMB_P2_DEBUG=1: say why the closed-form path declined (it falls back silently otherwise)
static int p2_unsupported(int line) {
    static const bool dbg = [] { const char* e = getenv("FEMB_P2_DEBUG"); return e && e[0] == '1'; }();
    if (dbg) fprintf(stderr, "femb_p2.cu:%d: closed-form P2 path not taken\n", line);
    return FEMB_ERR_UNSUPPORTED;
}
#define P2_UNSUPPORTED p2_unsupported(__LINE__)

namespace {

#ifndef P2_FILT_ON
#define P2_FILT_ON 1
#endif
#ifndef P2_MINB
#define P2_MINB 4
#endif
#ifndef P2_DEPTH
#define P2_DEPTH 1
#endif
#ifndef P2_MINB0
#define P2_MINB0 4
#endif

// doubles per cell of the edge line (one aligned 128-byte line): [G01 G02 G03 G12] and three sectors [G13 G23 bl_a bl_b],
// (a b) = (4 5) (6 7) (8 9): edge dof 4 + e reads sectors 0 and 1 + e / 2.  The vertex sectors [G_k,o1 G_k,o2 G_k,o3 bl_k] live
// in the visit-order array FembHMap::vrec.
constexpr int P2_REC = 16;

struct P2Const {
    double alpha, beta, gd, go, mus, mud;
};

struct P2Wphi {
    double v[FEMB_COEF_STRIDE - 6][10];  // w_q phi_j(xref_q)
};

struct P2Args {
    const double* rec;
    const int32_t* slice_ptr;
    const uint8_t* slice_zero;
    const uint32_t* map;
    size_t stride;
    const int64_t* colptr;
    const int32_t *sub0, *sublen;
    int64_t ndofs;
    double* nzval;
    double* bvec;
    int acc_A, acc_b, has_rhs;
    int filt;         // value filter |v| <= drop_tol (Example200:150) and / or positions missing from a compressed pattern (0xFF -> dump row `maxlen`)
    double drop_tol;
    int32_t* errflag;
    int64_t slice_begin, slice_end;
    const double* vrec;         // 3-D: vertex-dof sectors [G_k,o1 G_k,o2 G_k,o3 bl_k] in VISIT order: entry (vrow[slice] + r) * 32 + lane
    const int32_t* vrow;        // per slice: first row of the slice in vrec, or -1 (no vertex dofs in the slice)
    const double* vrhs;         // 2-D vector maps: local rhs entries [cell][6][2] of a two-component load (fused rhs of Example210), or NULL
    const int32_t* slice_list;  // NULL, or the slices of this launch (overlapped interface exchange): slice_begin .. slice_end then index the list
    int twins;            // 2-D vector maps: the columns c, c + ndofs, ... (one per component) receive the same values
    int maxlen, maxrows;  // of the launch range: longest CSC column, most visits of a slice (shared-memory layout)
    P2Const c;
};

// local edge (a, b), a < b, of a tetrahedron -> 0..5 (ExtendableGrids local edge numbering, as femb_h1.cu p2_edge_a/b)
constexpr int edge_of(int a, int b) { return a == 0 ? b - 1 : (a == 1 ? b + 1 : 5); }

// role order of the rows of local column k (host + symbolic phase): vertex k -> [k, o1, o2, o3, (k o1), (k o2), (k o3),
// (o1 o2), (o1 o3), (o2 o3)] with o ascending; edge (i j), others k < l -> [i, j, k, l, (ij), (ik), (il), (jk), (jl), (kl)]
void p2_role_perm(uint8_t perm[10][10]) {
    auto E = [](int a, int b) { return 4 + (a < b ? edge_of(a, b) : edge_of(b, a)); };
    for (int k = 0; k < 4; k++) {
        int o[3], n = 0;
        for (int v = 0; v < 4; v++)
            if (v != k) o[n++] = v;
        const int r[10] = {k, o[0], o[1], o[2], E(k, o[0]), E(k, o[1]), E(k, o[2]), E(o[0], o[1]), E(o[0], o[2]), E(o[1], o[2])};
        for (int i = 0; i < 10; i++) perm[k][i] = (uint8_t)r[i];
    }
    const int ea[6] = {0, 0, 0, 1, 1, 2}, eb[6] = {1, 2, 3, 2, 3, 3};
    for (int e = 0; e < 6; e++) {
        const int i = ea[e], j = eb[e];
        int o[2], n = 0;
        for (int v = 0; v < 4; v++)
            if (v != i && v != j) o[n++] = v;
        const int r[10] = {i, j, o[0], o[1], E(i, j), E(i, o[0]), E(i, o[1]), E(j, o[0]), E(j, o[1]), E(o[0], o[1])};
        for (int t = 0; t < 10; t++) perm[4 + e][t] = (uint8_t)r[t];
    }
}

// the ten entries of local column k in role order from the six off-diagonal G (order 01 02 03 12 13 23): host mirror of the
// device formulas, used by the launch-time check against the quadrature tables
void p2_column_closed(const P2Const& c, const double G6[6], int k, double v[10]) {
    auto G = [&](int a, int b) { return G6[a < b ? edge_of(a, b) : edge_of(b, a)]; };
    if (k < 4) {
        int o[3], n = 0;
        for (int t = 0; t < 4; t++)
            if (t != k) o[n++] = t;
        const double g1 = G(k, o[0]), g2 = G(k, o[1]), g3 = G(k, o[2]), gkk = -((g1 + g2) + g3);
        v[0] = c.alpha * gkk;
        v[1] = c.beta * g1; v[2] = c.beta * g2; v[3] = c.beta * g3;
        v[4] = c.gd * g1 + c.go * gkk; v[5] = c.gd * g2 + c.go * gkk; v[6] = c.gd * g3 + c.go * gkk;
        v[7] = c.go * (g1 + g2); v[8] = c.go * (g1 + g3); v[9] = c.go * (g2 + g3);
    } else {
        const int ea[6] = {0, 0, 0, 1, 1, 2}, eb[6] = {1, 2, 3, 2, 3, 3};
        const int i = ea[k - 4], j = eb[k - 4];
        int o[2], n = 0;
        for (int t = 0; t < 4; t++)
            if (t != i && t != j) o[n++] = t;
        const double g0 = G(i, j), g1 = G(i, o[0]), g2 = G(i, o[1]), g3 = G(j, o[0]), g4 = G(j, o[1]);
        const double gii = -((g0 + g1) + g2), gjj = -((g0 + g3) + g4);
        v[0] = c.gd * g0 + c.go * gii; v[1] = c.gd * g0 + c.go * gjj;
        v[2] = c.go * (g1 + g3); v[3] = c.go * (g2 + g4);
        v[4] = c.mus * (gii + gjj) + 2.0 * c.mud * g0;
        v[5] = c.mus * g3 - c.mud * g2; v[6] = c.mus * g4 - c.mud * g1;
        v[7] = c.mus * g1 - c.mud * g4; v[8] = c.mus * g2 - c.mud * g3;
        v[9] = c.mud * ((g1 + g2) + (g3 + g4));
    }
}

__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// cell pre-pass: G (six off-diagonals of mu |T| grad lambda_a . grad lambda_b) and the local rhs vector, laid out so that
// a vertex-column visit needs one 32-byte sector and an edge-column visit two or three
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_cell_p2rec(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol,
                                                    int64_t ncells, double mu, RhsDev rhs, QuadDev q, P2Wphi wp, double* __restrict__ out, const uint32_t* __restrict__ vinv,
                                                    double* __restrict__ vrec) {
    __shared__ double tile[128][P2_REC + 1];  // staged so that the records leave the SM coalesced
    const int64_t cell0 = blockIdx.x * (int64_t)blockDim.x;
    const int64_t cell = cell0 + threadIdx.x;
    if (cell < ncells) {
        double x[4][3];
        load_nodes<3>(coords, cellnodes, cell, x);
        double A[3][3], C[3][3];
#pragma unroll
        for (int d = 0; d < 3; d++)
#pragma unroll
            for (int i = 0; i < 3; i++) A[d][i] = x[i + 1][d] - x[0][d];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int ra = (r + 1) % 3, rb = (r + 2) % 3, ja = (j + 1) % 3, jb = (j + 2) % 3;
                C[r][j] = A[ra][ja] * A[rb][jb] - A[ra][jb] * A[rb][ja];
            }
        const double det = A[0][0] * C[0][0] + A[0][1] * C[0][1] + A[0][2] * C[0][2];
        const double vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * (1.0 / 6.0);
        const double sc = (mu * vol) / (det * det);
        // K_pr = mu |T| grad lambda_{p+1} . grad lambda_{r+1}  (the columns of the cofactor matrix over det)
        double K[3][3];
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int r = p; r < 3; r++) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 3; k++) s += C[k][p] * C[k][r];
                K[p][r] = K[r][p] = s * sc;
            }
        const double G01 = -((K[0][0] + K[0][1]) + K[0][2]), G02 = -((K[0][1] + K[1][1]) + K[1][2]), G03 = -((K[0][2] + K[1][2]) + K[2][2]);
        const double G12 = K[0][1], G13 = K[0][2], G23 = K[1][2];
        double bl[10];
#pragma unroll
        for (int j = 0; j < 10; j++) bl[j] = 0.0;
        if (rhs.kind != FEMB_RHS_NONE) {
            for (int qp = 0; qp < q.nqp; qp++) {
                double f;
                if (rhs.kind == FEMB_RHS_AFFINE) {  // f(x_qp), x_qp = A xref_qp + x_1
                    f = rhs.p[0];
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        double sx = x[0][d];
#pragma unroll
                        for (int i = 0; i < 3; i++) sx += A[d][i] * q.xref[qp * 3 + i];
                        f += rhs.p[1 + d] * sx;
                    }
                } else {
                    f = __ldg(rhs.fvals + (size_t)cell * q.nqp + qp);
                }
                f *= vol;
#pragma unroll
                for (int j = 0; j < 10; j++) bl[j] += wp.v[qp][j] * f;
            }
        }
        // vertex sectors: straight to the slot of the (vertex dof, cell) visit in the gather map's order (the vertex-column
        // kernel then streams them with its bulk copies instead of gathering 32 bytes per lane and visit)
        stg256(vrec + (size_t)__ldg(vinv + cell * 4 + 0) * 4, G01, G02, G03, bl[0]);
        stg256(vrec + (size_t)__ldg(vinv + cell * 4 + 1) * 4, G01, G12, G13, bl[1]);
        stg256(vrec + (size_t)__ldg(vinv + cell * 4 + 2) * 4, G02, G12, G23, bl[2]);
        stg256(vrec + (size_t)__ldg(vinv + cell * 4 + 3) * 4, G03, G13, G23, bl[3]);
        double* o = tile[threadIdx.x];
        o[0] = G01; o[1] = G02; o[2] = G03; o[3] = G12;
        o[4] = G13; o[5] = G23; o[6] = bl[4]; o[7] = bl[5];
        o[8] = G13; o[9] = G23; o[10] = bl[6]; o[11] = bl[7];
        o[12] = G13; o[13] = G23; o[14] = bl[8]; o[15] = bl[9];
    }
    __syncthreads();
    const int64_t nloc = min((int64_t)128, ncells - cell0);
    double* dst = out + cell0 * P2_REC;
    for (int i = threadIdx.x; i < nloc * P2_REC; i += 128) dst[i] = tile[i / P2_REC][i % P2_REC];
}

// record pieces of a visit, loaded one visit ahead (MODE 1: edge dofs only, MODE 2: any; MODE 0 reads its vertex sectors from
// the staged visit-order block instead).  Edge dof 4 + e: r[0..3] = (G01 G02 G03 G12), r[4..7] = (G13 G23 bl_a bl_b) of its pair;
// vertex dof (MODE 2): r[0..3] = its sector from the visit-order array (vp = this lane's slot of the visit).  The wide loads are
// unconditional (padding entries point at cell 0): a conditionally written destination would need a copy that waits for the load.
template <int MODE>
__device__ __forceinline__ void p2_load_rec(const double* __restrict__ rec, unsigned cell, unsigned k, const double* __restrict__ vp, double (&r)[8]) {
#ifdef P2_DBG_REC
    cell &= 1023u;  // timing experiment: every record load hits the cache
#endif
    const double* p = rec + (size_t)cell * P2_REC;
    ldg256((MODE == 2 && k < 4u) ? vp : p, r[0], r[1], r[2], r[3]);
    ldg256(p + 4 + 4u * min(((k - 4u) >> 1) & 3u, 2u), r[4], r[5], r[6], r[7]);
}

// the ten entries of the visit's local column in role order (p2_column_closed) and the local rhs entry
template <int MODE>
__device__ __forceinline__ double p2_column(const P2Const& c, unsigned k, const double (&r)[8], double (&v)[10]) {
    if (MODE == 0 || (MODE == 2 && k < 4u)) {
        const double g1 = r[0], g2 = r[1], g3 = r[2], gkk = -((g1 + g2) + g3);
        v[0] = c.alpha * gkk;
        v[1] = c.beta * g1; v[2] = c.beta * g2; v[3] = c.beta * g3;
        v[4] = c.gd * g1 + c.go * gkk; v[5] = c.gd * g2 + c.go * gkk; v[6] = c.gd * g3 + c.go * gkk;
        v[7] = c.go * (g1 + g2); v[8] = c.go * (g1 + g3); v[9] = c.go * (g2 + g3);
        return r[3];
    }
    // edge e = (i j), others k < l: G_ij G_ik G_il G_jk G_jl picked from (01 02 03 12 13 23) = r[0..5] by select chains
    // (selp: the local edge number differs from lane to lane, branches would serialise the warp)
    const unsigned e = k - 4u;
    double g0, g1, g2, g3, g4, bl;
    asm("{\n\t.reg .pred p0, p1, p2, p3, p4, p5, po, l2, l3;\n\t.reg .f64 t0, t1, t2, t3;\n\t.reg .b32 tq;\n\t"
        "setp.eq.u32 p0, %6, 0;\n\tsetp.eq.u32 p1, %6, 1;\n\tsetp.eq.u32 p2, %6, 2;\n\tsetp.eq.u32 p3, %6, 3;\n\tsetp.eq.u32 p4, %6, 4;\n\tsetp.eq.u32 p5, %6, 5;\n\t"
        "setp.lt.u32 l2, %6, 2;\n\tsetp.lt.u32 l3, %6, 3;\n\tor.pred po, p0, p5;\n\t"
        "selp.f64 t0, %8, %9, p1;\n\tselp.f64 t0, %7, t0, p0;\n\tselp.f64 t1, %11, %12, p4;\n\tselp.f64 t1, %10, t1, p3;\n\tselp.f64 %0, t0, t1, l3;\n\t"
        "selp.f64 %1, %8, %7, po;\n\tselp.f64 %4, %11, %12, po;\n\t"
        "selp.f64 t2, %11, %10, p3;\n\tselp.f64 t2, %8, t2, p2;\n\tselp.f64 %2, %9, t2, l2;\n\t"
        "selp.f64 t3, %8, %9, p3;\n\tselp.f64 t3, %11, t3, p2;\n\tselp.f64 %3, %10, t3, l2;\n\t"
        "and.b32 tq, %6, 1;\n\tsetp.ne.u32 p1, tq, 0;\n\tselp.f64 %5, %14, %13, p1;\n\t}"
        : "=d"(g0), "=d"(g1), "=d"(g2), "=d"(g3), "=d"(g4), "=d"(bl)
        : "r"(e), "d"(r[0]), "d"(r[1]), "d"(r[2]), "d"(r[3]), "d"(r[4]), "d"(r[5]), "d"(r[6]), "d"(r[7]));
    const double gii = -((g0 + g1) + g2), gjj = -((g0 + g3) + g4);
    v[0] = c.gd * g0 + c.go * gii; v[1] = c.gd * g0 + c.go * gjj;
    v[2] = c.go * (g1 + g3); v[3] = c.go * (g2 + g4);
    v[4] = c.mus * (gii + gjj) + 2.0 * c.mud * g0;
    v[5] = c.mus * g3 - c.mud * g2; v[6] = c.mus * g4 - c.mud * g1;
    v[7] = c.mus * g1 - c.mud * g4; v[8] = c.mus * g2 - c.mud * g3;
    v[9] = c.mud * ((g1 + g2) + (g3 + g4));
    return bl;
}

__device__ __forceinline__ unsigned lds32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// ordered hand-over between the four warps of a SPLIT = 4 CTA: warp w waits on barrier 1 + w, its predecessor arrives there
__device__ __forceinline__ void p2_token_wait(unsigned warp) { asm volatile("bar.sync %0, 64;" ::"r"(1u + warp) : "memory"); }
__device__ __forceinline__ void p2_token_pass(unsigned warp) { asm volatile("bar.arrive %0, 64;" ::"r"(1u + ((warp + 1u) & 3u)) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// one visit: compute with the record in `cur` (loaded DEPTH visits ago), then start the load of the record of the visit
// DEPTH visits ahead into the same registers (the caller unrolls over a ring of DEPTH buffers: no register copies).  The
// map entries of all visits sit in shared memory (sst: this lane's column of the staged planes, pst: plane stride).
template <int SPLIT, int MODE, int DEPTH>
__device__ __forceinline__ void p2_visit(const P2Args& a, double (&cur)[8], int r, int nrows, unsigned sst, unsigned pst, unsigned srec, const double* __restrict__ vp0,
                                         unsigned sacc, unsigned fmask, unsigned warp, double& bsum, double (&racc)[3], unsigned (&rad)[3], unsigned& err) {
    constexpr unsigned ncol = 128 / SPLIT;
    // staged planes: MODE 0 keeps the three position words only (slots 0..2), the others the cell plane first
    const unsigned ra = sst + (unsigned)r * 128u;
    const unsigned wbase = MODE == 0 ? ra : ra + pst;
    const unsigned wa = lds32(wbase), wb = lds32(wbase + pst), wc = lds32(wbase + 2u * pst);
    const unsigned k = (wc >> 16) & 0xFu;
    const unsigned first = (wc >> 20) & fmask;
    const bool on = k != 0xFu;
    if (MODE == 0) {  // this lane's sector of the visit from the staged visit-order block
        const unsigned sr = srec + (unsigned)r * 1024u;
        lds128(sr, cur[0], cur[1]);
        lds128(sr + 16u, cur[2], cur[3]);
    }
    double v[10];
    unsigned ad[10];
#pragma unroll
    for (int j = 0; j < 10; j++) {
        const unsigned w = j < 4 ? wa : (j < 8 ? wb : wc);
        ad[j] = sacc + ((w >> (8 * (j & 3))) & 0xFFu) * (ncol * 8u);
    }
    if (on) {
        const double bl = p2_column<MODE>(a.c, k, cur, v);
        if (a.has_rhs) bsum += bl;
        if (P2_FILT_ON && a.filt) {  // (uniform) the reference's value filter; rows without a slot go to the dump row and must be filtered values
#pragma unroll
            for (int j = 0; j < 10; j++) {
                const unsigned w = j < 4 ? wa : (j < 8 ? wb : wc);
                const bool missing = ((w >> (8 * (j & 3))) & 0xFFu) == 0xFFu, big = fabs(v[j]) > a.drop_tol;
                if (missing) ad[j] = sacc + (unsigned)a.maxlen * (ncol * 8u);
                if (missing && big) err = 1u;
                v[j] = big ? v[j] : 0.0;
            }
        }
    }
    if (MODE != 0 && r + DEPTH * SPLIT < nrows) {  // (uniform in the warp)
        const unsigned rn = ra + (unsigned)(DEPTH * SPLIT) * 128u;
        p2_load_rec<MODE>(a.rec, lds32(rn), (lds32(rn + 3u * pst) >> 16) & 0xFu, vp0 + (size_t)(r + DEPTH * SPLIT) * 128, cur);
    }
#ifdef P2_DBG_NOACC
    if (on) {  // timing experiment: no shared-memory accumulation
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < 10; j++) t += v[j];
        bsum += t;
        if (t == 123.456) sts64(ad[0], t);
    }
    return;
#endif
    if (SPLIT == 1 && MODE == 1) {
        // edge column (i j): the rows of its two vertices and its own row get a contribution from EVERY visit -> they are
        // summed in registers (lower / higher global row: the CSC positions are ascending in the row, so comparing them
        // orders the endpoints) and stored once after the last visit; seven accumulations per visit go through shared memory
        if (on) {
            const bool sw = ad[0] > ad[1];
            racc[0] += sw ? v[1] : v[0];
            racc[1] += sw ? v[0] : v[1];
            racc[2] += v[4];
            rad[0] = sw ? ad[1] : ad[0];
            rad[1] = sw ? ad[0] : ad[1];
            rad[2] = ad[4];
            constexpr int J[7] = {2, 3, 5, 6, 7, 8, 9};
            double o[7];
#pragma unroll
            for (int t = 0; t < 7; t++) o[t] = lds64_unless(ad[J[t]], (first >> J[t]) & 1u);
#pragma unroll
            for (int t = 0; t < 7; t++) sts64(ad[J[t]], o[t] + v[J[t]]);
        }
    } else if (SPLIT == 1) {
        if (on) {
            double o[10];
#pragma unroll
            for (int j = 0; j < 10; j++) o[j] = lds64_unless(ad[j], (first >> j) & 1u);
#pragma unroll
            for (int j = 0; j < 10; j++) sts64(ad[j], o[j] + v[j]);
        }
    } else {
        // visit order r = 0, 1, 2, ... = warp 0, 1, 2, 3, 0, ...: a token travels round the four warps through named barriers
        // (producer: bar.arrive after its stores, consumer: bar.sync before its loads — two warps per barrier), so a warp only
        // waits for its predecessor's accumulation and computes its next column meanwhile (no CTA-wide barrier per visit)
        p2_token_wait(warp);
        if (on) {
            double o[10];
#pragma unroll
            for (int j = 0; j < 10; j++) o[j] = lds64_unless(ad[j], (first >> j) & 1u);
#pragma unroll
            for (int j = 0; j < 10; j++) sts64(ad[j], o[j] + v[j]);
        }
        p2_token_pass(warp);
    }
}

// SPLIT = 1: one thread per column, 4 slices (warps) per CTA.  SPLIT = 4: one slice per CTA, warp w takes the visits
// r = w (mod 4) and the shared-memory accumulation is serialised in visit order between barriers (same sums in the same
// order as one thread per column).  Requires a complete map in role order, no value filter, no slot tracking.
// Latency plan (the visit chain map entry -> record -> accumulate bounded the first versions, not a throughput): the slice's
// whole map comes in with bulk copies (cp.async.bulk + mbarrier, one memory latency for all visits) — for vertex-dof slices
// (MODE 0) together with the visit-order sectors, so that kernel issues no record loads at all; edge-dof slices prefetch all
// their records into L2 at once and load them one visit ahead.
// dynamic shared memory: [4 mbarriers | pad to 128 B][acc: (maxlen + 1) x ncol doubles][per slice: map planes (3 or 4) x maxrows x 128 B
// (+ MODE 0: maxrows x 1 KB of sectors)]
template <int SPLIT, int MODE>
__global__ void __launch_bounds__(128, MODE == 0 ? P2_MINB0 : P2_MINB) k_p2_gather(const P2Args a) {
    constexpr unsigned ncol = 128 / SPLIT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double bpart[SPLIT == 1 ? 1 : 4][32];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    int64_t slice = a.slice_begin + (SPLIT == 1 ? (int64_t)blockIdx.x * 4 + warp : (int64_t)blockIdx.x);
    if (SPLIT == 1 && slice >= a.slice_end) return;
    if (a.slice_list) slice = __ldg(a.slice_list + slice);
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned pst = (unsigned)a.maxrows * 128u;
    const unsigned mbar = smem0 + (SPLIT == 1 ? warp : 0u) * 8u;
    // per slice (warp with SPLIT = 1, CTA with SPLIT = 4): MODE 0 stages 3 map planes + the visit-order sectors (8 x pst),
    // the other modes the 4 map planes
    constexpr unsigned NPL = MODE == 0 ? 3u : 4u;
    constexpr unsigned STG = MODE == 0 ? 11u : 4u;  // staging block per slice in units of pst
    const unsigned sstw = smem0 + 128u + (unsigned)(a.maxlen + 1) * (ncol * 8u) + (SPLIT == 1 ? warp * STG * pst : 0u);
    const int64_t c = slice * 32 + lane;
    const bool valid = c < a.ndofs;
    const int row0 = __ldg(a.slice_ptr + slice), nrows = __ldg(a.slice_ptr + slice + 1) - row0;
    const int vrow = (MODE != 1) ? __ldg(a.vrow + slice) : -1;
    if (nrows > 0 && (SPLIT == 1 ? lane == 0 : tid == 0)) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned)nrows * 128u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes * (MODE == 0 ? 11u : 4u)) : "memory");
        const uint32_t* g = a.map + ((size_t)row0 << 5) + (MODE == 0 ? a.stride : 0);
#pragma unroll
        for (int p = 0; p < (int)NPL; p++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sstw + (unsigned)p * pst), "l"(g + (size_t)p * a.stride),
                         "r"(bytes), "r"(mbar)
                         : "memory");
        if (MODE == 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sstw + 3u * pst), "l"(a.vrec + (size_t)vrow * 128),
                         "r"(bytes * 8u), "r"(mbar)
                         : "memory");
    }
    const bool zero_first = __ldg(a.slice_zero + slice) != 0;
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
    const unsigned sacc = smem0 + 128u + (SPLIT == 1 ? tid : lane) * 8u;
    if (zero_first) {
        for (int i = (SPLIT == 1 ? 0 : (int)warp); i < len; i += SPLIT) sts64(sacc + i * (ncol * 8u), 0.0);
    }
    if (SPLIT > 1) __syncthreads();
    else __syncwarp();
    const unsigned fmask = zero_first ? 0u : 0xFFFFFFFFu;  // zeroed accumulators: every visit adds
    double bsum = 0.0;
    const int niter = (nrows + SPLIT - 1) / SPLIT;
    const int rfirst = (SPLIT == 1) ? 0 : (int)warp;
    const unsigned sst = sstw + lane * 4u;
    if (nrows > 0) {  // wait for the bulk copies (phase 0 of the barrier)
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
    }
#ifndef P2_NO_PREFETCH
    if (MODE != 0) {  // every edge record of this lane's later visits -> L2, now (with one record in flight per lane this is worth 8 %)
        for (int r = rfirst + SPLIT; r < nrows; r += SPLIT) {
            const unsigned k = (lds32(sst + 3u * pst + (unsigned)r * 128u) >> 16) & 0xFu;
            const double* p = a.rec + (size_t)lds32(sst + (unsigned)r * 128u) * P2_REC;
            if (k - 4u < 6u) {
                prefetch_l2(p);
                prefetch_l2(p + 4 + 4u * ((k - 4u) >> 1));
            }
        }
    }
#endif
    constexpr int DEPTH = MODE == 0 ? 1 : P2_DEPTH;  // records in flight per lane (MODE 0: none, its sectors are staged)
    const unsigned srec = sstw + 3u * pst + lane * 32u;                                                    // MODE 0: this lane's sector of row 0
    const double* vp0 = (MODE == 2 && vrow >= 0) ? a.vrec + ((size_t)vrow * 32 + lane) * 4 : a.rec;  // MODE 2: this lane's slot of row 0
    double rec[DEPTH][8];
#pragma unroll
    for (int u = 0; u < DEPTH; u++) {
#pragma unroll
        for (int i = 0; i < 8; i++) rec[u][i] = 0.0;
        const int r = rfirst + u * SPLIT;
        if (MODE != 0 && r < nrows)
            p2_load_rec<MODE>(a.rec, lds32(sst + (unsigned)r * 128u), (lds32(sst + 3u * pst + (unsigned)r * 128u) >> 16) & 0xFu, vp0 + (size_t)r * 128, rec[u]);
    }
    double racc[3] = {0.0, 0.0, 0.0};
    unsigned rad[3] = {0u, 0u, 0u}, err = 0u;
    if (SPLIT > 1 && warp == 3u) p2_token_pass(warp);  // the token starts at warp 0
    for (int it = 0; it < niter; it += DEPTH) {
#pragma unroll
        for (int u = 0; u < DEPTH; u++) {
            if (u == 0 || it + u < niter) {  // (uniform in the CTA)
                const int r = (it + u) * SPLIT + rfirst;
                // (SPLIT = 4: a warp whose row index is past nrows only takes part in the barriers)
                if (SPLIT == 1 || r < nrows) p2_visit<SPLIT, MODE, DEPTH>(a, rec[u], r, nrows, sst, pst, srec, vp0, sacc, fmask, warp, bsum, racc, rad, err);
                else {  // (SPLIT = 4: a warp whose row index is past nrows only hands the token on)
                    p2_token_wait(warp);
                    p2_token_pass(warp);
                }
            }
        }
    }
    if (SPLIT > 1 && warp == 0u) p2_token_wait(warp);  // take the token back: every visit has been accumulated
    if (err) atomicAdd(a.errflag, 1);
    if (SPLIT == 1 && MODE == 1 && rad[2] != 0u) {  // (a column with at least one visit)
#pragma unroll
        for (int t = 0; t < 3; t++) sts64(rad[t], racc[t]);
    }
    if (SPLIT > 1) {
        bpart[warp][lane] = bsum;
        __syncthreads();
        bsum = ((bpart[0][lane] + bpart[1][lane]) + bpart[2][lane]) + bpart[3][lane];
    }
#ifdef P2_DBG_NOSTORE
    if (bsum != 123.456) return;  // timing experiment: no write-out
#endif
    // write-out.  The 32 columns of the slice are one contiguous range of nzval; every aligned 32-byte sector of it is
    // stored by exactly one lane (st.global.v4.f64): the lane whose column holds the sector's first entry — it reads the
    // entries that belong to the next column from that column's accumulators.  Only the head of the slice's first column
    // and the tail of its last one are scalar stores (the neighbouring slice owns the rest of those sectors).
    if (SPLIT == 1) __syncwarp();  // (SPLIT = 4: the barrier above)
    const int len_next = __shfl_down_sync(0xFFFFFFFFu, len, 1);
    const int len_prev = __shfl_up_sync(0xFFFFFFFFu, len, 1);
    if (!valid) return;
    double* out = a.nzval + base;
    {
        const int w0 = (SPLIT == 1 ? 0 : (int)warp);
        const int head = min(len, (int)((4 - (base & 3)) & 3));  // entries before the first sector that starts in this column
        // the previous lane stores the head when it has a column of its own that ends where this one begins (a straddling
        // sector) and this column holds all the sector's remaining entries
        const bool head_by_prev = lane > 0 && len_prev > 0 && head > 0 && len >= head;
        if (!head_by_prev && w0 == 0) {
            for (int i = 0; i < head; i++) {
                const double t = lds64(sacc + i * (ncol * 8u));
                out[i] = a.acc_A ? out[i] + t : t;
            }
        }
        const int nsec = (len - head + 3) >> 2;  // sectors that start in this column (the last one may straddle)
        for (int g = w0; g < nsec; g += SPLIT) {
            const int i = head + 4 * g;
            const int own = min(4, len - i);  // entries of the sector in this column
            if (own < 4 && !(lane < 31 && len_next >= 4 - own)) {  // nobody to complete the sector with: scalar tail
                for (int u = 0; u < own; u++) {
                    const double t = lds64(sacc + (i + u) * (ncol * 8u));
                    out[i + u] = a.acc_A ? out[i + u] + t : t;
                }
                continue;
            }
            double t[4];
#pragma unroll
            for (int u = 0; u < 4; u++) t[u] = lds64(u < own ? sacc + (i + u) * (ncol * 8u) : sacc + 8u + (u - own) * (ncol * 8u));
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

// vinv[cell * 4 + k] = slot of the (vertex dof k of the cell, cell) visit in the visit-order array (one warp per slice)
__global__ void k_p2_vinv(const uint32_t* __restrict__ map, size_t stride, const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ vrow, int64_t nslices,
                          uint32_t* __restrict__ vinv) {
    const int64_t slice = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (slice >= nslices) return;
    const int v0 = vrow[slice];
    if (v0 < 0) return;
    const int row0 = slice_ptr[slice], nrows = slice_ptr[slice + 1] - row0;
    for (int r = 0; r < nrows; r++) {
        const size_t idx = ((size_t)(row0 + r) << 5) + lane;
        const unsigned k = (map[3 * stride + idx] >> 16) & 0xFu;
        if (k < 4u) vinv[(size_t)map[idx] * 4 + k] = (uint32_t)(((size_t)(v0 + r) << 5) + lane);
    }
}

// per slice: bit 0 = holds vertex-dof visits, bit 1 = holds edge-dof visits (one warp per slice)
__global__ void k_p2_slice_class(const uint32_t* __restrict__ kwords, const int32_t* __restrict__ slice_ptr, int64_t nslices, uint8_t* __restrict__ cls) {
    const int64_t slice = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (slice >= nslices) return;
    const int row0 = slice_ptr[slice], nrows = slice_ptr[slice + 1] - row0;
    unsigned bits = 0u;
    for (int r = 0; r < nrows; r++) {
        const unsigned k = (kwords[((size_t)(row0 + r) << 5) + lane] >> 16) & 0xFu;
        bits |= k < 4u ? 1u : (k < 10u ? 2u : 0u);
    }
    bits = __reduce_or_sync(0xFFFFFFFFu, bits);
    if (lane == 0) cls[slice] = (uint8_t)bits;
}

// ===================================================================================================
// P2 on triangles (scalar, or one component of a component-blocked vector space: the Laplacian of Example210:284-290,368).
// Same construction with three G (01 02 12): every visit — vertex or edge dof — needs ONE 32-byte sector that the pre-pass
// writes in role order (vertex k: [G_k,o1 G_k,o2 bl_k -], edge (i j), other k: [G_ij G_ik G_jk bl]), so the numeric kernel has
// no selects at all.  Local edges (h1_p2.jl / ExtendableGrids): e0 = (0 1), e1 = (1 2), e2 = (0 2).
// ===================================================================================================
constexpr int P2_REC2 = 24;  // doubles per cell: six sectors

constexpr int edge_of2(int a, int b) { return a == 0 ? (b == 1 ? 0 : 2) : 1; }

void p2_role_perm2(uint8_t perm[6][6]) {
    auto E = [](int a, int b) { return 3 + (a < b ? edge_of2(a, b) : edge_of2(b, a)); };
    for (int k = 0; k < 3; k++) {
        const int o1 = k == 0 ? 1 : 0, o2 = k == 2 ? 1 : 2;
        const int r[6] = {k, o1, o2, E(k, o1), E(k, o2), E(o1, o2)};
        for (int i = 0; i < 6; i++) perm[k][i] = (uint8_t)r[i];
    }
    const int ea[3] = {0, 1, 0}, eb[3] = {1, 2, 2};
    for (int e = 0; e < 3; e++) {
        const int i = ea[e], j = eb[e], k = 3 - i - j;
        const int r[6] = {i, j, k, E(i, j), E(i, k), E(j, k)};
        for (int t = 0; t < 6; t++) perm[3 + e][t] = (uint8_t)r[t];
    }
}

// host mirror of the device formulas (launch-time check against the quadrature tables); G3 by local edge = (G01 G12 G02)
void p2_column_closed2(const P2Const& c, const double G3[3], int k, double v[6]) {
    auto G = [&](int a, int b) { return G3[a < b ? edge_of2(a, b) : edge_of2(b, a)]; };
    if (k < 3) {
        const int o1 = k == 0 ? 1 : 0, o2 = k == 2 ? 1 : 2;
        const double g1 = G(k, o1), g2 = G(k, o2), gkk = -(g1 + g2);
        v[0] = c.alpha * gkk;
        v[1] = c.beta * g1; v[2] = c.beta * g2;
        v[3] = c.gd * g1 + c.go * gkk; v[4] = c.gd * g2 + c.go * gkk;
        v[5] = c.go * (g1 + g2);
    } else {
        const int ea[3] = {0, 1, 0}, eb[3] = {1, 2, 2};
        const int i = ea[k - 3], j = eb[k - 3], o = 3 - i - j;
        const double g0 = G(i, j), g1 = G(i, o), g3 = G(j, o), gii = -(g0 + g1), gjj = -(g0 + g3);
        v[0] = c.gd * g0 + c.go * gii; v[1] = c.gd * g0 + c.go * gjj;
        v[2] = c.go * (g1 + g3);
        v[3] = c.mus * (gii + gjj) + 2.0 * c.mud * g0;
        v[4] = c.mus * g3; v[5] = c.mus * g1;
    }
}

struct P2Wphi2 {
    double v[FEMB_COEF_STRIDE - 6][6];
};

__global__ void __launch_bounds__(128) k_cell_p2rec2d(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol,
                                                      int64_t ncells, double mu, RhsDev rhs, QuadDev q, P2Wphi2 wp, double* __restrict__ out, double* __restrict__ vrhs) {
    __shared__ double tile[128][P2_REC2 + 1];
    const int64_t cell0 = blockIdx.x * (int64_t)blockDim.x;
    const int64_t cell = cell0 + threadIdx.x;
    if (cell < ncells) {
        double x[3][2];
        load_nodes<2>(coords, cellnodes, cell, x);
        const double a00 = x[1][0] - x[0][0], a01 = x[2][0] - x[0][0], a10 = x[1][1] - x[0][1], a11 = x[2][1] - x[0][1];
        const double det = a00 * a11 - a01 * a10;
        const double vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * 0.5;
        const double sc = (mu * vol) / (det * det);
        // det * grad lambda_1 = (a11, -a01), det * grad lambda_2 = (-a10, a00)
        const double K00 = (a11 * a11 + a01 * a01) * sc, K11 = (a10 * a10 + a00 * a00) * sc, K01 = -(a11 * a10 + a01 * a00) * sc;
        const double G01 = -(K00 + K01), G02 = -(K01 + K11), G12 = K01;
        double bl[6];
#pragma unroll
        for (int j = 0; j < 6; j++) bl[j] = 0.0;
        if (vrhs) {
            // two-component load (Example210:304-318 with the component-blocked basis): b_(c,j) = |T| sum_q w_q phi_j(x_q) f_c(x_q)
            double b0[6], b1[6];
#pragma unroll
            for (int j = 0; j < 6; j++) b0[j] = b1[j] = 0.0;
            for (int qp = 0; qp < q.nqp; qp++) {
                const double xq[2] = {x[0][0] + a00 * q.xref[qp * 2] + a01 * q.xref[qp * 2 + 1], x[0][1] + a10 * q.xref[qp * 2] + a11 * q.xref[qp * 2 + 1]};
                double f[3];
                eval_rhs<2>(rhs, xq, cell, qp, q.nqp, f);
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    b0[j] += wp.v[qp][j] * f[0];
                    b1[j] += wp.v[qp][j] * f[1];
                }
            }
            double2* o2 = reinterpret_cast<double2*>(vrhs) + cell * 6;
#pragma unroll
            for (int j = 0; j < 6; j++) o2[j] = make_double2(b0[j] * vol, b1[j] * vol);
        } else if (rhs.kind != FEMB_RHS_NONE) {
            for (int qp = 0; qp < q.nqp; qp++) {
                double f;
                if (rhs.kind == FEMB_RHS_AFFINE) {
                    const double px = x[0][0] + a00 * q.xref[qp * 2] + a01 * q.xref[qp * 2 + 1], py = x[0][1] + a10 * q.xref[qp * 2] + a11 * q.xref[qp * 2 + 1];
                    f = rhs.p[0] + rhs.p[1] * px + rhs.p[2] * py;
                } else {
                    f = __ldg(rhs.fvals + (size_t)cell * q.nqp + qp);
                }
                f *= vol;
#pragma unroll
                for (int j = 0; j < 6; j++) bl[j] += wp.v[qp][j] * f;
            }
        }
        double* o = tile[threadIdx.x];
        o[0] = G01; o[1] = G02; o[2] = bl[0]; o[3] = 0.0;
        o[4] = G01; o[5] = G12; o[6] = bl[1]; o[7] = 0.0;
        o[8] = G02; o[9] = G12; o[10] = bl[2]; o[11] = 0.0;
        o[12] = G01; o[13] = G02; o[14] = G12; o[15] = bl[3];  // e0 = (0 1), other 2
        o[16] = G12; o[17] = G01; o[18] = G02; o[19] = bl[4];  // e1 = (1 2), other 0
        o[20] = G02; o[21] = G01; o[22] = G12; o[23] = bl[5];  // e2 = (0 2), other 1
    }
    __syncthreads();
    const int64_t nloc = min((int64_t)128, ncells - cell0);
    double* dst = out + cell0 * P2_REC2;
    for (int i = threadIdx.x; i < nloc * P2_REC2; i += 128) dst[i] = tile[i / P2_REC2][i % P2_REC2];
}

// one thread per column, 4 slices (warps) per CTA; map planes (cell, 2 position words) staged by bulk copies; records one
// visit ahead; the rows of an edge column that every visit touches (its two vertices, itself) are summed in registers.
// dynamic shared memory: [4 mbarriers | pad to 128 B][acc: maxlen x 128 doubles][4 warps x 3 planes x maxrows x 128 B]
__global__ void __launch_bounds__(128, 6) k_p2_gather2d(const P2Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int64_t slice = a.slice_begin + (int64_t)blockIdx.x * 4 + warp;
    if (slice >= a.slice_end) return;
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned pst = (unsigned)a.maxrows * 128u;
    const unsigned mbar = smem0 + warp * 8u;
    const unsigned sstw = smem0 + 128u + (unsigned)(a.maxlen + 1) * 1024u + warp * 3u * pst;
    const int64_t c = slice * 32 + lane;
    const bool valid = c < a.ndofs;
    const int row0 = __ldg(a.slice_ptr + slice), nrows = __ldg(a.slice_ptr + slice + 1) - row0;
    if (nrows > 0 && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned)nrows * 128u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes * 3u) : "memory");
        const uint32_t* g = a.map + ((size_t)row0 << 5);
#pragma unroll
        for (int p = 0; p < 3; p++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sstw + (unsigned)p * pst), "l"(g + (size_t)p * a.stride),
                         "r"(bytes), "r"(mbar)
                         : "memory");
    }
    const bool zero_first = __ldg(a.slice_zero + slice) != 0;
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
    const unsigned sacc = smem0 + 128u + tid * 8u;
    if (zero_first) {
        for (int i = 0; i < len; i++) sts64(sacc + i * 1024u, 0.0);
    }
    __syncwarp();
    const unsigned fmask = zero_first ? 0u : 0xFFFFFFFFu;
    double bsum = 0.0;
    const unsigned sst = sstw + lane * 4u;
    if (nrows > 0) {
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
    }
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
    double2 q1 = make_double2(0.0, 0.0);
    double vb0 = 0.0, vb1 = 0.0;
    if (nrows > 0) {
        const unsigned k0 = min((lds32(sst + 2u * pst) >> 16) & 0xFu, 5u);
        ldg256(a.rec + (size_t)lds32(sst) * P2_REC2 + 4u * k0, r0, r1, r2, r3);
        if (a.vrhs) q1 = __ldg(reinterpret_cast<const double2*>(a.vrhs) + (size_t)lds32(sst) * 6 + k0);
    }
    double racc[3] = {0.0, 0.0, 0.0};
    unsigned rad[3] = {0u, 0u, 0u}, err = 0u;
    for (int r = 0; r < nrows; r++) {
        const unsigned ra = sst + (unsigned)r * 128u;
        const unsigned wa = lds32(ra + pst), wb = lds32(ra + 2u * pst);
        const unsigned k = (wb >> 16) & 0xFu;
        const unsigned first = (wb >> 20) & fmask;
        const double c0 = r0, c1 = r1, c2 = r2, c3 = r3;
        const double2 qc = q1;
        if (r + 1 < nrows) {
            const unsigned k1 = min((lds32(ra + 128u + 2u * pst) >> 16) & 0xFu, 5u);
            ldg256(a.rec + (size_t)lds32(ra + 128u) * P2_REC2 + 4u * k1, r0, r1, r2, r3);
            if (a.vrhs) q1 = __ldg(reinterpret_cast<const double2*>(a.vrhs) + (size_t)lds32(ra + 128u) * 6 + k1);
        }
        if (k == 0xFu) continue;
        vb0 += qc.x;
        vb1 += qc.y;
        unsigned ad[6];
#pragma unroll
        for (int j = 0; j < 6; j++) ad[j] = sacc + (((j < 4 ? wa : wb) >> (8 * (j & 3))) & 0xFFu) * 1024u;
        double v[6];
        if (k < 3u) {  // vertex dof: (G_k,o1 G_k,o2 bl_k)
            const double gkk = -(c0 + c1);
            v[0] = a.c.alpha * gkk;
            v[1] = a.c.beta * c0; v[2] = a.c.beta * c1;
            v[3] = a.c.gd * c0 + a.c.go * gkk; v[4] = a.c.gd * c1 + a.c.go * gkk;
            v[5] = a.c.go * (c0 + c1);
            if (a.has_rhs) bsum += c2;
            if (a.filt) {
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    const bool missing = (((j < 4 ? wa : wb) >> (8 * (j & 3))) & 0xFFu) == 0xFFu, big = fabs(v[j]) > a.drop_tol;
                    if (missing) ad[j] = sacc + (unsigned)a.maxlen * 1024u;
                    if (missing && big) err = 1u;
                    v[j] = big ? v[j] : 0.0;
                }
            }
            double o[6];
#pragma unroll
            for (int j = 0; j < 6; j++) o[j] = lds64_unless(ad[j], (first >> j) & 1u);
#pragma unroll
            for (int j = 0; j < 6; j++) sts64(ad[j], o[j] + v[j]);
        } else {  // edge dof (i j), other vertex k: (G_ij G_ik G_jk bl)
            const double gii = -(c0 + c1), gjj = -(c0 + c2);
            v[0] = a.c.gd * c0 + a.c.go * gii; v[1] = a.c.gd * c0 + a.c.go * gjj;
            v[2] = a.c.go * (c1 + c2);
            v[3] = a.c.mus * (gii + gjj) + 2.0 * a.c.mud * c0;
            v[4] = a.c.mus * c2; v[5] = a.c.mus * c1;
            if (a.has_rhs) bsum += c3;
            if (a.filt) {
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    const bool missing = (((j < 4 ? wa : wb) >> (8 * (j & 3))) & 0xFFu) == 0xFFu, big = fabs(v[j]) > a.drop_tol;
                    if (missing) ad[j] = sacc + (unsigned)a.maxlen * 1024u;
                    if (missing && big) err = 1u;
                    v[j] = big ? v[j] : 0.0;
                }
            }
            const bool sw = ad[0] > ad[1];  // CSC positions ascend with the row: orders the two endpoints globally
            racc[0] += sw ? v[1] : v[0];
            racc[1] += sw ? v[0] : v[1];
            racc[2] += v[3];
            rad[0] = sw ? ad[1] : ad[0];
            rad[1] = sw ? ad[0] : ad[1];
            rad[2] = ad[3];
            constexpr int J[3] = {2, 4, 5};
            double o[3];
#pragma unroll
            for (int t = 0; t < 3; t++) o[t] = lds64_unless(ad[J[t]], (first >> J[t]) & 1u);
#pragma unroll
            for (int t = 0; t < 3; t++) sts64(ad[J[t]], o[t] + v[J[t]]);
        }
    }
    if (err) atomicAdd(a.errflag, 1);
    if (rad[2] != 0u) {
#pragma unroll
        for (int t = 0; t < 3; t++) sts64(rad[t], racc[t]);
    }
    __syncwarp();
    const int len_next = __shfl_down_sync(0xFFFFFFFFu, len, 1);
    const int len_prev = __shfl_up_sync(0xFFFFFFFFu, len, 1);
    if (!valid) return;
    for (int comp = 0; comp < a.twins; comp++) {
        if (comp > 0) base = a.colptr[c + comp * a.ndofs] - 1 + a.sub0[c + comp * a.ndofs];  // (twins: vector maps only)
        double* out = a.nzval + base;
        // every aligned 32-byte sector of the slice's nzval range is stored by the lane whose column holds its first entry
        const int head = min(len, (int)((4 - (base & 3)) & 3));
        const bool contiguous = a.sub0 == nullptr;  // (vector maps: the runs of neighbouring columns do not touch)
        const bool head_by_prev = contiguous && lane > 0 && len_prev > 0 && head > 0 && len >= head;
        if (!head_by_prev) {
            for (int i = 0; i < head; i++) {
                const double t = lds64(sacc + i * 1024u);
                out[i] = a.acc_A ? out[i] + t : t;
            }
        }
        const int nsec = (len - head + 3) >> 2;
        for (int g = 0; g < nsec; g++) {
            const int i = head + 4 * g;
            const int own = min(4, len - i);
            if (own < 4 && !(contiguous && lane < 31 && len_next >= 4 - own)) {
                for (int u = 0; u < own; u++) {
                    const double t = lds64(sacc + (i + u) * 1024u);
                    out[i + u] = a.acc_A ? out[i + u] + t : t;
                }
                continue;
            }
            double t[4];
#pragma unroll
            for (int u = 0; u < 4; u++) t[u] = lds64(u < own ? sacc + (i + u) * 1024u : sacc + 8u + (u - own) * 1024u);
            if (a.acc_A) {
                double o0, o1, o2, o3;
                asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o0), "=d"(o1), "=d"(o2), "=d"(o3) : "l"(out + i));
                t[0] += o0; t[1] += o1; t[2] += o2; t[3] += o3;
            }
            stg256(out + i, t[0], t[1], t[2], t[3]);
        }
    }
    if (a.has_rhs) {
        if (a.acc_b) a.bvec[c] += bsum;
        else a.bvec[c] = bsum;
    }
    if (a.vrhs) {  // (twins = 2: the scalar dof c carries both components)
        a.bvec[c] += vb0;
        a.bvec[c + a.ndofs] += vb1;
    }
}

}  // namespace

// role permutation of the gather map (femb_pattern.cu applies it in place)
void femb_p2_role_perm(uint8_t perm[12][12], int dim) {
    for (int k = 0; k < 12; k++)
        for (int r = 0; r < 12; r++) perm[k][r] = (uint8_t)r;
    if (dim == 3) {
        uint8_t p[10][10];
        p2_role_perm(p);
        for (int k = 0; k < 10; k++)
            for (int r = 0; r < 10; r++) perm[k][r] = p[k][r];
    } else {
        uint8_t p[6][6];
        p2_role_perm2(p);
        for (int k = 0; k < 6; k++)
            for (int r = 0; r < 6; r++) perm[k][r] = p[k][r];
    }
}

// P2 on triangles: H = the scalar map (Example200) or the component-diagonal map of a vector space (Example210's Laplacian:
// EI = nullptr, the result is added to nzval).  FEMB_ERR_UNSUPPORTED (no message) when the closed form does not apply.
static int launch_p2_gather2d(femb_ctx* ctx, FembHMap& H, const FembEval& EG, const FembEval* EI, double mu, const RhsDev& rhs, bool vec, bool vec_fresh = false,
                              bool vec_rhs = false, double drop_tol = 0.0) {
    const FembSpace& s = ctx->spaces[EG.space];
    const int nc = s.ncomp;
    if (EG.nqp > FEMB_COEF_STRIDE - 6 || EG.h_refderivs.size() != (size_t)EG.nqp * s.nd_all * nc * 2) return P2_UNSUPPORTED;
    if (EI && EI->h_refvals.size() != (size_t)EG.nqp * s.nd_all * nc) return P2_UNSUPPORTED;
    for (auto& r : H.ranges)
        if (r.split != 1) return P2_UNSUPPORTED;
    // exact integrals on a triangle: int lambda_a lambda_b = |T| / 12, int lambda_a^2 = |T| / 6
    const P2Const cst = {1.0, -1.0 / 3.0, 4.0 / 3.0, 0.0, 8.0 / 3.0, 4.0 / 3.0};
    {
        const double K[2][2] = {{1.3, -0.21}, {-0.21, 0.9}};
        const double G3[3] = {-(K[0][0] + K[0][1]), K[0][1], -(K[0][1] + K[1][1])};  // by local edge: (0 1), (1 2), (0 2)
        uint8_t perm[6][6];
        p2_role_perm2(perm);
        auto D = [&](int q, int j, int d) { return EG.h_refderivs[((size_t)q * s.nd_all + j) * nc * 2 + d]; };  // component 0 of the first 6 functions
        for (int k = 0; k < 6; k++) {
            double v[6];
            p2_column_closed2(cst, G3, k, v);
            for (int r = 0; r < 6; r++) {
                const int j = perm[k][r];
                double ref = 0.0;
                for (int q = 0; q < EG.nqp; q++)
                    for (int aa = 0; aa < 2; aa++)
                        for (int b = 0; b < 2; b++) ref += EG.w[q] * D(q, j, aa) * K[aa][b] * D(q, k, b);
                if (fabs(ref - v[r]) > 5e-12) {
                    if (getenv("FEMB_P2_DEBUG")) fprintf(stderr, "2-D check: k=%d r=%d j=%d ref=%.17g closed=%.17g nqp=%d nd_all=%d nc=%d\n", k, r, j, ref, v[r], EG.nqp, s.nd_all, nc);
                    return P2_UNSUPPORTED;
                }
            }
        }
    }
    if (!ctx->p2rec) FEMB_TRY(femb_alloc(ctx, &ctx->p2rec, (size_t)ctx->ncells * P2_REC2));
    if (ctx->ncells == 0) return FEMB_OK;
    FEMB_TRY(femb_hmap_set_canon(ctx, H, true));
    if (H.cranges.empty()) {
        const int64_t nslices = (s.ndofs + 31) / 32;
        std::vector<int32_t> hptr((size_t)nslices + 1);
        CUDA_TRY(ctx, cudaMemcpyAsync(hptr.data(), H.slice_ptr, sizeof(int32_t) * (nslices + 1), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        std::vector<FembHMap::Range> out = H.ranges;
        for (auto& c : out) {
            c.maxrows = 1;
            for (int64_t sl = c.s0; sl < c.s1; sl++) c.maxrows = std::max(c.maxrows, hptr[sl + 1] - hptr[sl]);
        }
        H.cranges = out;
    }
    auto smem_of = [](const FembHMap::Range& r) { return 128 + (size_t)(std::max(1, r.maxlen) + 1) * 1024 + 4 * 3 * (size_t)r.maxrows * 128; };
    size_t need = 0;
    for (auto& r : H.cranges) need = std::max(need, smem_of(r));
    if (need > 200 * 1024) return P2_UNSUPPORTED;
    P2Wphi2 wp;
    for (int q = 0; q < FEMB_COEF_STRIDE - 6; q++)
        for (int j = 0; j < 6; j++) wp.v[q][j] = (EI && q < EG.nqp) ? EG.w[q] * EI->h_refvals[((size_t)q * s.nd_all + j) * nc] : 0.0;
    QuadDev q = quad_dev(EG, 2);
    if (vec_rhs && !ctx->p2vrhs) FEMB_TRY(femb_alloc(ctx, &ctx->p2vrhs, (size_t)ctx->ncells * 12));
    k_cell_p2rec2d<<<(unsigned)((ctx->ncells + 127) / 128), 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, mu, rhs, q, wp, ctx->p2rec,
                                                                                 vec_rhs ? ctx->p2vrhs : nullptr);
    KERNEL_CHECK(ctx);
    P2Args a;
    a.rec = ctx->p2rec; a.slice_ptr = H.slice_ptr; a.slice_zero = H.slice_zero; a.map = H.map; a.stride = H.stride; a.colptr = ctx->colptr + H.col_off; a.sub0 = H.sub0; a.sublen = H.sublen; a.ndofs = s.ndofs;
    a.nzval = ctx->nzval; a.bvec = ctx->bvec;
    a.has_rhs = (!vec && rhs.kind != FEMB_RHS_NONE) ? 1 : 0;
    if (vec) {
        a.acc_A = vec_fresh ? 0 : 1;  // the columns also hold rows of other blocks: the caller flushed the lazy zero, only the block's runs are touched
        a.acc_b = 1;
    } else {
        FEMB_TRY(femb_settle_partial_zero(ctx));
        a.acc_A = ctx->zero_pending_A ? 0 : 1;
        a.acc_b = ctx->zero_pending_b ? 0 : 1;
        ctx->zero_pending_A = false;
        if (a.has_rhs) ctx->zero_pending_b = false;
    }
    a.filt = (!H.complete || drop_tol > 0.0) ? 1 : 0;
    a.drop_tol = drop_tol > 0.0 ? drop_tol : 0.0;
    a.errflag = ctx->errflag;
    a.slice_begin = a.slice_end = 0;
    a.slice_list = nullptr;
    a.vrhs = vec_rhs ? ctx->p2vrhs : nullptr;
    a.vrec = nullptr;
    a.vrow = nullptr;
    a.c = cst;
    a.twins = 1;
    int64_t slice_cap = INT64_MAX;
    if (vec && H.twins == s.ncomp && H.twins > 1) {  // compute component 0, write every component
        a.twins = H.twins;
        a.ndofs = s.ndofs / H.twins;
        slice_cap = (a.ndofs + 31) / 32;
    }
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_p2_gather2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_p2_gather2d, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (auto& r : H.cranges) {
        if (r.s1 <= r.s0) continue;
        a.slice_begin = r.s0;
        a.slice_end = std::min(r.s1, slice_cap);
        if (a.slice_end <= a.slice_begin) continue;
        a.maxlen = std::max(1, r.maxlen);
        a.maxrows = r.maxrows;
        k_p2_gather2d<<<(unsigned)((a.slice_end - a.slice_begin + 3) / 4), 128, smem_of(r), ctx->stream>>>(a);
        KERNEL_CHECK(ctx);
    }
    return FEMB_OK;
}

static bool p2_closed_enabled() {
    static const bool enabled = [] { const char* e = getenv("FEMB_P2_CLOSED"); return !(e && e[0] == '0'); }();
    return enabled;
}

// mu (grad u, grad v) of a component-blocked vector-valued P2 space on triangles in block (0,0) (Example210's Laplacian), added to nzval
int launch_p2_gather_vec(femb_ctx* ctx, int eg, double mu, int ei, const RhsDev* rhs, bool* rhs_done) {
    if (rhs_done) *rhs_done = false;
    if (!p2_closed_enabled()) return P2_UNSUPPORTED;
    const FembEval& EG = ctx->evals[eg];
    const FembSpace& s = ctx->spaces[EG.space];
    if (ctx->dim != 2 || s.ncomp < 2 || s.nd != 6 * s.ncomp || ctx->touched) return P2_UNSUPPORTED;
    int st = femb_build_hmap_vec(ctx);
    if (st != FEMB_OK) return st;
    FembHMap& H = ctx->hm_vec;
    if (H.space != EG.space || !H.complete) return P2_UNSUPPORTED;
    // right after femb_matrix_zero the runs can be overwritten instead of read, added to and written back
    FEMB_TRY(femb_settle_partial_zero(ctx));
    const bool fresh = ctx->zero_pending_A;
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    // the load of the same space rides along (Example210:304-318): its local entries come from the pre-pass, the gather sums them
    const bool with_rhs = rhs && rhs_done && rhs->kind != FEMB_RHS_NONE && rhs->ncomp == 2 && s.ncomp == 2 && H.twins == 2 && ei >= 0 && ctx->evals[ei].space == EG.space &&
                          ctx->evals[ei].op == FEMB_OP_IDENTITY && ctx->evals[ei].nqp == EG.nqp && ctx->bvec;
    if (with_rhs) {
        FEMB_TRY(femb_flush_zero(ctx, false, true));
        const int st2 = launch_p2_gather2d(ctx, H, EG, &ctx->evals[ei], mu, *rhs, true, fresh, true);
        if (st2 == FEMB_OK) *rhs_done = true;
        return st2;
    }
    RhsDev norhs;
    norhs.kind = FEMB_RHS_NONE;
    norhs.ncomp = 1;
    norhs.fvals = nullptr;
    return launch_p2_gather2d(ctx, H, EG, nullptr, mu, norhs, true, fresh);
}

// returns FEMB_ERR_UNSUPPORTED (no error message) when the closed form does not cover the space / rule / policy
int launch_p2_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol, const P2HostOut* host) {
    if (!p2_closed_enabled()) return P2_UNSUPPORTED;
    const FembEval& EG = ctx->evals[eg];
    const FembSpace& s = ctx->spaces[EG.space];
    FembHMap& H = ctx->hm;
    if (ctx->dim == 2) {
        if (s.nd != 6 || s.ncomp != 1 || !H.map || H.space != EG.space || ctx->touched) return P2_UNSUPPORTED;
        if (rhs.kind != FEMB_RHS_NONE && rhs.kind != FEMB_RHS_AFFINE && rhs.kind != FEMB_RHS_QPVALUES) return P2_UNSUPPORTED;
        if (host) return P2_UNSUPPORTED;  // (the streamed variant exists for tetrahedra)
        return launch_p2_gather2d(ctx, H, EG, rhs.kind != FEMB_RHS_NONE ? &ctx->evals[ei] : nullptr, mu, rhs, false, false, false, drop_tol);
    }
    if (ctx->dim != 3 || s.nd != 10 || s.ncomp != 1 || !H.map || H.space != EG.space || ctx->touched) return P2_UNSUPPORTED;
    if (EG.nqp > FEMB_COEF_STRIDE - 6 || EG.h_refderivs.size() != (size_t)EG.nqp * s.nd_all * 3) return P2_UNSUPPORTED;
    if (rhs.kind != FEMB_RHS_NONE && rhs.kind != FEMB_RHS_AFFINE && rhs.kind != FEMB_RHS_QPVALUES) return P2_UNSUPPORTED;
    const FembEval* EI = rhs.kind != FEMB_RHS_NONE ? &ctx->evals[ei] : nullptr;
    if (EI && EI->h_refvals.size() != (size_t)EG.nqp * s.nd_all) return P2_UNSUPPORTED;
    // exact integrals of the P2 shape-function products on a tetrahedron (int lambda_a lambda_b = |T| / 20, int lambda_a^2 = |T| / 10)
    const P2Const cst = {0.6, -0.2, 0.6, -0.2, 1.6, 0.8};
    {
        // accept the closed form only if it reproduces  sum_q w_q D_q[:,j]^T K D_q[:,k]  of the evaluator's tables (Example200:139-146)
        const double K[3][3] = {{1.3, -0.21, 0.17}, {-0.21, 0.9, 0.33}, {0.17, 0.33, 1.1}};
        const double G6[6] = {-((K[0][0] + K[0][1]) + K[0][2]), -((K[0][1] + K[1][1]) + K[1][2]), -((K[0][2] + K[1][2]) + K[2][2]), K[0][1], K[0][2], K[1][2]};
        uint8_t perm[10][10];
        p2_role_perm(perm);
        for (int k = 0; k < 10; k++) {
            double v[10];
            p2_column_closed(cst, G6, k, v);
            for (int r = 0; r < 10; r++) {
                const int j = perm[k][r];
                double ref = 0.0;
                for (int q = 0; q < EG.nqp; q++)
                    for (int aa = 0; aa < 3; aa++)
                        for (int b = 0; b < 3; b++)
                            ref += EG.w[q] * EG.h_refderivs[((size_t)q * s.nd_all + j) * 3 + aa] * K[aa][b] * EG.h_refderivs[((size_t)q * s.nd_all + k) * 3 + b];
                if (fabs(ref - v[r]) > 5e-12) return P2_UNSUPPORTED;  // (rules from an eigen-solver carry ~1e-14 in points and weights)
            }
        }
    }
    if (!ctx->p2rec) FEMB_TRY(femb_alloc(ctx, &ctx->p2rec, (size_t)ctx->ncells * P2_REC));
    if (ctx->ncells == 0) return FEMB_OK;
    FEMB_TRY(femb_hmap_set_canon(ctx, H, true));
    P2Args a;
    a.rec = ctx->p2rec; a.slice_ptr = H.slice_ptr; a.slice_zero = H.slice_zero; a.map = H.map; a.stride = H.stride; a.colptr = ctx->colptr; a.sub0 = a.sublen = nullptr; a.ndofs = s.ndofs;
    a.nzval = ctx->nzval; a.bvec = ctx->bvec;
    FEMB_TRY(femb_settle_partial_zero(ctx));
    a.acc_A = ctx->zero_pending_A ? 0 : 1;
    a.acc_b = ctx->zero_pending_b ? 0 : 1;
    a.has_rhs = rhs.kind != FEMB_RHS_NONE;
    a.filt = (!H.complete || drop_tol > 0.0) ? 1 : 0;
    a.drop_tol = drop_tol > 0.0 ? drop_tol : 0.0;
    a.errflag = ctx->errflag;
    a.slice_begin = a.slice_end = 0;
    a.slice_list = nullptr;
    a.vrhs = nullptr;
    a.c = cst;
    a.twins = 1;
    if (H.cranges.empty()) {
        // cut the launch ranges where the kind of dof changes (vertex dofs | the slice that holds both | edge dofs): the
        // kernels are specialised per kind (MODE), which keeps the vertex-dof kernel at one 32-byte record load per visit
        const int64_t nslices = (s.ndofs + 31) / 32;
        uint8_t* d_cls = nullptr;
        FEMB_TRY(femb_alloc(ctx, &d_cls, (size_t)nslices));
        k_p2_slice_class<<<(unsigned)((nslices * 32 + 127) / 128), 128, 0, ctx->stream>>>(H.map + 3 * H.stride, H.slice_ptr, nslices, d_cls);
        ctx->launches++;
        std::vector<uint8_t> cls((size_t)nslices);
        cudaError_t e = cudaMemcpyAsync(cls.data(), d_cls, (size_t)nslices, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        femb_free(&d_cls);
        CUDA_TRY(ctx, e);
        std::vector<int32_t> hptr((size_t)nslices + 1);
        CUDA_TRY(ctx, cudaMemcpy(hptr.data(), H.slice_ptr, sizeof(int32_t) * (nslices + 1), cudaMemcpyDeviceToHost));
        auto kind = [&](int64_t sl) { return cls[sl] == 1 ? 0 : (cls[sl] == 2 ? 1 : 2); };
        std::vector<FembHMap::Range> out;
        for (auto& r : H.ranges) {
            for (int64_t sl = r.s0; sl < r.s1;) {
                const int mode = kind(sl);
                int64_t e1 = sl + 1;
                while (e1 < r.s1 && kind(e1) == mode) e1++;
                FembHMap::Range c = r;
                c.s0 = sl; c.s1 = e1; c.mode = mode;
                out.push_back(c);
                sl = e1;
            }
        }
        if (out.size() > 64) {  // dof kinds interleaved: one launch per original range with the general variant
            out = H.ranges;
            for (auto& c : out) c.mode = 2;
        }
        for (auto& c : out) {
            c.maxrows = 1;
            for (int64_t sl = c.s0; sl < c.s1; sl++) c.maxrows = std::max(c.maxrows, hptr[sl + 1] - hptr[sl]);
        }
        H.cranges = out;
    }
    if (!H.vrow) {
        // visit-order slots of the vertex-dof sectors: per slice the first row in vrec (slices without vertex dofs: -1), and the
        // inverse map (cell, local vertex) -> slot for the cell pre-pass
        const int64_t nslices = (s.ndofs + 31) / 32;
        std::vector<int32_t> hptr((size_t)nslices + 1), hv((size_t)nslices, -1);
        CUDA_TRY(ctx, cudaMemcpy(hptr.data(), H.slice_ptr, sizeof(int32_t) * (nslices + 1), cudaMemcpyDeviceToHost));
        int64_t rows = 0;
        for (auto& r : H.cranges) {
            if (r.mode == 1) continue;
            for (int64_t sl = r.s0; sl < r.s1; sl++) {
                hv[sl] = (int32_t)rows;
                rows += hptr[sl + 1] - hptr[sl];
            }
        }
        if (rows >= ((int64_t)1 << 26)) return P2_UNSUPPORTED;  // (slots are 32-bit: rows * 32 < 2^31)
        FEMB_TRY(femb_alloc(ctx, &H.vrow, (size_t)nslices));
        FEMB_TRY(femb_h2d(ctx, H.vrow, hv.data(), (size_t)nslices));
        FEMB_TRY(femb_alloc(ctx, &H.vinv, (size_t)ctx->ncells * 4));
        FEMB_TRY(femb_alloc(ctx, &H.vrec, (size_t)std::max<int64_t>(rows, 1) * 128));
        CUDA_TRY(ctx, cudaMemsetAsync(H.vrec, 0, sizeof(double) * (size_t)std::max<int64_t>(rows, 1) * 128, ctx->stream));
        k_p2_vinv<<<(unsigned)((nslices * 32 + 127) / 128), 128, 0, ctx->stream>>>(H.map, H.stride, H.slice_ptr, H.vrow, nslices, H.vinv);
        KERNEL_CHECK(ctx);
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    P2Wphi wp;
    for (int q = 0; q < FEMB_COEF_STRIDE - 6; q++)
        for (int j = 0; j < 10; j++) wp.v[q][j] = (EI && q < EG.nqp) ? EG.w[q] * EI->h_refvals[(size_t)q * s.nd_all + j] : 0.0;
    QuadDev q = quad_dev(EG, 3);
    k_cell_p2rec<<<(unsigned)((ctx->ncells + 127) / 128), 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, mu, rhs, q, wp, ctx->p2rec, H.vinv, H.vrec);
    KERNEL_CHECK(ctx);
    a.vrec = H.vrec;
    a.vrow = H.vrow;
    {
        const void* kern[2][3] = {{(const void*)k_p2_gather<1, 0>, (const void*)k_p2_gather<1, 1>, (const void*)k_p2_gather<1, 2>},
                                  {(const void*)k_p2_gather<4, 0>, (const void*)k_p2_gather<4, 1>, (const void*)k_p2_gather<4, 2>}};
        auto smem_of = [](const FembHMap::Range& r) {
            const size_t ncol = r.split == 4 ? 32 : 128, nstage = r.split == 4 ? 1 : 4, planes = r.mode == 0 ? 11 : 4;  // (mode 0: 3 map planes + 8 of sectors)
            return 128 + (size_t)(std::max(1, r.maxlen) + 1) * ncol * sizeof(double) + nstage * planes * (size_t)r.maxrows * 128;
        };
        size_t need[2][3] = {{0, 0, 0}, {0, 0, 0}};
        for (auto& r : H.cranges) need[r.split == 4][r.mode] = std::max(need[r.split == 4][r.mode], smem_of(r));
        for (int sp = 0; sp < 2; sp++)
            for (int m = 0; m < 3; m++) {
                if (need[sp][m] > 200 * 1024) return P2_UNSUPPORTED;  // columns too long for the shared-memory accumulators: the general kernel decides
                if (!need[sp][m]) continue;
                CUDA_TRY(ctx, cudaFuncSetAttribute(kern[sp][m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need[sp][m]));
                CUDA_TRY(ctx, cudaFuncSetAttribute(kern[sp][m], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            }
        ctx->zero_pending_A = false;  // (from here on the call cannot decline any more)
        if (a.has_rhs) ctx->zero_pending_b = false;
        auto launch = [&](const FembHMap::Range& r, const int32_t* list, int64_t b0, int64_t b1, cudaStream_t stream) -> int {
            if (b1 <= b0) return FEMB_OK;
            a.slice_list = list;
            a.slice_begin = b0;
            a.slice_end = b1;
            a.maxlen = std::max(1, r.maxlen);
            a.maxrows = r.maxrows;
            void* params[1] = {(void*)&a};
            const unsigned grid = r.split == 4 ? (unsigned)(b1 - b0) : (unsigned)((b1 - b0 + 3) / 4);
            CUDA_TRY(ctx, cudaLaunchKernel(kern[r.split == 4][r.mode], dim3(grid), dim3(128), params, smem_of(r), stream));
            KERNEL_CHECK(ctx);
            return FEMB_OK;
        };
        const bool fused = !host && ctx->fused_exchange && ctx->comm && !ctx->peers.empty() && ctx->stream_x && !ctx->xslices.empty();
        if (host) {
            // streamed: every launch range is cut into chunks of slices; a finished chunk's nzval / b range goes to the host on
            // stream_d2h while the next chunk is assembled (the columns of a slice range are one contiguous range of nzval)
            const int64_t nsl_all = (s.ndofs + 31) / 32;
            int64_t total = 0;
            for (auto& r : H.cranges) total += r.s1 - r.s0;
            std::vector<std::pair<size_t, std::pair<int64_t, int64_t>>> chunks;  // (crange, [b0, b1))
            for (size_t i = 0; i < H.cranges.size(); i++) {
                const auto& r = H.cranges[i];
                const int64_t n = r.s1 - r.s0;
                if (n <= 0) continue;
                const int64_t nc = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)((double)std::max(1, host->nchunks) * (double)n / (double)std::max<int64_t>(1, total) + 0.5)));
                for (int64_t c = 0; c < nc; c++) chunks.push_back({i, {r.s0 + n * c / nc, r.s0 + n * (c + 1) / nc}});
            }
            // column pointers at the chunk boundaries
            std::vector<int64_t> cp(chunks.size() * 2);
            for (size_t c = 0; c < chunks.size(); c++) {
                const int64_t d0 = std::min<int64_t>(chunks[c].second.first * 32, s.ndofs), d1 = std::min<int64_t>(chunks[c].second.second * 32, s.ndofs);
                CUDA_TRY(ctx, cudaMemcpyAsync(&cp[2 * c], ctx->colptr + d0, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_TRY(ctx, cudaMemcpyAsync(&cp[2 * c + 1], ctx->colptr + d1, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            }
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            (void)nsl_all;
            while (ctx->chunk_ev.size() < chunks.size()) {
                cudaEvent_t ev;
                CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                ctx->chunk_ev.push_back(ev);
            }
            for (size_t c = 0; c < chunks.size(); c++) {
                const auto& r = H.cranges[chunks[c].first];
                const int64_t b0 = chunks[c].second.first, b1 = chunks[c].second.second;
                FEMB_TRY(launch(r, nullptr, b0, b1, ctx->stream));
                CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[c], ctx->stream));
                CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_d2h, ctx->chunk_ev[c], 0));
                const int64_t p0 = cp[2 * c] - 1, p1 = cp[2 * c + 1] - 1;
                if (p1 > p0) CUDA_TRY(ctx, cudaMemcpyAsync(host->nzval_out + p0, ctx->nzval + p0, sizeof(double) * (p1 - p0), cudaMemcpyDeviceToHost, ctx->stream_d2h));
                const int64_t d0 = std::min<int64_t>(b0 * 32, s.ndofs), d1 = std::min<int64_t>(b1 * 32, s.ndofs);
                if (a.has_rhs && host->b_out && d1 > d0)
                    CUDA_TRY(ctx, cudaMemcpyAsync(host->b_out + d0, ctx->bvec + d0, sizeof(double) * (d1 - d0), cudaMemcpyDeviceToHost, ctx->stream_d2h));
            }
        } else if (fused) {
            // overlapped interface exchange (as the P1 path): the slices that hold ghost or owned-interface columns are
            // assembled first on the high-priority stream, followed there by pack / NCCL / add, while the main stream
            // assembles all other slices.  Vertex and edge dofs are numbered apart, so the interface is a LIST of slices.
            if (H.xlists_version != ctx->exchange_version) {
                for (auto& x : H.xlists) {
                    femb_free(&x.ifc);
                    femb_free(&x.rest);
                }
                H.xlists.assign(H.cranges.size(), FembHMap::XList());
                for (size_t i = 0; i < H.cranges.size(); i++) {
                    const auto& r = H.cranges[i];
                    std::vector<int32_t> li, lr;
                    auto it = std::lower_bound(ctx->xslices.begin(), ctx->xslices.end(), r.s0);
                    for (int64_t sl = r.s0; sl < r.s1; sl++) {
                        if (it != ctx->xslices.end() && *it == sl) {
                            li.push_back((int32_t)sl);
                            ++it;
                        } else {
                            lr.push_back((int32_t)sl);
                        }
                    }
                    auto& x = H.xlists[i];
                    x.n_ifc = (int64_t)li.size();
                    x.n_rest = (int64_t)lr.size();
                    if (x.n_ifc) {
                        FEMB_TRY(femb_alloc(ctx, &x.ifc, li.size()));
                        FEMB_TRY(femb_h2d(ctx, x.ifc, li.data(), li.size()));
                    }
                    if (x.n_rest) {
                        FEMB_TRY(femb_alloc(ctx, &x.rest, lr.size()));
                        FEMB_TRY(femb_h2d(ctx, x.rest, lr.data(), lr.size()));
                    }
                }
                CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
                H.xlists_version = ctx->exchange_version;
            }
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_x0, ctx->stream));  // the records (and everything queued before) come first
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_x, ctx->ev_x0, 0));
            for (size_t i = 0; i < H.cranges.size(); i++) FEMB_TRY(launch(H.cranges[i], H.xlists[i].ifc, 0, H.xlists[i].n_ifc, ctx->stream_x));
            FEMB_TRY(femb_exchange_on(ctx, ctx->stream_x));
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_x1, ctx->stream_x));
            for (size_t i = 0; i < H.cranges.size(); i++) FEMB_TRY(launch(H.cranges[i], H.xlists[i].rest, 0, H.xlists[i].n_rest, ctx->stream));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_x1, 0));
            ctx->exchange_done = true;
        } else {
            // the vertex-column kernel (barriers of its 4-warp split, L1 67 %) and the edge-column kernel (L1 79 %) leave different
            // resources idle: run them side by side on two streams
            static const bool conc = [] { const char* e = getenv("FEMB_P2_CONCURRENT"); return !(e && e[0] == '0'); }();
            bool any0 = false, any1 = false;
            for (auto& r : H.cranges) (r.mode == 0 ? any0 : any1) = true;
            if (conc && any0 && any1) {
                if (!ctx->stream_aux) {
                    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_aux, cudaStreamNonBlocking));
                    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_a0, cudaEventDisableTiming));
                    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_a1, cudaEventDisableTiming));
                }
                CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a0, ctx->stream));
                CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_aux, ctx->ev_a0, 0));
                for (auto& r : H.cranges)
                    if (r.mode == 0) FEMB_TRY(launch(r, nullptr, r.s0, r.s1, ctx->stream_aux));
                CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a1, ctx->stream_aux));
                for (auto& r : H.cranges)
                    if (r.mode != 0) FEMB_TRY(launch(r, nullptr, r.s0, r.s1, ctx->stream));
                CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_a1, 0));
            } else {
                for (auto& r : H.cranges) FEMB_TRY(launch(r, nullptr, r.s0, r.s1, ctx->stream));
            }
        }
    }
    return FEMB_OK;
}
