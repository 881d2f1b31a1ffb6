// HDIVRT0 on tetrahedra (C4 of BASELINE.json: mass (v_j, v_k) + (div v_j, q) against L2P0) without the generic quadrature
// machinery.  The contravariant Piola image of the reference basis (src/fedefs/hdiv_rt0.jl:86-100, feevaluator_hdiv.jl:2-19)
// is  phi_j(x) = s_j (2 / det) (x - p_j),  p_j = the vertex opposite face j, s_j = xgrid[CellFaceSigns][j, cell], so
//   M_jk = s_j s_k (4 |T| / det^2) ( (c - p_j).(c - p_k) + (1/20) sum_i |x_i - c|^2 ),  c = centroid        (exact; the order-2
//   rule the caller uses integrates the same quadratic exactly), and  int div phi_j = s_j |T| trace / det  (= +-1).
// The launcher accepts the closed form only if it reproduces the evaluator's own tables on a test cell.
//  * mass: cell pre-pass writes the 4 x 4 local matrix as four 32-byte columns; one thread per face-dof column gathers its
//    (at most two) cells: write-once, bit-reproducible, one sector per visit (map staged by bulk copies as in femb_p2.cu);
//  * divergence blocks: every entry of (v, q) and (q, v) has exactly ONE contributing cell -> plain stores through the
//    slot maps, no reductions.
#include <cstdlib>

#include "femb_kernels.cuh"

namespace {

__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ unsigned lds32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// vertex opposite face j (0-based, hdiv_rt0.jl:86-100: the reference functions are 2 (xref - p_j) with p = e3, e2, 0, e1)
__host__ __device__ constexpr int rt0_opp(int j) { return j == 0 ? 3 : (j == 1 ? 2 : (j == 2 ? 0 : 1)); }

__global__ void __launch_bounds__(128) k_cell_rt0rec(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol,
                                                     const int32_t* __restrict__ signs, int64_t ncells, const uint32_t* __restrict__ finv, double* __restrict__ out,
                                                     const int32_t* __restrict__ dslots, double fdiv, double* __restrict__ nzval) {
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cell < ncells) {
        double x[4][3];
        load_nodes<3>(coords, cellnodes, cell, x);
        double A[3][3];
#pragma unroll
        for (int d = 0; d < 3; d++)
#pragma unroll
            for (int i = 0; i < 3; i++) A[d][i] = x[i + 1][d] - x[0][d];
        const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
        const double vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * (1.0 / 6.0);
        double c[3], S = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) c[d] = 0.25 * (((x[0][d] + x[1][d]) + x[2][d]) + x[3][d]);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int d = 0; d < 3; d++) S += (x[i][d] - c[d]) * (x[i][d] - c[d]);
        S *= 0.05;
        const double sc = 4.0 * vol / (det * det);
        double dv[4][3], sg[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            sg[j] = (double)__ldg(signs + cell * 4 + j);
#pragma unroll
            for (int d = 0; d < 3; d++) dv[j][d] = c[d] - x[rt0_opp(j)][d];
        }
        // column k of the local mass matrix goes straight to the slot of the (face dof, cell) visit in the gather map's order: the
        // gather streams its records with the bulk copies of the map instead of fetching one 32-byte sector per lane and visit
#pragma unroll
        for (int k = 0; k < 4; k++) {
            double m[4];
#pragma unroll
            for (int j = 0; j < 4; j++) m[j] = (sg[j] * sg[k]) * sc * (((dv[j][0] * dv[k][0] + dv[j][1] * dv[k][1]) + dv[j][2] * dv[k][2]) + S);
            double* o = out + (size_t)__ldg(finv + cell * 4 + k) * 4;
            asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(o), "d"(m[0]), "d"(m[1]), "d"(m[2]), "d"(m[3]) : "memory");
        }
        if (dslots) {
            // mixed pass: the (div v_j, q_T) entries of the cell's own column of block (0,1) — one sector, stored, not added (k_rt0_div's value)
#pragma unroll
            for (int j = 0; j < 4; j++) nzval[__ldg(dslots + cell * 4 + j)] = (fdiv * (sg[j] / det)) * vol;
        }
    }
}

// mixed pass: per visit (cell, local face k) of the gather map the slot of the cell's (0,1) entry and the position of its transposed copy
// in the (1,0) run that follows the mass run of the face column; bad counts what the fused gather does not cover
__global__ void k_hdiv_xsrc(const uint32_t* __restrict__ map, size_t stride, int nw, const int32_t* __restrict__ celldofs, const int64_t* __restrict__ colptr,
                            const int32_t* __restrict__ sub0, const int32_t* __restrict__ sublen, const int32_t* __restrict__ slots01, const int32_t* __restrict__ slots10,
                            uint32_t* __restrict__ xsrc, int32_t* __restrict__ bad) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= stride) return;
    const unsigned k = (map[(size_t)nw * stride + idx] >> 16) & 0xFu;
    uint32_t v = 0u;
    if (k < 4u) {
        const int64_t cell = map[idx];
        const int64_t col = celldofs[cell * 4 + k] - 1;
        const int32_t src = slots01[cell * 4 + k], dst = slots10[cell * 4 + k];
        const int64_t rank = (int64_t)dst - (colptr[col] - 1 + sublen[col]);
        const int64_t dlen = (colptr[col + 1] - colptr[col]) - sublen[col];
        if (src < 0 || dst < 0 || sub0[col] != 0 || rank < 0 || rank > 1 || rank >= dlen || dlen > 2) atomicAdd(bad, 1);
        else v = (uint32_t)src | ((uint32_t)rank << 31);
    }
    xsrc[idx] = v;
}

// finv[cell * nd + k] = entry (row * 32 + lane) of the (local dof k of the cell, cell) visit in the gather map (nw position words)
__global__ void k_hdiv_finv(const uint32_t* __restrict__ map, size_t stride, int nd, int nw, uint32_t* __restrict__ finv) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= stride) return;
    const unsigned k = (map[(size_t)nw * stride + idx] >> 16) & 0xFu;
    if (k < (unsigned)nd) finv[(size_t)map[idx] * nd + k] = (uint32_t)idx;
}

struct RGArgs {
    const double* rec;
    const int32_t* slice_ptr;
    const uint8_t* slice_zero;
    const uint32_t* map;
    size_t stride;
    const int64_t* colptr;
    const int32_t *sub0, *sublen;
    int64_t ndofs;
    double* nzval;
    double factor;
    int acc_A;
    int64_t slice_begin, slice_end;
    int maxlen, maxrows;
    const uint32_t* xsrc;  // MIX
};

// one thread per face-dof column, 4 slices (warps) per CTA; ND local dofs per cell (RT0: 4, BDM1: 12).  Map planes (NW position
// words; the cell plane is not needed) staged by bulk copies together with the visit-order records (ND doubles per visit).
// dynamic shared memory: [4 mbarriers | pad to 128 B][acc: maxlen x 128 doubles][4 warps x (NW + 2 ND) x maxrows x 128 B]
// MIX (RT0 in a [v; q] layout, fresh matrix): the thread also owns the (1,0) run that follows the mass run of its column — the transposed
// copies of the (div v, q) entries the pre-pass has just stored in block (0,1), fetched through one more staged plane — and writes the
// whole column in one piece (maxlen counts the two extra accumulator rows).
template <int ND, bool MIX = false>
__global__ void __launch_bounds__(128, ND == 4 ? 8 : 4) k_hdiv_mass_gather(const RGArgs a) {
    constexpr unsigned NW = (ND + 5) / 4, STG = NW + 2u * ND, STGX = STG + (MIX ? 1u : 0u);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int64_t slice = a.slice_begin + (int64_t)blockIdx.x * 4 + warp;
    if (slice >= a.slice_end) return;
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned pst = (unsigned)a.maxrows * 128u;
    const unsigned mbar = smem0 + warp * 8u;
    const unsigned sstw = smem0 + 128u + (unsigned)a.maxlen * 1024u + warp * STGX * pst;
    const int64_t c = slice * 32 + lane;
    const bool valid = c < a.ndofs;
    const int row0 = __ldg(a.slice_ptr + slice), nrows = __ldg(a.slice_ptr + slice + 1) - row0;
    if (nrows > 0 && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned)nrows * 128u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes * STGX) : "memory");
        const uint32_t* g = a.map + ((size_t)row0 << 5) + a.stride;
        if (MIX)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sstw + STG * pst), "l"(a.xsrc + ((size_t)row0 << 5)),
                         "r"(bytes), "r"(mbar)
                         : "memory");
#pragma unroll
        for (int p = 0; p < (int)NW; p++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sstw + (unsigned)p * pst), "l"(g + (size_t)p * a.stride),
                         "r"(bytes), "r"(mbar)
                         : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sstw + NW * pst), "l"(a.rec + (size_t)row0 * 32 * ND),
                     "r"(bytes * 2u * ND), "r"(mbar)
                     : "memory");
    }
    const bool zero_first = __ldg(a.slice_zero + slice) != 0;
    int64_t base = 0;
    int len = 0, dlen = 0;
    if (valid) {
        base = a.colptr[c] - 1;
        len = (int)(a.colptr[c + 1] - 1 - base);
        if (a.sub0) {  // the block's run of the column
            if (MIX) dlen = len - a.sublen[c];  // (sub0 == 0, checked when xsrc was built: the (1,0) run follows directly)
            else base += a.sub0[c];
            len = a.sublen[c];
        }
    }
    const unsigned sacc = smem0 + 128u + tid * 8u;
    if (zero_first) {
        for (int i = 0; i < len; i++) sts64(sacc + i * 1024u, 0.0);
    }
    __syncwarp();
    const unsigned fmask = zero_first ? 0u : 0xFFFFFFFFu;
    const unsigned sst = sstw + lane * 4u;
    if (nrows > 0) {
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
    }
    const unsigned srec = sstw + NW * pst + lane * (ND * 8u);
    for (int r = 0; r < nrows; r++) {
        const unsigned ra = sst + (unsigned)r * 128u;
        unsigned w[NW];
#pragma unroll
        for (int p = 0; p < (int)NW; p++) w[p] = lds32(ra + (unsigned)p * pst);
        const unsigned k = (w[NW - 1] >> 16) & 0xFu;
        const unsigned first = (w[NW - 1] >> 20) & fmask;
        if (k == 0xFu) continue;
        double v[ND];
#pragma unroll
        for (int j = 0; j < ND; j += 2) lds128(srec + (unsigned)r * (ND * 256u) + j * 8u, v[j], v[j + 1]);
        unsigned ad[ND];
        double o[ND];
#pragma unroll
        for (int j = 0; j < ND; j++) ad[j] = sacc + ((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) * 1024u;
#pragma unroll
        for (int j = 0; j < ND; j++) o[j] = lds64_unless(ad[j], (first >> j) & 1u);
#pragma unroll
        for (int j = 0; j < ND; j++) sts64(ad[j], o[j] + v[j] * a.factor);
        if (MIX) {
            const unsigned xs = lds32(sstw + STG * pst + lane * 4u + (unsigned)r * 128u);
            sts64(sacc + (unsigned)(len + (int)(xs >> 31)) * 1024u, __ldg(a.nzval + (xs & 0x7FFFFFFFu)));
        }
    }
    if (!valid) return;
    double* out = a.nzval + base;
    len += dlen;
    for (int i = 0; i < len; i++) {
        const double t = lds64(sacc + i * 1024u);
        out[i] = a.acc_A ? out[i] + t : t;
    }
}

// (div v_j, q_T) and its transposed copy: one contributing cell per entry -> plain adds through the slot maps, or plain stores
// (store bit 0: the block, bit 1: its transposed copy) when the block was claimed from the pending zero
// (ndl local dofs per cell, the flux function of face j is local dof j * fstep: RT0 4 / 1, BDM1 12 / 3 — the other BDM1 functions are divergence-free)
__global__ void k_rt0_div(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol, const int32_t* __restrict__ signs,
                          int64_t ncells, double f0, double f1, double f2, double f3, const int32_t* __restrict__ slots, const int32_t* __restrict__ slots_t,
                          double* __restrict__ nzval, int32_t* __restrict__ errflag, int ndl, int fstep, int store) {
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cell >= ncells) return;
    double x[4][3];
    load_nodes<3>(coords, cellnodes, cell, x);
    double A[3][3];
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int i = 0; i < 3; i++) A[d][i] = x[i + 1][d] - x[0][d];
    const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                       A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    const double vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * (1.0 / 6.0);
    const double f[4] = {f0, f1, f2, f3};
    unsigned err = 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const double v = (f[j] * ((double)__ldg(signs + cell * 4 + j) / det)) * vol;
        const int32_t s = __ldg(slots + cell * ndl + j * fstep);
        if (s < 0) err = 1u;
        else nzval[s] = (store & 1) ? v : nzval[s] + v;
        if (slots_t) {
            const int32_t st = __ldg(slots_t + cell * ndl + j * fstep);
            if (st < 0) err = 1u;
            else nzval[st] = (store & 2) ? v : nzval[st] + v;
        }
    }
    if (err) atomicAdd(errflag, 1);
}

// ---------------------------------------------------------------------------------------------------
// HDIVBDM1 on tetrahedra: every function of the basis is linear, so with its Piola image at the four vertices u_j(x_a)
//   M_jk = |T| / 20 ( (sum_a u_j(x_a)).(sum_b u_k(x_b)) + sum_a u_j(x_a).u_k(x_a) )       (int lambda_a lambda_b = |T| (1 + delta_ab) / 20)
// exactly what the order-2 rule integrates.  The reference functions at the reference vertices come from the evaluator's own
// table (extrapolated from its four quadrature points); subset (face orientations) and coefficients (face signs) per cell are
// the handlers of hdiv_bdm1.jl:215-249 (femb_device.cuh subset_of / coef_of).  One thread per (cell, local dof k): column k of
// the local matrix, written as 12 doubles to the slot of the (dof, cell) visit in the gather map's order.
// ---------------------------------------------------------------------------------------------------
struct Bdm1Tab {
    double V[16][4][3];
};

__global__ void __launch_bounds__(120) k_cell_bdm1rec(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol, SpaceDev s,
                                                      int64_t ncells, Bdm1Tab tab, const uint32_t* __restrict__ finv, double* __restrict__ out) {
    // 10 cells per CTA, one thread per (cell, local dof): phase 1 = the Piola image of its own function at the four vertices (and
    // their sum) into shared memory, phase 2 = column k of the local matrix from the twelve functions of the cell
    __shared__ double sV[16 * 12];
    __shared__ double su[10][12 * 16 + 2];  // [cell][dof][u(4 x 3), S(3), pad]; the cells of a warp start in different banks (their lanes read broadcast words)
    for (int i = threadIdx.x; i < 16 * 12; i += blockDim.x) sV[i] = (&tab.V[0][0][0])[i];
    __syncthreads();
    const int cl = threadIdx.x / 12, k = threadIdx.x - cl * 12;
    const int64_t cell = blockIdx.x * (int64_t)10 + cl;
    const bool active = cell < ncells;
    double vol = 0.0;
    if (active) {
        double x[4][3];
        load_nodes<3>(coords, cellnodes, cell, x);
        double A[3][3];
#pragma unroll
        for (int d = 0; d < 3; d++)
#pragma unroll
            for (int i = 0; i < 3; i++) A[d][i] = x[i + 1][d] - x[0][d];
        const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
        vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * (1.0 / 6.0);
        const double* v = sV + subset_of(s, cell, k) * 12;
        const double cf = coef_of(s, cell, k) / det;
        double S[3] = {0.0, 0.0, 0.0};
        double* o = su[cl] + k * 16;
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double u = cf * ((A[d][0] * v[a * 3] + A[d][1] * v[a * 3 + 1]) + A[d][2] * v[a * 3 + 2]);
                o[a * 3 + d] = u;
                S[d] += u;
            }
        o[12] = S[0]; o[13] = S[1]; o[14] = S[2];
    }
    __syncthreads();
    if (!active) return;
    double uk[15];
#pragma unroll
    for (int i = 0; i < 15; i++) uk[i] = su[cl][k * 16 + i];
    double* o = out + (size_t)__ldg(finv + cell * 12 + k) * 12;
    const double sc = vol * 0.05;
#pragma unroll
    for (int j0 = 0; j0 < 12; j0 += 4) {
        double m[4];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const double* uj = su[cl] + (j0 + jj) * 16;
            double acc = (uj[12] * uk[12] + uj[13] * uk[13]) + uj[14] * uk[14];
#pragma unroll
            for (int i = 0; i < 12; i++) acc += uj[i] * uk[i];
            m[jj] = acc * sc;
        }
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(o + j0), "d"(m[0]), "d"(m[1]), "d"(m[2]), "d"(m[3]) : "memory");
    }
}

bool is_bdm1_3d(const femb_ctx* ctx, const FembSpace& s) {
    return ctx->dim == 3 && s.family == FEMB_FAMILY_HDIV && s.handler == FEMB_HANDLER_BDM1_3D && s.nd == 12 && s.nd_all == 16 && s.signs && s.orient;
}

// reference functions at the reference vertices from the table at the quadrature points (linear functions: exact), false if
// the rule does not determine them
bool bdm1_vertex_values(const FembEval& E, Bdm1Tab& tab) {
    if (E.nqp != 4 || E.h_refvals.size() != (size_t)4 * 16 * 3) return false;
    double L[4][8];
    for (int q = 0; q < 4; q++) {
        const double xq = E.xref[q * 3], yq = E.xref[q * 3 + 1], zq = E.xref[q * 3 + 2];
        L[q][0] = 1.0 - xq - yq - zq; L[q][1] = xq; L[q][2] = yq; L[q][3] = zq;
        for (int a = 0; a < 4; a++) L[q][4 + a] = q == a ? 1.0 : 0.0;
    }
    for (int c = 0; c < 4; c++) {  // Gauss-Jordan with partial pivoting: [L | I] -> [I | L^-1]
        int piv = c;
        for (int r = c + 1; r < 4; r++)
            if (fabs(L[r][c]) > fabs(L[piv][c])) piv = r;
        if (fabs(L[piv][c]) < 1e-10) return false;
        for (int i = 0; i < 8; i++) std::swap(L[c][i], L[piv][i]);
        const double d = L[c][c];
        for (int i = 0; i < 8; i++) L[c][i] /= d;
        for (int r = 0; r < 4; r++)
            if (r != c) {
                const double f = L[r][c];
                for (int i = 0; i < 8; i++) L[r][i] -= f * L[c][i];
            }
    }
    // values(q) = sum_a lambda_a(x_q) V_a  =>  V_a = sum_q Linv[a][q] values(q)
    for (int f = 0; f < 16; f++)
        for (int a = 0; a < 4; a++)
            for (int l = 0; l < 3; l++) {
                double v = 0.0;
                for (int q = 0; q < 4; q++) v += L[a][4 + q] * E.h_refvals[((size_t)q * 16 + f) * 3 + l];
                tab.V[f][a][l] = v;
            }
    return true;
}

bool rt0_enabled() {
    static const bool on = [] { const char* e = getenv("FEMB_RT0_CLOSED"); return !(e && e[0] == '0'); }();
    return on;
}

bool is_rt0_3d(const femb_ctx* ctx, const FembSpace& s) {
    return ctx->dim == 3 && s.family == FEMB_FAMILY_HDIV && s.handler == FEMB_HANDLER_RT0_SIGNS && s.nd == 4 && s.nd_all == 4 && s.signs && s.nsign == 4;
}

}  // namespace

struct HdivMix {
    double fdiv;             // (sum_q w_q div phi_ref q_ref) * factor, the same for the four RT0 functions
    const int32_t* slots01;  // [cell][4] slots of block (0,1)
};

// shared tail of the RT0 / BDM1 mass paths: gather map of block (0,0), inverse slot map, cell pre-pass, column gather
// mix != nullptr (RT0, all three blocks still awaiting the lazy zero, H.xsrc built): the divergence blocks ride along
static int hdiv_mass_gather(femb_ctx* ctx, const FembEval& E, int nd, double factor, const Bdm1Tab* tab, const HdivMix* mix = nullptr) {
    const FembSpace& s = ctx->spaces[E.space];
    int st = femb_build_hmap_vec(ctx);
    if (st != FEMB_OK) return st;
    FembHMap& H = ctx->hm_vec;
    const int nw = (nd + 5) / 4;
    if (H.space != E.space || !H.complete || H.words != nw) return FEMB_ERR_UNSUPPORTED;
    for (auto& r : H.ranges)
        if (r.split != 1) return FEMB_ERR_UNSUPPORTED;
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
    const size_t stg = (size_t)nw + 2 * (size_t)nd + (mix ? 1 : 0);
    const int xrows = mix ? 2 : 0;  // accumulator rows of the (1,0) run
    auto smem_of = [stg, xrows](const FembHMap::Range& r) { return 128 + (size_t)(std::max(1, r.maxlen) + xrows) * 1024 + 4 * stg * (size_t)r.maxrows * 128; };
    size_t need = 0;
    for (auto& r : H.cranges) need = std::max(need, smem_of(r));
    if (need > 200 * 1024) return FEMB_ERR_UNSUPPORTED;
    if (!ctx->rt0rec) {
        FEMB_TRY(femb_alloc(ctx, &ctx->rt0rec, (size_t)std::max<size_t>(H.stride, 1) * nd));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->rt0rec, 0, sizeof(double) * std::max<size_t>(H.stride, 1) * nd, ctx->stream));
    }
    if (!H.vinv) {
        FEMB_TRY(femb_alloc(ctx, &H.vinv, (size_t)std::max<int64_t>(ctx->ncells, 1) * nd));
        if (H.stride) {
            k_hdiv_finv<<<(unsigned)((H.stride + 255) / 256), 256, 0, ctx->stream>>>(H.map, H.stride, nd, nw, H.vinv);
            KERNEL_CHECK(ctx);
        }
    }
    if (ctx->ncells == 0) return FEMB_OK;
    // right after femb_matrix_zero the gather claims its block: every run is overwritten (not read, added to and written back) and the
    // zeroing of the block is skipped altogether; the other blocks keep waiting for their own kernels or femb_flush_zero
    const bool fresh = femb_claim_block(ctx, 0, 0);
    if (mix && !(fresh && femb_claim_block(ctx, 0, 1) && femb_claim_block(ctx, 1, 0))) return femb_fail(ctx, FEMB_ERR_ARG, "mixed RT0 pass on a matrix that is not freshly zeroed");
    if (!fresh) FEMB_TRY(femb_flush_zero(ctx, true, false));
    if (nd == 4) {
        k_cell_rt0rec<<<(unsigned)((ctx->ncells + 127) / 128), 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, s.signs, ctx->ncells, H.vinv, ctx->rt0rec,
                                                                                     mix ? mix->slots01 : nullptr, mix ? mix->fdiv : 0.0, ctx->nzval);
    } else {
        k_cell_bdm1rec<<<(unsigned)((ctx->ncells + 9) / 10), 120, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, space_dev(s), ctx->ncells, *tab, H.vinv,
                                                                                         ctx->rt0rec);
    }
    KERNEL_CHECK(ctx);
    RGArgs a;
    a.rec = ctx->rt0rec; a.slice_ptr = H.slice_ptr; a.slice_zero = H.slice_zero; a.map = H.map; a.stride = H.stride;
    a.colptr = ctx->colptr + H.col_off; a.sub0 = H.sub0; a.sublen = H.sublen; a.ndofs = s.ndofs; a.nzval = ctx->nzval;
    a.factor = factor;
    a.acc_A = fresh ? 0 : 1;
    a.xsrc = mix ? H.xsrc : nullptr;
    const void* kern = nd == 4 ? (mix ? (const void*)k_hdiv_mass_gather<4, true> : (const void*)k_hdiv_mass_gather<4>) : (const void*)k_hdiv_mass_gather<12>;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (auto& r : H.cranges) {
        if (r.s1 <= r.s0) continue;
        a.slice_begin = r.s0;
        a.slice_end = r.s1;
        a.maxlen = std::max(1, r.maxlen) + xrows;
        a.maxrows = r.maxrows;
        void* params[1] = {(void*)&a};
        CUDA_TRY(ctx, cudaLaunchKernel(kern, dim3((unsigned)((r.s1 - r.s0 + 3) / 4)), dim3(128), params, smem_of(r), ctx->stream));
        KERNEL_CHECK(ctx);
    }
    return FEMB_OK;
}

// Measured on B200 (Kuhn m = 64, 1.57M tets): pre-pass 1.9 ms + gather 1.6 ms vs 2.4 ms for the mass block through the generic
// kernel with its 144 reductions per cell — the 96-byte columns scattered to their visit slots and the 3 KB per cell the gather
// moves cost more than the reductions.  Kept behind FEMB_BDM1_MASS_GATHER=1 (bit-reproducible, parity-tested), off by default;
// the divergence blocks of BDM1 do use the closed form (launch_rt0_div: 1.23 -> 0.13 ms).
int launch_bdm1_mass_gather(femb_ctx* ctx, int er, int brow, int bcol, double factor) {
    static const bool on = [] { const char* e = getenv("FEMB_BDM1_MASS_GATHER"); return e && e[0] == '1'; }();
    if (!on || !rt0_enabled()) return FEMB_ERR_UNSUPPORTED;
    const FembEval& E = ctx->evals[er];
    const FembSpace& s = ctx->spaces[E.space];
    if (!is_bdm1_3d(ctx, s) || brow != 0 || bcol != 0 || ctx->nrs < 1 || ctx->row_spaces[0] != E.space || ctx->col_spaces[0] != E.space) return FEMB_ERR_UNSUPPORTED;
    Bdm1Tab tab;
    if (!bdm1_vertex_values(E, tab)) return FEMB_ERR_UNSUPPORTED;
    {
        // the vertex form must reproduce  sum_q w_q (A phi_j(q)).(A phi_k(q))  of the evaluator's table for all 16 x 16 reference functions
        const double A[3][3] = {{1.0, 0.2, 0.1}, {0.2, 0.9, 0.1}, {-0.1, 0.1, 1.2}};
        for (int j = 0; j < 16; j++)
            for (int k = j; k < 16; k++) {
                double ref = 0.0, closed = 0.0, Sj[3] = {0, 0, 0}, Sk[3] = {0, 0, 0};
                for (int q = 0; q < 4; q++) {
                    double pj[3] = {0, 0, 0}, pk[3] = {0, 0, 0};
                    for (int d = 0; d < 3; d++)
                        for (int l = 0; l < 3; l++) {
                            pj[d] += A[d][l] * E.h_refvals[((size_t)q * 16 + j) * 3 + l];
                            pk[d] += A[d][l] * E.h_refvals[((size_t)q * 16 + k) * 3 + l];
                        }
                    ref += E.w[q] * (pj[0] * pk[0] + pj[1] * pk[1] + pj[2] * pk[2]);
                }
                for (int a = 0; a < 4; a++) {
                    double uj[3] = {0, 0, 0}, uk[3] = {0, 0, 0};
                    for (int d = 0; d < 3; d++)
                        for (int l = 0; l < 3; l++) {
                            uj[d] += A[d][l] * tab.V[j][a][l];
                            uk[d] += A[d][l] * tab.V[k][a][l];
                        }
                    for (int d = 0; d < 3; d++) {
                        Sj[d] += uj[d];
                        Sk[d] += uk[d];
                        closed += uj[d] * uk[d];
                    }
                }
                closed = (closed + Sj[0] * Sk[0] + Sj[1] * Sk[1] + Sj[2] * Sk[2]) * 0.05;
                if (fabs(ref - closed) > 1e-11 * (1.0 + fabs(ref))) return FEMB_ERR_UNSUPPORTED;
            }
    }
    return hdiv_mass_gather(ctx, E, 12, factor, &tab);
}

static int rt0_mass_checked(femb_ctx* ctx, int er, double factor, const HdivMix* mix);

int launch_rt0_mass_gather(femb_ctx* ctx, int er, int brow, int bcol, double factor) {
    if (brow != 0 || bcol != 0) return FEMB_ERR_UNSUPPORTED;
    return rt0_mass_checked(ctx, er, factor, nullptr);
}

static int rt0_mass_checked(femb_ctx* ctx, int er, double factor, const HdivMix* mix) {
    if (!rt0_enabled()) return FEMB_ERR_UNSUPPORTED;
    const FembEval& E = ctx->evals[er];
    const FembSpace& s = ctx->spaces[E.space];
    if (!is_rt0_3d(ctx, s) || ctx->nrs < 1 || ctx->row_spaces[0] != E.space || ctx->col_spaces[0] != E.space) return FEMB_ERR_UNSUPPORTED;
    if (E.h_refvals.size() != (size_t)E.nqp * 4 * 3) return FEMB_ERR_UNSUPPORTED;
    {
        // the closed form must reproduce  |T| sum_q w_q (A phi_j(q) / det).(A phi_k(q) / det)  of the evaluator's table on a test cell
        const double X[4][3] = {{0.1, 0.0, 0.05}, {1.1, 0.2, 0.0}, {0.3, 0.9, 0.1}, {0.2, 0.1, 1.2}};
        double A[3][3], c[3], S = 0.0;
        for (int d = 0; d < 3; d++) {
            for (int i = 0; i < 3; i++) A[d][i] = X[i + 1][d] - X[0][d];
            c[d] = 0.25 * (X[0][d] + X[1][d] + X[2][d] + X[3][d]);
        }
        const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
        const double vol = fabs(det) / 6.0;
        for (int i = 0; i < 4; i++)
            for (int d = 0; d < 3; d++) S += (X[i][d] - c[d]) * (X[i][d] - c[d]);
        S *= 0.05;
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) {
                double ref = 0.0, dot = 0.0;
                for (int q = 0; q < E.nqp; q++) {
                    double pj[3] = {0, 0, 0}, pk[3] = {0, 0, 0};
                    for (int d = 0; d < 3; d++)
                        for (int l = 0; l < 3; l++) {
                            pj[d] += A[d][l] * E.h_refvals[((size_t)q * 4 + j) * 3 + l];
                            pk[d] += A[d][l] * E.h_refvals[((size_t)q * 4 + k) * 3 + l];
                        }
                    ref += E.w[q] * (pj[0] * pk[0] + pj[1] * pk[1] + pj[2] * pk[2]);
                }
                ref *= vol / (det * det);
                for (int d = 0; d < 3; d++) dot += (c[d] - X[rt0_opp(j)][d]) * (c[d] - X[rt0_opp(k)][d]);
                const double closed = 4.0 * vol / (det * det) * (dot + S);
                if (fabs(ref - closed) > 1e-12 * fabs(closed) + 1e-14) return FEMB_ERR_UNSUPPORTED;
            }
    }
    return hdiv_mass_gather(ctx, E, 4, factor, nullptr, mix);
}

int launch_rt0_div(femb_ctx* ctx, int er, int ec, int brow, int bcol, double factor, int flags) {
    if (!rt0_enabled()) return FEMB_ERR_UNSUPPORTED;
    const FembEval& ER = ctx->evals[er];
    const FembEval& EC = ctx->evals[ec];
    const FembSpace& rs = ctx->spaces[ER.space];
    const FembSpace& cs = ctx->spaces[EC.space];
    const bool bdm = is_bdm1_3d(ctx, rs);
    if ((!is_rt0_3d(ctx, rs) && !bdm) || cs.family != FEMB_FAMILY_L2 || cs.nd != 1 || cs.ncomp != 1 || ER.nqp != EC.nqp) return FEMB_ERR_UNSUPPORTED;
    const int nall = bdm ? 16 : 4, ndl = bdm ? 12 : 4, fstep = bdm ? 3 : 1;
    if (brow < 0 || brow >= ctx->nrs || bcol < 0 || bcol >= ctx->ncs || ctx->row_spaces[brow] != ER.space || ctx->col_spaces[bcol] != EC.space) return FEMB_ERR_UNSUPPORTED;
    if (ER.h_refderivs.size() != (size_t)ER.nqp * nall * 9 || EC.h_refvals.size() != (size_t)EC.nqp) return FEMB_ERR_UNSUPPORTED;
    const FembBlock* blk = find_block(ctx, brow, bcol);
    const FembBlock* blk_t = nullptr;
    int tr = -1, tc = -1;
    if (flags & FEMB_BIL_TRANSPOSE_COPY) {
        for (int i = 0; i < ctx->nrs; i++) if (ctx->row_spaces[i] == EC.space) tr = i;
        for (int i = 0; i < ctx->ncs; i++) if (ctx->col_spaces[i] == ER.space) tc = i;
        blk_t = (tr >= 0 && tc >= 0) ? find_block(ctx, tr, tc) : nullptr;
        if (!blk_t) return FEMB_ERR_UNSUPPORTED;
    }
    if (!blk) return FEMB_ERR_UNSUPPORTED;
    FEMB_TRY(femb_build_slotmaps(ctx, brow, bcol));
    if (blk_t) FEMB_TRY(femb_build_slotmaps(ctx, tr, tc));
    blk = find_block(ctx, brow, bcol);
    blk_t = blk_t ? find_block(ctx, tr, tc) : nullptr;
    // sum_q w_q trace(D phi_j)(q) q_T(q): constant per reference function (feevaluator_hdiv.jl:66-83, the L2P0 value is 1)
    double f[4] = {0.0, 0.0, 0.0, 0.0};
    for (int fn = 0; fn < nall; fn++) {
        double t = 0.0;
        for (int q = 0; q < ER.nqp; q++) {
            double tr3 = 0.0;
            for (int d = 0; d < 3; d++) tr3 += ER.h_refderivs[((size_t)q * nall + fn) * 9 + d * 3 + d];
            t += ER.w[q] * tr3 * EC.h_refvals[q];
        }
        if (!bdm) f[fn] = t * factor;
        else if (fn % 4 == 0) f[fn / 4] = t * factor;  // the RT0 function of face fn / 4 (subset index 4 j, hdiv_bdm1.jl:240)
        else if (fabs(t) > 1e-13) return FEMB_ERR_UNSUPPORTED;  // the moment functions must be divergence-free in the mean
    }
    if (ctx->ncells == 0) return femb_flush_zero(ctx, true, false);
    // a block whose pattern holds nothing but the (face dof, cell) couplings — 4 per cell (RT0), 12 per cell with the exact zeros of the
    // divergence-free BDM1 functions — is overwritten entry by entry: claim it from the pending zero (RT0) instead of zeroing and adding
    int store = 0;
    if (ctx->zero_pending_A && !bdm) {
        int64_t n = 0, nt = 0;
        FEMB_TRY(femb_block_entries(ctx, brow, bcol, &n));
        if (blk_t) FEMB_TRY(femb_block_entries(ctx, tr, tc, &nt));
        if (n == ctx->ncells * 4 && femb_claim_block(ctx, brow, bcol)) store |= 1;
        if (blk_t && nt == ctx->ncells * 4 && femb_claim_block(ctx, tr, tc)) store |= 2;
    }
    if (store != (blk_t ? 3 : 1)) {
        // something is added to: settle the rest of the pending zero (the claimed blocks stay out of it)
        FEMB_TRY(femb_flush_zero(ctx, true, false));
    }
    k_rt0_div<<<(unsigned)((ctx->ncells + 127) / 128), 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, rs.signs, ctx->ncells, f[0], f[1], f[2], f[3], blk->slots,
                                                                             blk_t ? blk_t->slots : nullptr, ctx->nzval, ctx->errflag, ndl, fstep, store);
    KERNEL_CHECK(ctx);
    return FEMB_OK;
}

// Mass block and both divergence blocks of a [v; q] layout (v: HDIVRT0 on tetrahedra, q: L2P0) in two kernels on a freshly zeroed matrix: the cell
// pre-pass also stores the cell's own column of block (0,1) (one sector), the face-column gather also copies those values into the (1,0) run
// that follows its mass run, so that every column of the matrix is written exactly once and in one piece — no zeroing, no partial sectors.
// FEMB_ERR_UNSUPPORTED (the caller runs the two separate assemblies) whenever something is not exactly that.
int launch_rt0_mixed(femb_ctx* ctx, int eid, int ediv, int eq, double fmass, double fdivf) {
    static const bool on = [] { const char* e = getenv("FEMB_RT0_MIXED"); return !(e && e[0] == '0'); }();
    if (!on || !rt0_enabled()) return FEMB_ERR_UNSUPPORTED;
    const FembEval& EI = ctx->evals[eid];
    const FembEval& ED = ctx->evals[ediv];
    const FembEval& EQ = ctx->evals[eq];
    const FembSpace& vs = ctx->spaces[EI.space];
    const FembSpace& qs = ctx->spaces[EQ.space];
    if (!is_rt0_3d(ctx, vs) || ED.space != EI.space || EI.op != FEMB_OP_IDENTITY || ED.op != FEMB_OP_DIVERGENCE || EQ.op != FEMB_OP_IDENTITY) return FEMB_ERR_UNSUPPORTED;
    if (qs.family != FEMB_FAMILY_L2 || qs.nd != 1 || qs.ncomp != 1 || ED.nqp != EQ.nqp) return FEMB_ERR_UNSUPPORTED;
    if (ctx->nrs != 2 || ctx->ncs != 2 || ctx->row_spaces[0] != EI.space || ctx->col_spaces[0] != EI.space || ctx->row_spaces[1] != EQ.space || ctx->col_spaces[1] != EQ.space)
        return FEMB_ERR_UNSUPPORTED;
    if (!find_block(ctx, 0, 0) || !find_block(ctx, 0, 1) || !find_block(ctx, 1, 0) || ctx->ncells == 0) return FEMB_ERR_UNSUPPORTED;
    const uint64_t bits = (1ull << 0) | (1ull << 1) | (1ull << FEMB_MAX_SPACES);
    if (!ctx->zero_pending_A || (ctx->zero_mask & bits) != bits) return FEMB_ERR_UNSUPPORTED;  // something was assembled since femb_matrix_zero: add, do not store
    if (ED.h_refderivs.size() != (size_t)ED.nqp * 4 * 9 || EQ.h_refvals.size() != (size_t)EQ.nqp) return FEMB_ERR_UNSUPPORTED;
    double f[4];
    for (int fn = 0; fn < 4; fn++) {
        double t = 0.0;
        for (int q = 0; q < ED.nqp; q++) {
            double tr3 = 0.0;
            for (int d = 0; d < 3; d++) tr3 += ED.h_refderivs[((size_t)q * 4 + fn) * 9 + d * 3 + d];
            t += ED.w[q] * tr3 * EQ.h_refvals[q];
        }
        f[fn] = t * fdivf;
    }
    if (f[1] != f[0] || f[2] != f[0] || f[3] != f[0]) return FEMB_ERR_UNSUPPORTED;
    int64_t n01 = 0, n10 = 0;
    FEMB_TRY(femb_block_entries(ctx, 0, 1, &n01));
    FEMB_TRY(femb_block_entries(ctx, 1, 0, &n10));
    if (n01 != ctx->ncells * 4 || n10 != ctx->ncells * 4) return FEMB_ERR_UNSUPPORTED;  // entries beyond the (face dof, cell) couplings would stay unwritten
    int st = femb_build_hmap_vec(ctx);
    if (st != FEMB_OK) return st;
    FembHMap& H = ctx->hm_vec;
    if (H.space != EI.space || !H.complete || H.words != 2 || !H.sub0 || !H.sublen) return FEMB_ERR_UNSUPPORTED;
    FEMB_TRY(femb_build_slotmaps(ctx, 0, 1));
    FEMB_TRY(femb_build_slotmaps(ctx, 1, 0));
    const FembBlock* b01 = find_block(ctx, 0, 1);
    const FembBlock* b10 = find_block(ctx, 1, 0);
    if (H.xsrc_state == 0) {
        int32_t* bad = nullptr;
        int32_t hbad = 0;
        FEMB_TRY(femb_alloc(ctx, &H.xsrc, std::max<size_t>(H.stride, 1)));
        CUDA_TRY(ctx, cudaMalloc(&bad, sizeof(int32_t)));
        cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(int32_t), ctx->stream);
        if (e == cudaSuccess && H.stride) {
            k_hdiv_xsrc<<<(unsigned)((H.stride + 255) / 256), 256, 0, ctx->stream>>>(H.map, H.stride, 2, vs.celldofs, ctx->colptr + H.col_off, H.sub0, H.sublen, b01->slots, b10->slots,
                                                                                    H.xsrc, bad);
            e = cudaGetLastError();
            ctx->launches++;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hbad, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(bad);
        if (e != cudaSuccess) return femb_fail(ctx, FEMB_ERR_CUDA, "k_hdiv_xsrc: %s", cudaGetErrorString(e));
        H.xsrc_state = hbad == 0 ? 1 : -1;
    }
    if (H.xsrc_state != 1) return FEMB_ERR_UNSUPPORTED;
    // the closed-form mass matrix must pass the same table check as the stand-alone gather (launch_rt0_mass_gather); run it through that door
    HdivMix mix{f[0], b01->slots};
    return rt0_mass_checked(ctx, eid, fmass, &mix);
}
