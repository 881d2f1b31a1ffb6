// Scalar H1P1 Poisson (Example200 with order 1): the fused stages-1-5 owner-computes kernels (triangles: the headline
// kernel; tetrahedra), their launchers (slice ranges, overlapped NCCL exchange) and the streamed host-to-host entry.
#include "femb_kernels.cuh"

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
    const double* gvol;  // CellVolumes in visit order [row][lane] (femb_ctx::gvol), or NULL
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

// one visit of the 2-D P1 gather: column k of the cell's local stiffness matrix (+ rhs) into the accumulators
template <bool TRACK, int RHS, bool HAS_VOL, int NQP>
__device__ __forceinline__ void p1_visit(const unsigned w, const unsigned cell, const double2 self, const double2 pa, const double2 pb, const double volr,
                                         const P1Tab2& tab, const RhsDev& rhs, const double drop_tol, const double mu_wsum, const unsigned sacc,
                                         double& diag, unsigned& dpos, unsigned& err, unsigned long long& mask, double& bsum) {
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
        const unsigned ad = sacc + pA * 8u;
        sts64(ad, lds64(ad) + va);
        if (TRACK) mask |= 1ull << (pA & 63u);
    }
    if (kb && pB != 0xFFu) {
        const unsigned ad = sacc + pB * 8u;
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
#ifndef FEMB_P1_OPAQUE
#define FEMB_P1_OPAQUE 1
#endif
#ifndef FEMB_P1_UNROLL
#define FEMB_P1_UNROLL 2
#endif
#ifndef FEMB_P1_STAGE16
#define FEMB_P1_STAGE16 1
#endif
#ifndef FEMB_P1_MINB
#define FEMB_P1_MINB 8  // resident CTAs per SM the 2-D gather kernel is compiled for (register cap 65536 / (128 * MINB))
#endif
template <bool TRACK, int RHS, bool HAS_VOL, int NQP, int STAGE, int MINB = FEMB_P1_MINB>  // NQP: compile-time number of quadrature points, 0 = tab.nqp; STAGE: 0 = map entries by loads, 1 = staged by bulk copies, 2 = + CellVolumes in visit order
__global__ void __launch_bounds__(128, MINB) k_p1_poisson_gather2d(const P1Args a, const RhsDev rhs, const P1Tab2 tab) {
    // VOLP: xgrid[CellVolumes] comes as one more staged plane, in visit order (built once per mesh from the registered volumes):
    // no cell-index plane and no scattered 8-byte load per visit — the given-volume kernel then runs like the |det| / 2 one
    constexpr bool VOLP = HAS_VOL && STAGE == 2;
    constexpr bool NEED_CELL = (HAS_VOL && !VOLP) || RHS == FEMB_RHS_QPVALUES;
    constexpr int NPL = NEED_CELL ? 4 : 3;  // gather-map planes this variant reads
    // dynamic shared memory: [4 mbarriers | pad to 128 B][4 warps x flat accumulators (maxlen * 32 + 2 doubles)][(STAGE) u32 [4 warps][4 planes][maxrows][32]]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int64_t slice = a.slice_begin + (int64_t)blockIdx.x * 4 + warp;
    if (slice >= a.slice_end) return;  // whole warp out of range
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned accw = ((unsigned)a.maxlen * 32u + 2u) * 8u;  // bytes of one warp's accumulator block
    const unsigned wacc = smem0 + 128u + warp * accw;
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
    const uint32_t* __restrict__ ga = a.gmap + ((size_t)row0 << 5) + lane;
    const size_t gs = a.gstride;
    // STAGE: the slice's gather-map entries (all visits, NPL planes of nrows x 128 bytes, each contiguous in global memory)
    // come in as NPL bulk copies (cp.async.bulk, SASS UBLKCP) issued by ONE lane and signalled on a per-warp mbarrier: no
    // per-lane copy instructions, no L1 wavefronts, one memory latency for all visits — and they fly while the warp
    // reads colptr / its own coordinates and clears the accumulators.
    const unsigned mbar = smem0 + warp * 8u;
    const unsigned sstw = smem0 + 128u + 4u * accw + warp * (4u * (unsigned)a.maxrows * 128u);  // this warp's staging block
    const unsigned pst = (unsigned)a.maxrows * 128u;                                              // plane stride in bytes
    const unsigned svolw = smem0 + 128u + 4u * accw + 16u * pst + warp * 2u * pst;                // (VOLP) this warp's volume plane: maxrows x 32 doubles
    if (STAGE && nrows > 0) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const unsigned bytes = (unsigned)nrows * 128u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes * (NPL + (VOLP ? 2 : 0))) : "memory");
            if (VOLP)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(svolw), "l"(a.gvol + ((size_t)row0 << 5)),
                             "r"(2u * bytes), "r"(mbar)
                             : "memory");
#pragma unroll
            for (int p = 0; p < NPL; p++)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sstw + (unsigned)p * pst),
                             "l"(ga + (size_t)p * gs), "r"(bytes), "r"(mbar)
                             : "memory");
        }
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
    // The 32 columns of the slice are one contiguous range [base0, base0 + total) of nzval: the accumulators mirror it
    // (element i of the range at wacc + (i + par) * 8, par = parity of base0 so that 16-byte aligned global runs are 16-byte
    // aligned in shared memory).  Lanes with the same position in equally long columns hit different banks (odd column
    // lengths; two lanes per 8-byte bank pair = the minimum of a 64-bit access), and the finished range leaves with ONE bulk
    // store instead of `len` strided 8-byte stores per lane (one L1 wavefront per lane and entry before).
    const int nvalid = (int)min((int64_t)32, a.ndofs - slice * 32);
    const int64_t base0 = __shfl_sync(0xFFFFFFFFu, base, 0);
    const int64_t endl = __shfl_sync(0xFFFFFFFFu, base + len, nvalid - 1);
    const int total = (int)(endl - base0);
    const unsigned par = (unsigned)(base0 & 1);
    if (!valid) base = base0;
    unsigned sacc = wacc + ((unsigned)(base - base0) + par) * 8u;
#if FEMB_P1_OPAQUE
    asm volatile("" : "+r"(sacc));  // keep the accumulator address in its register: re-deriving it costs six instructions per use
#endif
    for (int i = (int)lane; i < total + 2; i += 32) sts64(wacc + (unsigned)i * 8u, 0.0);
    __syncwarp();
    unsigned long long mask = 0ull;
    double bsum = 0.0, diag = 0.0;
    unsigned err = 0u, dpos = 0xFFu;
    const double drop_tol = a.drop_tol, mu_wsum = tab.mu_wsum;

    if (STAGE) {
        const unsigned sst = sstw + lane * 4u;
        if (nrows > 0) {  // wait for the bulk copies (phase 0 of the barrier)
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
        }
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
            if (HAS_VOL && !VOLP) vol1 = __ldg(a.cellvol + lds32(sst + 3u * pst));
        }
        const unsigned svol = svolw + lane * 8u;
        constexpr int UNR = FEMB_P1_UNROLL;
#pragma unroll UNR
        for (int r = 0; r < nrows; r++) {
            const unsigned d = sst + (unsigned)r * 128u;
            const unsigned w = lds32(d + 2u * pst);
            const unsigned cell = NEED_CELL ? lds32(d + 3u * pst) : 0u;
            const double2 pa = pa1, pb = pb1;
            const double volr = VOLP ? lds64(svol + (unsigned)r * 256u) : vol1;
            if (r + 1 < nrows) {  // coordinates / volume of the next visit
                pa1 = __ldg(xy + lds32(d + 128u));
                pb1 = __ldg(xy + lds32(d + 128u + pst));
                if (HAS_VOL && !VOLP) vol1 = __ldg(a.cellvol + lds32(d + 128u + 3u * pst));
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
    if (valid && err) atomicAdd(a.errflag, 1);
    if (valid && dpos != 0xFFu) {
        sts64(sacc + dpos * 8u, diag);  // the diagonal row received no shared-memory update
        if (TRACK) mask |= 1ull << (dpos & 63u);
    }
    double* out0 = a.nzval + base0;
    if (a.acc_A) {
        __syncwarp();
        for (int i = (int)lane; i < total; i += 32) out0[i] += lds64(wacc + ((unsigned)i + par) * 8u);  // coalesced
    } else if (total > 0) {
        // one bulk store of the 16-byte aligned part of the range (cp.async.bulk.global.shared::cta, SASS UBLKCP), scalar head / tail
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
        __syncwarp();
        if (lane == 0) {
            const int head = (int)par, body = (total - head) & ~1;
            if (head) out0[0] = lds64(wacc + 8u);
            if (head + body < total) out0[total - 1] = lds64(wacc + ((unsigned)(total - 1) + par) * 8u);
            if (body > 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out0 + head), "r"(wacc + 2u * par * 8u), "r"((unsigned)body * 8u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must stay alive until it has been read
            }
        }
    }
    if (!valid) return;
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

// CellVolumes in visit order: out[row][lane] = cellvol[cell of that gather-map entry] (padding entries: cell 0)
__global__ void k_gvol(const uint32_t* __restrict__ cells, const double* __restrict__ cellvol, size_t n, int64_t ncells, double* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cells[i];
    out[i] = c < (uint64_t)ncells ? cellvol[c] : 0.0;
}

static int prepare_p1_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol, P1Launch& L, bool use_volplane) {
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
    a.coords = ctx->coords; a.cellvol = ctx->cellvol; a.gvol = nullptr; a.gslice_ptr = ctx->gslice_ptr; a.gmap = ctx->gmap; a.gstride = ctx->gmap_stride;
    a.colptr = ctx->colptr; a.ndofs = s.ndofs; a.nzval = ctx->nzval; a.touched = ctx->touched; a.bvec = ctx->bvec; a.errflag = ctx->errflag;
    a.mu = mu; a.drop_tol = drop_tol;
    FEMB_TRY(femb_settle_partial_zero(ctx));
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
        L.smem = 128 + 4 * ((size_t)a.maxlen * 32 + 2) * sizeof(double);  // mbarriers + per-warp flat accumulators
        // staged variant: + u32 [4 warps][4 planes][maxrows][32] for the slice's gather-map entries
        static const int stage_env = getenv("FEMB_P1_STAGE") ? atoi(getenv("FEMB_P1_STAGE")) : 1;
        a.maxrows = (stage_env && ctx->gmap_maxrows > 0 && ctx->gmap_maxrows <= 12) ? ctx->gmap_maxrows : 0;
        L.smem += (size_t)a.maxrows * 4 * 4 * 32 * sizeof(uint32_t);
        // CellVolumes in visit order (one more staged plane): built once per registered mesh, for the device-resident assembly
        static const int volp_env = getenv("FEMB_P1_VOLPLANE") ? atoi(getenv("FEMB_P1_VOLPLANE")) : 1;
        if (volp_env && use_volplane && a.maxrows > 0 && ctx->cellvol && !ctx->touched && rhs.kind != FEMB_RHS_QPVALUES) {
            if (!ctx->gvol_valid) {
                if (!ctx->gvol) FEMB_TRY(femb_alloc(ctx, &ctx->gvol, ctx->gmap_stride));
                if (ctx->gmap_stride) {
                    k_gvol<<<(unsigned)((ctx->gmap_stride + 255) / 256), 256, 0, ctx->stream>>>(ctx->gmap + 3 * ctx->gmap_stride, ctx->cellvol, ctx->gmap_stride, ctx->ncells, ctx->gvol);
                    KERNEL_CHECK(ctx);
                }
                ctx->gvol_valid = true;
            }
            a.gvol = ctx->gvol;
            L.smem += (size_t)a.maxrows * 4 * 32 * sizeof(double);
        }
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
#define FEMB_P1_LAUNCH2Q(T, R, V, Q)                                           \
    do {                                                                       \
        if (a.maxrows > 0 && V && !T && a.gvol) FEMB_P1_LAUNCH2S(T, R, V, Q, (V && !T) ? 2 : 1); \
        else if (a.maxrows > 0) FEMB_P1_LAUNCH2S(T, R, V, Q, 1);               \
        else FEMB_P1_LAUNCH2S(T, R, V, Q, 0);                                  \
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

int launch_p1_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol) {
    P1Launch L;
    FEMB_TRY(prepare_p1_gather(ctx, eg, ei, mu, rhs, drop_tol, L, true));
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
            ctx->exchange_done = true;
            return FEMB_OK;
        }
        FEMB_TRY(launch_p1_slices(ctx, L, 0, L.nslices, ctx->stream));  // interface too large to be worth splitting
        FEMB_TRY(femb_exchange_on(ctx, ctx->stream));
        ctx->exchange_done = true;
        return FEMB_OK;
    }
    return launch_p1_slices(ctx, L, 0, L.nslices, ctx->stream);
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
    if (!(ctx->gmap && ctx->gmap_space == EG.space) && ctx->hm.map && ctx->hm.space == EG.space && !ctx->touched && ctx->dim == 3 &&
        (rhs_kind == FEMB_RHS_NONE || rhs_kind == FEMB_RHS_AFFINE)) {
        // scalar P2 on tetrahedra (closed-form gather, femb_p2.cu): upload the geometry, assemble range by range, download the
        // finished nzval / b ranges behind the kernels
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        RhsDev rhs2;
        double* dev_f2 = nullptr;
        FEMB_TRY(make_rhs(ctx, rhs_kind, params, 1, EG.nqp, rhs2, &dev_f2));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (!ctx->stream_h2d) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_h2d, cudaStreamNonBlocking));
        if (!ctx->stream_d2h) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_d2h, cudaStreamNonBlocking));
        ctx->gvol_valid = false;
        ctx->zero_pending_A = true;
        ctx->zero_mask = ~0ull;
        ctx->zero_pending_b = true;
        FEMB_TRY(femb_timer_begin(ctx));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->coords, coords, sizeof(double) * ctx->nnodes * 3, cudaMemcpyHostToDevice, ctx->stream_h2d));
        if (cellvol) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->cellvol, cellvol, sizeof(double) * ctx->ncells, cudaMemcpyHostToDevice, ctx->stream_h2d));
        if (ctx->chunk_ev.empty()) {
            cudaEvent_t ev;
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->chunk_ev.push_back(ev);
        }
        CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[0], ctx->stream_h2d));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[0], 0));
        P2HostOut ho = {nzval_out, b_out, std::max(1, (int)nchunks)};
        int st2 = launch_p2_gather(ctx, eg, ei, mu, rhs2, drop_tol, &ho);
        if (st2 == FEMB_OK && cellnodes) st2 = femb_cellnodes_compare_async(ctx, cellnodes, ctx->stream_h2d);
        if (st2 == FEMB_OK) st2 = femb_timer_end(ctx);
        cudaError_t e1 = cudaStreamSynchronize(ctx->stream_h2d), e2 = cudaStreamSynchronize(ctx->stream_d2h);
        cudaFree(dev_f2);
        if (st2 == FEMB_ERR_UNSUPPORTED) return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "the streamed path covers scalar P1 and (closed-form) P2 on tetrahedra; use femb_mesh_set + femb_assemble_poisson + femb_matrix_get");
        if (st2 != FEMB_OK) return st2;
        CUDA_TRY(ctx, e1);
        CUDA_TRY(ctx, e2);
        if (cellnodes) {
            bool differ = false;
            FEMB_TRY(femb_cellnodes_differ(ctx, &differ));
            if (differ)
                return femb_fail(ctx, FEMB_ERR_ARG, "femb_assemble_poisson_host: cellnodes differs from the connectivity the pattern was built from; call femb_mesh_set and rebuild the pattern");
        }
        return femb_sync(ctx);
    }
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
    ctx->zero_mask = ~0ull;
    ctx->zero_pending_b = true;
    P1Launch L;
    ctx->gvol_valid = false;  // this call uploads new volumes chunk by chunk: the visit-order copy is not used and goes stale
    FEMB_TRY(prepare_p1_gather(ctx, eg, ei, mu, rhs, drop_tol, L, false));
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
    // cellnodes != NULL: the caller is not sure the connectivity is the registered one.  The numeric phase never reads it
    // (the gather map of the symbolic phase is the topology), so it is uploaded behind the geometry and compared on the
    // device: a difference means the result above belongs to the OLD topology -> FEMB_ERR_ARG, the caller must
    // femb_mesh_set + rebuild the pattern.  NULL = connectivity declared unchanged (geometry-only step).
    int st = FEMB_OK;
    if (cellnodes) st = femb_cellnodes_compare_async(ctx, cellnodes, ctx->stream_h2d);
    if (st == FEMB_OK) st = femb_timer_end(ctx);
    cudaError_t e1 = cudaStreamSynchronize(ctx->stream_h2d), e2 = cudaStreamSynchronize(ctx->stream_d2h);
    cudaFree(dev_f);
    if (st != FEMB_OK) return st;
    CUDA_TRY(ctx, e1);
    CUDA_TRY(ctx, e2);
    if (cellnodes) {
        bool differ = false;
        FEMB_TRY(femb_cellnodes_differ(ctx, &differ));
        if (differ)
            return femb_fail(ctx, FEMB_ERR_ARG, "femb_assemble_poisson_host: cellnodes differs from the connectivity the pattern was built from; call femb_mesh_set and rebuild the pattern");
    }
    return femb_sync(ctx);
}
