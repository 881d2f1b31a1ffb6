// Example210 (Taylor-Hood Navier-Stokes): Newton-linearised convection kernels and the update_system! entry point.
#include <cstdlib>

#include "femb_kernels.cuh"

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
// One thread per (cell, 6x6 block): 36 accumulators in registers, no shared-memory staging of cvals.  The kernel is
// latency-bound at 2 CTAs / SM (180 registers); capping it at 168 (3 CTAs / SM, 24 bytes of spill) measured 2.38 vs 2.91 ms
// on 2.1M cells, 128 registers (4 CTAs, 400 bytes of spill) 3.98 ms.
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
__global__ void __launch_bounds__(128, 3) k_convection2d_blocks(const Conv2Args a) {
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

// ---------------------------------------------------------------------------------------------------
// Example210 div-pressure blocks in closed form (Example210:293-301,376-380): Bloc[(c,j),k] = -|T| int d_c phi_j lambda_k for the
// component-blocked P2 velocity / P1 pressure pair on triangles.  With D_ca = |T| d_c lambda_a:
//   vertex function i:  -(4 m_ik - 1/3) D_ci          (= -D_ck / 3 for i = k, 0 otherwise)
//   edge function (ab): -4 (m_ak D_cb + m_bk D_ca),   m = 1/6 (equal indices) or 1/12
// One thread per cell, 36 values, scattered into (u,p) and (p,u) through the slot maps.  The launcher accepts the closed form
// only if it reproduces the evaluators' quadrature tables on a test cell; else the generic kernel runs.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_th_divpressure2d(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol,
                                                          int64_t ncells, double factor, const int32_t* __restrict__ slots, const int32_t* __restrict__ slots_t,
                                                          double* __restrict__ nzval, int32_t* __restrict__ errflag) {
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cell >= ncells) return;
    double x[3][2];
    load_nodes<2>(coords, cellnodes, cell, x);
    const double a00 = x[1][0] - x[0][0], a01 = x[2][0] - x[0][0], a10 = x[1][1] - x[0][1], a11 = x[2][1] - x[0][1];
    const double det = a00 * a11 - a01 * a10;
    const double vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * 0.5;
    const double sc = vol / det;
    // det * grad lambda_1 = (a11, -a01), det * grad lambda_2 = (-a10, a00), grad lambda_0 = -(sum)
    double D[2][3];
    D[0][1] = a11 * sc; D[1][1] = -a01 * sc;
    D[0][2] = -a10 * sc; D[1][2] = a00 * sc;
    D[0][0] = -(D[0][1] + D[0][2]); D[1][0] = -(D[1][1] + D[1][2]);
    unsigned err = 0u;
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int j = 0; j < 6; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                double v;
                if (j < 3) {
                    if (j != k) continue;  // (exact zero: the reference adds 0.0 there)
                    v = D[c][k] * (1.0 / 3.0);
                } else {
                    const int ea = (j == 3) ? 0 : ((j == 4) ? 1 : 0), eb = (j == 3) ? 1 : 2;  // local edges (0 1), (1 2), (0 2)
                    const double ma = (ea == k) ? (1.0 / 6.0) : (1.0 / 12.0), mb = (eb == k) ? (1.0 / 6.0) : (1.0 / 12.0);
                    v = 4.0 * (ma * D[c][eb] + mb * D[c][ea]);
                }
                v *= factor;
                const int32_t s = __ldg(slots + (cell * 12 + c * 6 + j) * 3 + k);
                const int32_t t = __ldg(slots_t + (cell * 3 + k) * 12 + c * 6 + j);
                if (s < 0 || t < 0) err = 1u;
                else {
                    atomicAdd(nzval + s, v);
                    atomicAdd(nzval + t, v);
                }
            }
    if (err) atomicAdd(errflag, 1);
}

// ---------------------------------------------------------------------------------------------------
// The same blocks written by the column owners (no reductions, every entry stored once), for a matrix that was zero when
// femb_assemble_navierstokes started.  No map beyond the component-diagonal gather map of the velocity block is needed: the vertex dofs
// of P2 are the dofs of the P1 pressure, hence
//   * the pressure rows [p rows] that follow the [ux rows | uy rows] of a velocity column are that column's vertex rows in the same
//     order: the position bytes the map holds for local dofs 0..2 of every visit address them;
//   * the pressure column of node v has the rows, visits and positions of the velocity column of vertex dof v (ux part, and the uy
//     part one run length further down).
// One thread per scalar velocity dof j (both components: twins): (p, ux_j), (p, uy_j) and, for vertex dofs, column p_j.
// Records: |T| d_c lambda_a for a = 1, 2 (one sector per cell).  Map words in role order (femb_p2.cu: perm[k][role] = local dof).
// dynamic shared memory: [4 mbarriers | pad to 128 B][acc: nacc x R rows x 128 doubles][4 warps x 3 planes x maxrows x 128 B],
// nacc = 4 (vertex-dof slices: P_x, P_y, Q_x, Q_y) or 2.
// ---------------------------------------------------------------------------------------------------
struct ThGArgs {
    const double* rec;  // [cell][4] = D[0][1], D[1][1], D[0][2], D[1][2]
    const int32_t* slice_ptr;
    const uint32_t* map;
    size_t stride;
    const int64_t* colptr;  // of the whole matrix
    const int32_t* sublen;  // run length of the component-diagonal block per velocity column
    int64_t n;              // scalar dofs per component
    int64_t nnodes;         // vertex dofs = pressure dofs
    int64_t ucol0, pcol0;   // first velocity / pressure column of the matrix
    double* nzval;
    double factor;
    int64_t slice_begin, slice_end;
    int R, nacc, maxrows;
    uint8_t perm[6][6];
};

__device__ __forceinline__ double th_pick(int a, double x0, double x1, double x2) { return a == 0 ? x0 : (a == 1 ? x1 : x2); }

// -|T| int d_c phi_i lambda_a without the factor, D = (D_c0, D_c1, D_c2): the arithmetic of k_th_divpressure2d
__device__ __forceinline__ double th_value(int i, int a, double d0, double d1, double d2) {
    if (i < 3) return i == a ? th_pick(a, d0, d1, d2) * (1.0 / 3.0) : 0.0;
    const int ea = (i == 3) ? 0 : ((i == 4) ? 1 : 0), eb = (i == 3) ? 1 : 2;
    const double ma = (ea == a) ? (1.0 / 6.0) : (1.0 / 12.0), mb = (eb == a) ? (1.0 / 6.0) : (1.0 / 12.0);
    return 4.0 * (ma * th_pick(eb, d0, d1, d2) + mb * th_pick(ea, d0, d1, d2));
}

__global__ void __launch_bounds__(128) k_th_rec(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, const double* __restrict__ cellvol, int64_t ncells,
                                                double* __restrict__ rec) {
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cell >= ncells) return;
    double x[3][2];
    load_nodes<2>(coords, cellnodes, cell, x);
    const double a00 = x[1][0] - x[0][0], a01 = x[2][0] - x[0][0], a10 = x[1][1] - x[0][1], a11 = x[2][1] - x[0][1];
    const double det = a00 * a11 - a01 * a10;
    const double vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * 0.5;
    const double sc = vol / det;
    double* o = rec + cell * 4;
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(o), "d"(a11 * sc), "d"(-a01 * sc), "d"(-a10 * sc), "d"(a00 * sc) : "memory");
}

// what the gather takes for granted about the columns, checked once per pattern: velocity column = [ux rows | uy rows | p rows] with
// equally long ux / uy parts (= the run of the component-diagonal block), the vertex rows first in a run and as many as there are
// pressure rows; pressure column of node v = [ux rows | uy rows | ...] with the run length of velocity column v
__global__ void k_th_layout_check(const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, const int32_t* __restrict__ sub0, const int32_t* __restrict__ sublen,
                                  int64_t n, int64_t nnodes, int64_t ucol0, int64_t pcol0, int64_t urow0, int64_t prow0, int32_t* __restrict__ bad) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    bool ok = true;
    const int L = sublen[j];
    ok = ok && sublen[j + n] == L && sub0[j] == 0 && sub0[j + n] == L;
    for (int c = 0; c < 2 && ok; c++) {
        const int64_t b = colptr[ucol0 + j + c * n] - 1, e = colptr[ucol0 + j + c * n + 1] - 1;
        const int np = (int)(e - b) - 2 * L;
        ok = ok && np >= 0 && np <= L;
        if (!ok) break;
        // rows (1-based): ux part in (urow0, urow0 + n], uy part in (urow0 + n, urow0 + 2n], p part in (prow0, prow0 + nnodes]
        if (L > 0) ok = ok && rowval[b] > urow0 && rowval[b + L - 1] <= urow0 + n && rowval[b + L] > urow0 + n && rowval[b + 2 * L - 1] <= urow0 + 2 * n;
        if (np > 0) ok = ok && rowval[b + 2 * L] > prow0 && rowval[e - 1] <= prow0 + nnodes;
        // vertex rows of the run: exactly np, at its start
        if (np > 0) ok = ok && rowval[b + c * L + np - 1] <= urow0 + c * n + nnodes;
        if (np < L) ok = ok && rowval[b + c * L + np] > urow0 + c * n + nnodes;
    }
    if (ok && j < nnodes) {
        const int64_t b = colptr[pcol0 + j] - 1, e = colptr[pcol0 + j + 1] - 1;
        ok = ok && (e - b) >= 2 * L;
        if (ok && L > 0) ok = ok && rowval[b] > urow0 && rowval[b + L - 1] <= urow0 + n && rowval[b + L] > urow0 + n && rowval[b + 2 * L - 1] <= urow0 + 2 * n;
        if (ok && (e - b) > 2 * L) ok = ok && rowval[b + 2 * L] > urow0 + 2 * n;
    }
    if (!ok) atomicAdd(bad, 1);
}

__device__ __forceinline__ unsigned th_lds32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(128) k_th_divp_gather(const ThGArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int64_t slice = a.slice_begin + (int64_t)blockIdx.x * 4 + warp;
    if (slice >= a.slice_end) return;
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned pst = (unsigned)a.maxrows * 128u;
    const unsigned mbar = smem0 + warp * 8u;
    const unsigned plane = 128u * 8u;                      // one accumulator row: 128 doubles
    const unsigned asz = (unsigned)a.R * plane;            // one accumulator array
    const unsigned sstw = smem0 + 128u + (unsigned)a.nacc * asz + warp * 3u * pst;
    const int64_t j = slice * 32 + lane;
    const bool valid = j < a.n;
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
    const int L = valid ? __ldg(a.sublen + j) : 0;
    const bool isv = valid && j < a.nnodes && a.nacc == 4;
    int64_t ub[2] = {0, 0};
    int np[2] = {0, 0};
    if (valid) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            ub[c] = a.colptr[a.ucol0 + j + c * a.n] - 1;
            np[c] = (int)(a.colptr[a.ucol0 + j + c * a.n + 1] - 1 - ub[c]) - 2 * L;
        }
    }
    const unsigned sP = smem0 + 128u + tid * 8u;  // P_c at sP + c * asz, Q_c at sP + (2 + c) * asz
    for (int i = 0; i < max(np[0], np[1]); i++) {
        sts64(sP + i * plane, 0.0);
        sts64(sP + asz + i * plane, 0.0);
    }
    if (isv)
        for (int i = 0; i < L; i++) {
            sts64(sP + 2u * asz + i * plane, 0.0);
            sts64(sP + 3u * asz + i * plane, 0.0);
        }
    __syncwarp();
    const unsigned sst = sstw + lane * 4u;
    if (nrows > 0) {
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
    }
    for (int r = 0; r < nrows; r++) {
        const unsigned ra = sst + (unsigned)r * 128u;
        const unsigned cell = th_lds32(ra), wa = th_lds32(ra + pst), wb = th_lds32(ra + 2u * pst);
        const unsigned k = (wb >> 16) & 0xFu;
        if (k == 0xFu || !valid) continue;
        double d01, d11, d02, d12;
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(d01), "=d"(d11), "=d"(d02), "=d"(d12) : "l"(a.rec + (size_t)cell * 4));
        const double d00 = -(d01 + d02), d10 = -(d11 + d12);
#pragma unroll
        for (int rr = 0; rr < 6; rr++) {
            const unsigned pos = ((rr < 4 ? wa : wb) >> (8 * (rr & 3))) & 0xFFu;
            const int il = a.perm[k][rr];  // local dof of the cell in this role
            if (il < 3 && pos < (unsigned)a.R) {
                // pressure row of local vertex il in the columns (ux_j, uy_j): -|T| int d_c phi_k lambda_il
                const unsigned ad = sP + pos * plane;
                sts64(ad, lds64(ad) + th_value((int)k, il, d00, d01, d02) * a.factor);
                sts64(ad + asz, lds64(ad + asz) + th_value((int)k, il, d10, d11, d12) * a.factor);
            }
            if (isv && pos < (unsigned)a.R) {
                // pressure column of this vertex (local index k): rows ux / uy of local dof il
                const unsigned ad = sP + 2u * asz + pos * plane;
                sts64(ad, lds64(ad) + th_value(il, (int)k, d00, d01, d02) * a.factor);
                sts64(ad + asz, lds64(ad + asz) + th_value(il, (int)k, d10, d11, d12) * a.factor);
            }
        }
    }
    if (!valid) return;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        double* out = a.nzval + ub[c] + 2 * L;
        for (int i = 0; i < np[c]; i++) out[i] = lds64(sP + (unsigned)c * asz + i * plane);
    }
    if (isv) {
        double* out = a.nzval + (a.colptr[a.pcol0 + j] - 1);
        for (int i = 0; i < L; i++) out[i] = lds64(sP + 2u * asz + i * plane);
        for (int i = 0; i < L; i++) out[L + i] = lds64(sP + 3u * asz + i * plane);
    }
}

// FEMB_ERR_UNSUPPORTED (no message) when the closed form does not apply
static int launch_th_divp_gather(femb_ctx* ctx, const FembSpace& us, const FembSpace& ps, double factor);

static int launch_th_divpressure(femb_ctx* ctx, int egu, int eip, double factor, bool zero_at_entry) {
    static const bool on = [] { const char* e = getenv("FEMB_TH_CLOSED"); return !(e && e[0] == '0'); }();
    const FembEval& EG = ctx->evals[egu];
    const FembEval& EP = ctx->evals[eip];
    const FembSpace& us = ctx->spaces[EG.space];
    const FembSpace& ps = ctx->spaces[EP.space];
    if (!on || ctx->dim != 2 || ctx->scatter == FEMB_SCATTER_COLORED || ctx->touched || us.family != FEMB_FAMILY_H1 || us.ncomp != 2 || us.nd != 12 || us.nd_all != 12 ||
        us.handler != FEMB_HANDLER_NONE || ps.family != FEMB_FAMILY_H1 || ps.ncomp != 1 || ps.nd != 3 || ps.nd_all != 3 || EG.nqp != EP.nqp)
        return FEMB_ERR_UNSUPPORTED;
    if (EG.h_refderivs.size() != (size_t)EG.nqp * 12 * 4 || EP.h_refvals.size() != (size_t)EP.nqp * 3) return FEMB_ERR_UNSUPPORTED;
    {
        // reference cell check: sum_q w_q (d_c phi_j)(q) lambda_k(q) for the first-component functions against the constants above
        for (int j = 0; j < 6; j++)
            for (int k = 0; k < 3; k++)
                for (int c = 0; c < 2; c++) {
                    double ref = 0.0;
                    for (int q = 0; q < EG.nqp; q++) ref += EG.w[q] * EG.h_refderivs[(((size_t)q * 12 + j) * 2 + 0) * 2 + c] * EP.h_refvals[(size_t)q * 3 + k];
                    // on the reference triangle: d_c lambda_1 = (1,0), d_c lambda_2 = (0,1), d_c lambda_0 = (-1,-1)
                    const double dl[3] = {-1.0, c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0};
                    double closed;
                    if (j < 3) closed = (j == k) ? dl[k] / 3.0 : 0.0;
                    else {
                        const int ea = (j == 3) ? 0 : ((j == 4) ? 1 : 0), eb = (j == 3) ? 1 : 2;
                        closed = 4.0 * (((ea == k) ? 1.0 / 6.0 : 1.0 / 12.0) * dl[eb] + ((eb == k) ? 1.0 / 6.0 : 1.0 / 12.0) * dl[ea]);
                    }
                    if (fabs(ref - closed) > 5e-12) return FEMB_ERR_UNSUPPORTED;
                    // the second-component functions must be the same scalar functions in component 1
                    double ref2 = 0.0;
                    for (int q = 0; q < EG.nqp; q++) ref2 += EG.w[q] * EG.h_refderivs[(((size_t)q * 12 + 6 + j) * 2 + 1) * 2 + c] * EP.h_refvals[(size_t)q * 3 + k];
                    if (fabs(ref2 - closed) > 5e-12) return FEMB_ERR_UNSUPPORTED;
                }
    }
    if (zero_at_entry && ctx->scatter == FEMB_SCATTER_GATHER) {
        // nothing but zeros in the (u,p) / (p,u) blocks so far: the column owners store them, no reductions
        const int st = launch_th_divp_gather(ctx, us, ps, factor);
        if (st != FEMB_ERR_UNSUPPORTED) return st;
    }
    FEMB_TRY(femb_build_slotmaps(ctx, 0, 1));
    FEMB_TRY(femb_build_slotmaps(ctx, 1, 0));
    const FembBlock* b01 = find_block(ctx, 0, 1);
    const FembBlock* b10 = find_block(ctx, 1, 0);
    if (!b01 || !b10 || !b01->slots || !b10->slots) return FEMB_ERR_UNSUPPORTED;
    FEMB_TRY(femb_flush_zero(ctx, true, false));
    if (ctx->ncells == 0) return FEMB_OK;
    k_th_divpressure2d<<<(unsigned)((ctx->ncells + 127) / 128), 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, factor, b01->slots, b10->slots,
                                                                                      ctx->nzval, ctx->errflag);
    KERNEL_CHECK(ctx);
    return FEMB_OK;
}

static int launch_th_divp_gather(femb_ctx* ctx, const FembSpace& us, const FembSpace& ps, double factor) {
    static const bool on = [] { const char* e = getenv("FEMB_TH_GATHER"); return !(e && e[0] == '0'); }();
    if (!on || ctx->ncells == 0 || !find_block(ctx, 0, 0) || !find_block(ctx, 0, 1) || !find_block(ctx, 1, 0)) return FEMB_ERR_UNSUPPORTED;
    int st = femb_build_hmap_vec(ctx);
    if (st != FEMB_OK) return st;
    FembHMap& H = ctx->hm_vec;
    if (H.space != ctx->row_spaces[0] || !H.complete || H.words != 2 || H.twins != 2 || !H.sub0 || !H.sublen) return FEMB_ERR_UNSUPPORTED;
    for (auto& r : H.ranges)
        if (r.split != 1) return FEMB_ERR_UNSUPPORTED;
    const int64_t n = us.ndofs / 2, nnodes = ps.ndofs;
    if (nnodes > n || nnodes != ctx->nnodes) return FEMB_ERR_UNSUPPORTED;
    if (H.th_state == 0) {
        int32_t* bad = nullptr;
        int32_t hbad = 0;
        CUDA_TRY(ctx, cudaMalloc(&bad, sizeof(int32_t)));
        cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(int32_t), ctx->stream);
        if (e == cudaSuccess) {
            k_th_layout_check<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->colptr, ctx->rowval, H.sub0, H.sublen, n, nnodes, ctx->col_off[0], ctx->col_off[1], ctx->row_off[0],
                                                                                   ctx->row_off[1], bad);
            e = cudaGetLastError();
            ctx->launches++;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hbad, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(bad);
        if (e != cudaSuccess) return femb_fail(ctx, FEMB_ERR_CUDA, "k_th_layout_check: %s", cudaGetErrorString(e));
        // (the stores cover the two blocks entirely by construction: the whole pressure run of every velocity column, both velocity runs of every
        // pressure column; the transposed blocks must at least agree in size)
        int64_t n01 = 0, n10 = 0;
        FEMB_TRY(femb_block_entries(ctx, 0, 1, &n01));
        FEMB_TRY(femb_block_entries(ctx, 1, 0, &n10));
        H.th_state = (hbad == 0 && n01 == n10) ? 1 : -1;
    }
    if (H.th_state != 1) return FEMB_ERR_UNSUPPORTED;
    if (H.cranges.empty()) {
        const int64_t nslices = (us.ndofs + 31) / 32;
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
    FEMB_TRY(femb_hmap_set_canon(ctx, H, true));
    if (!ctx->threc) FEMB_TRY(femb_alloc(ctx, &ctx->threc, (size_t)ctx->ncells * 4));
    FEMB_TRY(femb_flush_zero(ctx, true, false));  // (normally a no-op here: the Laplacian has settled the lazy zero)
    k_th_rec<<<(unsigned)((ctx->ncells + 127) / 128), 128, 0, ctx->stream>>>(ctx->coords, ctx->cellnodes, ctx->cellvol, ctx->ncells, ctx->threc);
    KERNEL_CHECK(ctx);
    ThGArgs a;
    a.rec = ctx->threc; a.slice_ptr = H.slice_ptr; a.map = H.map; a.stride = H.stride; a.colptr = ctx->colptr; a.sublen = H.sublen;
    a.n = n; a.nnodes = nnodes; a.ucol0 = ctx->col_off[0]; a.pcol0 = ctx->col_off[1]; a.nzval = ctx->nzval; a.factor = factor;
    uint8_t perm[12][12];
    femb_p2_role_perm(perm, 2);
    for (int k = 0; k < 6; k++)
        for (int r = 0; r < 6; r++) a.perm[k][r] = perm[k][r];
    const int64_t vend = (nnodes + 31) / 32, send = (n + 31) / 32;  // slices with vertex dofs (they also own the pressure columns) / all slices of component 0
    size_t maxsmem = 0;
    auto smem_of = [](const FembHMap::Range& r, int nacc) { return 128 + (size_t)nacc * (size_t)std::max(1, r.maxlen) * 1024 + 4 * 3 * (size_t)r.maxrows * 128; };
    for (auto& r : H.cranges) {
        if (r.s0 < vend) maxsmem = std::max(maxsmem, smem_of(r, 4));
        if (r.s1 > vend && r.s0 < send) maxsmem = std::max(maxsmem, smem_of(r, 2));
    }
    if (maxsmem > 200 * 1024) return FEMB_ERR_UNSUPPORTED;
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_th_divp_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxsmem));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_th_divp_gather, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (auto& r : H.cranges) {
        for (int part = 0; part < 2; part++) {
            const int64_t s0 = part == 0 ? r.s0 : std::max(r.s0, vend), s1 = part == 0 ? std::min(r.s1, vend) : std::min(r.s1, send);
            if (s1 <= s0) continue;
            a.slice_begin = s0;
            a.slice_end = s1;
            a.nacc = part == 0 ? 4 : 2;
            a.R = std::max(1, r.maxlen);  // (ranges of a block map carry the longest run of the block, not the longest column)
            a.maxrows = r.maxrows;
            k_th_divp_gather<<<(unsigned)((s1 - s0 + 3) / 4), 128, smem_of(r, a.nacc), ctx->stream>>>(a);
            KERNEL_CHECK(ctx);
        }
    }
    return FEMB_OK;
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
    const bool zero_at_entry = ctx->zero_pending_A && ctx->zero_mask == ~0ull && !ctx->dirty_A;  // femb_matrix_zero was the last thing that happened to A
    ctx->dirty_A = ctx->dirty_b = true;
    int st = femb_timer_begin(ctx);
    if (st == FEMB_OK && linear) {
        // Aloc[k,j] = mu * sum_qp w <grad_j, grad_k>, then * |T|                 (Example210:284-290,368)
        // component-blocked basis: the 12x12 local matrix is two copies of the scalar 6x6 P2 stiffness; its
        // off-diagonal blocks are exact zeros (their slots stay in the pattern, adding 0.0 changes nothing)
        // default policy: owner-computes gather per component (write-once per launch, bit-reproducible); else the DMMA scatter
        bool rhs_done = false;  // the load of the velocity space is fused into the Laplacian gather where that kernel covers it
        st = ctx->scatter == FEMB_SCATTER_GATHER ? launch_h1_gather_vec(ctx, egu, mu, eiu, &rhs, &rhs_done) : FEMB_ERR_UNSUPPORTED;
        if (st == FEMB_ERR_UNSUPPORTED) st = ctx->scatter == FEMB_SCATTER_COLORED ? FEMB_ERR_UNSUPPORTED : launch_stiffness_dmma(ctx, egu, mu, -1.0);
        if (st == FEMB_ERR_UNSUPPORTED) st = launch_bilinear(ctx, egu, egu, FEMB_OP_GRADIENT, FEMB_OP_GRADIENT, 0, 0, mu, 0, 0.0);
        // Bloc[j,k] = -sum_qp w (g[1,j]+g[4,j]) psi_k * |T| into (u,p) and (p,u) (Example210:293-301,376-380)
        if (st == FEMB_OK) {
            st = launch_th_divpressure(ctx, egu, eip, -1.0, zero_at_entry);  // closed form for the P2 / P1 pair on triangles
            if (st == FEMB_ERR_UNSUPPORTED) st = launch_bilinear(ctx, egu, eip, FEMB_OP_GRADTRACE, FEMB_OP_IDENTITY, 0, 1, -1.0, FEMB_BIL_TRANSPOSE_COPY, 0.0);
        }
        if (st == FEMB_OK && rhs.kind != FEMB_RHS_NONE && !rhs_done) st = launch_linear(ctx, eiu, 0, 1.0, rhs);
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
