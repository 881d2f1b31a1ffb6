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

// FEMB_ERR_UNSUPPORTED (no message) when the closed form does not apply
static int launch_th_divpressure(femb_ctx* ctx, int egu, int eip, double factor) {
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
            st = launch_th_divpressure(ctx, egu, eip, -1.0);  // closed form for the P2 / P1 pair on triangles
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
