// Device-side building blocks shared by the assembly kernels: stage 1 (affine map), the per-cell
// handlers, stage 2 (push-forward / Piola) and the rhs functions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "femb_internal.h"

// ---- stage 1: L2GTransformer ([EXT E1]; call sites src/feevaluator.jl:344-363) --------------------
// A[:,i] = x_{i+1} - x_1, b = x_1; M = A^{-T} (mapderiv!), det = det(A); |T| = |det|/d! unless given.
template <int DIM>
struct Geo {
    double A[DIM][DIM];  // forward Jacobian (Piola)
    double M[DIM][DIM];  // A^{-T}: d/dx_k = sum_j M[k][j] d/dxref_j
    double x0[DIM];
    double det;
    double vol;
};

template <int DIM>
__device__ __forceinline__ void load_nodes(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes,
                                           int64_t cell, double (&x)[DIM + 1][DIM]) {
#pragma unroll
    for (int i = 0; i <= DIM; i++) {
        int64_t n = (int64_t)__ldg(cellnodes + cell * (DIM + 1) + i) - 1;
        if (DIM == 2) {
            double2 p = __ldg(reinterpret_cast<const double2*>(coords) + n);
            x[i][0] = p.x;
            x[i][1] = p.y;
        } else {
#pragma unroll
            for (int d = 0; d < DIM; d++) x[i][d] = __ldg(coords + n * DIM + d);
        }
    }
}

template <int DIM>
__device__ __forceinline__ void make_geo(const double (&x)[DIM + 1][DIM], const double* __restrict__ cellvol, int64_t cell,
                                         Geo<DIM>& g) {
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        g.x0[d] = x[0][d];
#pragma unroll
        for (int i = 0; i < DIM; i++) g.A[d][i] = x[i + 1][d] - x[0][d];
    }
    if (DIM == 2) {
        double det = g.A[0][0] * g.A[1][1] - g.A[0][1] * g.A[1][0];
        g.det = det;
        g.M[0][0] = g.A[1][1] / det;
        g.M[0][1] = -g.A[1][0] / det;
        g.M[1][0] = -g.A[0][1] / det;
        g.M[1][1] = g.A[0][0] / det;
        g.vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * 0.5;
    } else {
        // cofactors C[j][k]; M = C / det
        double c00 = g.A[1][1] * g.A[2][2] - g.A[1][2] * g.A[2][1];
        double c01 = g.A[1][2] * g.A[2][0] - g.A[1][0] * g.A[2][2];
        double c02 = g.A[1][0] * g.A[2][1] - g.A[1][1] * g.A[2][0];
        double c10 = g.A[0][2] * g.A[2][1] - g.A[0][1] * g.A[2][2];
        double c11 = g.A[0][0] * g.A[2][2] - g.A[0][2] * g.A[2][0];
        double c12 = g.A[0][1] * g.A[2][0] - g.A[0][0] * g.A[2][1];
        double c20 = g.A[0][1] * g.A[1][2] - g.A[0][2] * g.A[1][1];
        double c21 = g.A[0][2] * g.A[1][0] - g.A[0][0] * g.A[1][2];
        double c22 = g.A[0][0] * g.A[1][1] - g.A[0][1] * g.A[1][0];
        double det = g.A[0][0] * c00 + g.A[0][1] * c01 + g.A[0][2] * c02;
        g.det = det;
        g.M[0][0] = c00 / det; g.M[0][1] = c01 / det; g.M[0][2] = c02 / det;
        g.M[1][0] = c10 / det; g.M[1][1] = c11 / det; g.M[1][2] = c12 / det;
        g.M[2][0] = c20 / det; g.M[2][1] = c21 / det; g.M[2][2] = c22 / det;
        g.vol = cellvol ? __ldg(cellvol + cell) : fabs(det) * (1.0 / 6.0);
    }
}

template <int DIM>
__device__ __forceinline__ void load_geo(const double* __restrict__ coords, const int32_t* __restrict__ cellnodes,
                                         const double* __restrict__ cellvol, int64_t cell, Geo<DIM>& g) {
    double x[DIM + 1][DIM];
    load_nodes<DIM>(coords, cellnodes, cell, x);
    make_geo<DIM>(x, cellvol, cell, g);
}

// eval_trafo!(x, L2G, xref) = A xref + b
template <int DIM>
__device__ __forceinline__ void eval_trafo(const Geo<DIM>& g, const double* xr, double* x) {
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        double s = g.x0[d];
#pragma unroll
        for (int i = 0; i < DIM; i++) s += g.A[d][i] * xr[i];
        x[d] = s;
    }
}

// ---- per-cell handlers (get_basissubset / get_coefficients) ----------------------------------------
struct SpaceDev {
    int family, ncomp, nd, nd_all, handler, nsign;
    const int32_t* celldofs;
    const int32_t* signs;
    const int32_t* orient;
};

// subset[dof] (0-based in, 0-based out)
__device__ __forceinline__ int subset_of(const SpaceDev& s, int64_t cell, int dof) {
    switch (s.handler) {
        case FEMB_HANDLER_H1_FACESWAP2D: {  // h1_pk.jl:280-304, h1_p3.jl:200-217
            int nds = s.nd / s.ncomp;
            int c = dof / nds, l = dof - c * nds;
            int order = (nds == 10) ? 3 : (nds == 15) ? 4 : (nds == 21) ? 5 : 2;
            int nf = order - 1;
            if (l >= 3 && l < 3 + 3 * nf) {
                int j = (l - 3) / nf, k = (l - 3) - j * nf;
                if (__ldg(s.signs + cell * 3 + j) != 1) l = 3 + j * nf + (nf - 1 - k);
            }
            return c * nds + l;
        }
        case FEMB_HANDLER_H1_EDGESWAP3D: {  // h1_p3.jl:220-241
            int c = dof / 20, l = dof - c * 20;
            if (l >= 4 && l < 16) {
                int j = (l - 4) >> 1;
                if (__ldg(s.signs + cell * 6 + j) != 1) l ^= 1;  // 4+2j <-> 4+2j+1
            }
            return c * 20 + l;
        }
        case FEMB_HANDLER_BDM1_3D: {  // hdiv_bdm1.jl:234-249
            int j = dof / 3, m = dof - 3 * j;
            if (m == 0) return 4 * j;
            int o = __ldg(s.orient + cell * 4 + j);  // 1..4
            // 1-based: 4(j+1) - shift[o]; shift1 = [1,0,1,2], shift2 = [2,2,0,1]
            int sh = (m == 1) ? ((0x2101 >> (4 * (o - 1))) & 0xF) : ((0x1022 >> (4 * (o - 1))) & 0xF);
            return 4 * (j + 1) - sh - 1;
        }
        default:
            return dof;
    }
}

// coefficients[comp, dof] (same for all components in every on-path element)
__device__ __forceinline__ double coef_of(const SpaceDev& s, int64_t cell, int dof) {
    switch (s.handler) {
        case FEMB_HANDLER_RT0_SIGNS:  // hdiv_rt0.jl:114-124
            return (double)__ldg(s.signs + cell * s.nsign + dof);
        case FEMB_HANDLER_N0_EDGESIGNS3D:  // hcurl_n0.jl:156-166
            return (double)__ldg(s.signs + cell * 6 + dof);
        case FEMB_HANDLER_BDM1_2D:  // hdiv_bdm1.jl:200-212
            return (dof & 1) ? 1.0 : (double)__ldg(s.signs + cell * 3 + (dof >> 1));
        case FEMB_HANDLER_BDM1_3D: {  // hdiv_bdm1.jl:215-228
            int j = dof / 3, m = dof - 3 * j;
            return m == 0 ? (double)__ldg(s.signs + cell * 4 + j) : (m == 1 ? -1.0 : 1.0);
        }
        default:
            return 1.0;
    }
}

// ---- stage 2: update_basis! for one (dof, qp) ------------------------------------------------------
// refvals  [qp][nd_all][ncomp]; refderivs [qp][nd_all][ncomp][DIM]
// out[r], r < resultdim.  Accumulation order follows the reference loops.
// RT > 0: the caller knows the result length at compile time (k_bilinear_rows): every index into `out` is then static
// and the array lives in registers instead of local memory.
template <int DIM, int RT = 0>
__device__ __forceinline__ void eval_op(const SpaceDev& s, int op, const Geo<DIM>& g, int64_t cell, int dof, int qp,
                                        const double* __restrict__ refvals, const double* __restrict__ refderivs,
                                        double* out) {
    const int sub = subset_of(s, cell, dof);
    const int nc = s.ncomp;
    if (s.family != FEMB_FAMILY_HDIV && s.family != FEMB_FAMILY_HCURL) {
        if (op == FEMB_OP_IDENTITY) {  // feevaluator_h1.jl:2-14
            const double* v = refvals + ((size_t)qp * s.nd_all + sub) * nc;
            if (RT > 0) {
#pragma unroll
                for (int k = 0; k < RT; k++) out[k] = v[k];
            } else {
                for (int k = 0; k < nc; k++) out[k] = v[k];
            }
        } else if (op == FEMB_OP_GRADIENT) {  // feevaluator_h1.jl:61-74
            const double* d = refderivs + ((size_t)qp * s.nd_all + sub) * nc * DIM;
            if (RT > 0) {
#pragma unroll
                for (int c = 0; c < RT / DIM; c++) {
#pragma unroll
                    for (int k = 0; k < DIM; k++) {
                        double a = 0.0;
#pragma unroll
                        for (int j = 0; j < DIM; j++) a += g.M[k][j] * d[c * DIM + j];
                        out[k + DIM * c] = a;
                    }
                }
            } else
            for (int c = 0; c < nc; c++) {
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    double a = 0.0;
#pragma unroll
                    for (int j = 0; j < DIM; j++) a += g.M[k][j] * d[c * DIM + j];
                    out[k + DIM * c] = a;
                }
            }
        } else if (op == FEMB_OP_DIVERGENCE) {  // feevaluator_h1.jl:119-130
            const double* d = refderivs + ((size_t)qp * s.nd_all + sub) * nc * DIM;
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; k++)
#pragma unroll
                for (int j = 0; j < DIM; j++) a += g.M[k][j] * d[k * DIM + j];
            out[0] = a;
        } else if (op == FEMB_OP_GRADTRACE) {  // Example210:296: grad[1] + grad[4]
            const double* d = refderivs + ((size_t)qp * s.nd_all + sub) * nc * DIM;
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < DIM; j++) a += g.M[k][j] * d[k * DIM + j];
                t = (k == 0) ? a : t + a;
            }
            out[0] = t;
        }
    } else if (s.family == FEMB_FAMILY_HCURL) {
        if (op == FEMB_OP_IDENTITY) {  // feevaluator_hcurl.jl:2-16: covariant Piola map, (sum_l L2GAinv[k,l] phi[sub,l]) * coeff
            const double cf = coef_of(s, cell, dof);
            const double* v = refvals + ((size_t)qp * s.nd_all + sub) * DIM;
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                double a = 0.0;
#pragma unroll
                for (int l = 0; l < DIM; l++) a += g.M[k][l] * v[l];
                if (RT == 0 || k < RT) out[k] = a * cf;
            }
        }
    } else {
        const double cf = coef_of(s, cell, dof);
        if (op == FEMB_OP_IDENTITY) {  // feevaluator_hdiv.jl:2-19
            const double* v = refvals + ((size_t)qp * s.nd_all + sub) * DIM;
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                double a = 0.0;
#pragma unroll
                for (int l = 0; l < DIM; l++) a += g.A[k][l] * v[l];
                if (RT == 0 || k < RT) out[k] = a * (cf / g.det);
            }
        } else if (op == FEMB_OP_DIVERGENCE) {  // feevaluator_hdiv.jl:66-83
            const double* d = refderivs + ((size_t)qp * s.nd_all + sub) * DIM * DIM;
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < DIM; j++) a += d[j * DIM + j];
            out[0] = a * (cf / g.det);
        }
    }
}

__host__ __device__ inline int result_dim(int op, int dim, int ncomp) {
    // Length4Operator (functionoperators.jl:182-196)
    switch (op) {
        case FEMB_OP_IDENTITY: return ncomp;
        case FEMB_OP_GRADIENT: return dim * ncomp;
        case FEMB_OP_DIVERGENCE: return (ncomp + dim - 1) / dim;
        case FEMB_OP_GRADTRACE: return 1;
        case FEMB_OP_CONVECTION: return ncomp;
        default: return 0;
    }
}

// ---- rhs data -----------------------------------------------------------------------------------------
struct RhsDev {
    int kind;
    int ncomp;
    double p[16];         // affine coefficients or {mu, t}
    const double* fvals;  // [cell][qp][comp]
};

// NC > 0: the caller knows the number of components at compile time (every index into f is then static: registers, not local memory)
template <int DIM, int NC = 0>
__device__ __forceinline__ void eval_rhs(const RhsDev& r, const double* x, int64_t cell, int qp, int nqp, double* f) {
    const int nc = NC > 0 ? NC : r.ncomp;
    if (r.kind == FEMB_RHS_AFFINE) {
#pragma unroll
        for (int c = 0; c < nc; c++) {
            double s = r.p[c * (DIM + 1)];
#pragma unroll
            for (int d = 0; d < DIM; d++) s += r.p[c * (DIM + 1) + 1 + d] * x[d];
            f[c] = s;
        }
    } else if (r.kind == FEMB_RHS_QPVALUES) {
#pragma unroll
        for (int c = 0; c < nc; c++) f[c] = __ldg(r.fvals + ((size_t)cell * nqp + qp) * nc + c);
    } else if (r.kind == FEMB_RHS_EX210) {  // Example210:44-48
        const double pi = 3.14159265358979323846;
        double mu = r.p[0], t = r.p[1];
        double e = 8.0 * pi * pi * mu * exp(-8.0 * pi * pi * mu * t);
        if (NC == 0 || NC >= 2) {
            f[0] = e * sin(2.0 * pi * x[0]) * sin(2.0 * pi * x[1]);
            f[NC == 1 ? 0 : 1] = e * cos(2.0 * pi * x[0]) * cos(2.0 * pi * x[1]);
        } else {
            f[0] = 0.0;  // (a two-component load)
        }
    } else {
#pragma unroll
        for (int c = 0; c < nc; c++) f[c] = 0.0;
    }
}
