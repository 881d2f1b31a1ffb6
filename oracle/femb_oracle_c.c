/*
 * TEST INFRASTRUCTURE — plain-C restatement of the reference's serial assembly loop
 *   /root/reference/examples/Example200_LowLevelPoisson.jl:125-180  (barrier(): for cell in 1:ncells ...)
 * with the per-cell evaluator update it calls
 *   /root/reference/src/feevaluator.jl:344-353       (_update_trafo!: update_trafo! + mapderiv!)
 *   /root/reference/src/feevaluator_h1.jl:61-74      (update_basis! Gradient: 5-deep loop)
 * and the sparse insert into an EXISTING (flushed) CSC pattern
 *   ExtendableSparse rawupdateindex!(A, +, v, i, j)  [EXT E2]: binary search in column j, in-place add
 *   (steady state of the reference: second assembly, zero allocations, Example200:205-211).
 *
 * Used (a) by tests/ as a second, independently written oracle (serial mode has the reference's exact
 * summation order), (b) by bench.py as the CPU baseline ("port") on the GPU box's host cores.  The
 * reference loop itself is serial; the threaded mode (OpenMP over cells, atomic adds) is what a
 * Threads.@threads variant with FEMatrix(...; npartitions) would do and is labelled as such.
 * Never linked into libfemb.so.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXND 32
#define MAXQP 32

/* rawupdateindex!(A, +, v, i, j) on the CSC part: 1-based i, j; returns 0 if (i,j) is not in the pattern */
static inline int csc_add(const int64_t* colptr, const int64_t* rowval, double* nzval, int64_t i, int64_t j, double v, int atomic) {
    int64_t lo = colptr[j - 1] - 1, hi = colptr[j] - 1;
    const int64_t end = hi;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rowval[mid] < i) lo = mid + 1;
        else hi = mid;
    }
    if (lo >= end || rowval[lo] != i) return 0;
    if (atomic) {
#pragma omp atomic
        nzval[lo] += v;
    } else {
        nzval[lo] += v;
    }
    return 1;
}

int femb_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * Example200.assemble! for a scalar H1 element without subset handler (H1Pk order 1, 2; H1P1; H1P2) on
 * Triangle2D / Tetrahedron3D, cells [cell_begin, cell_end) (0-based), rhs(x) = rhs[0] + sum_d rhs[1+d] x_d
 * (NULL = no rhs).  refderivs[qp][dof][j], refvals[qp][dof].  Returns the number of contributions that
 * found no slot (must be 0), or -1 on bad arguments.
 */
int64_t femb_oracle_example200(int dim, int64_t ncells, const double* coords, const int32_t* cellnodes, const double* cellvol,
                               const int32_t* celldofs, int nd, int nqp, const double* w, const double* xref, const double* refvals,
                               const double* refderivs, double mu, const double* rhs, double drop_tol, const int64_t* colptr,
                               const int64_t* rowval, double* nzval, double* b, int nthreads, int64_t cell_begin, int64_t cell_end) {
    if ((dim != 2 && dim != 3) || nd > MAXND || nqp > MAXQP || cell_end > ncells) return -1;
    int64_t missing = 0;
    const int atomic = nthreads > 1;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(+ : missing)
#endif
    {
        double cvals[3][MAXND][MAXQP]; /* FEB.cvals[k, dof, qp] */
        double Aloc[MAXND][MAXND];
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t cell = cell_begin; cell < cell_end; cell++) {
            /* update_trafo!: A = [x2-x1 | x3-x1 | ...], b = x1 */
            double A[3][3], M[3][3], x0[3], det;
            const int32_t* cn = cellnodes + cell * (dim + 1);
            const double* p0 = coords + (int64_t)(cn[0] - 1) * dim;
            for (int k = 0; k < dim; k++) {
                x0[k] = p0[k];
                for (int i = 0; i < dim; i++) A[k][i] = coords[(int64_t)(cn[i + 1] - 1) * dim + k] - p0[k];
            }
            /* mapderiv!: M = A^{-T}, det */
            if (dim == 2) {
                det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
                M[0][0] = A[1][1] / det;
                M[0][1] = -A[1][0] / det;
                M[1][0] = -A[0][1] / det;
                M[1][1] = A[0][0] / det;
            } else {
                double C[3][3];
                for (int j = 0; j < 3; j++)
                    for (int k = 0; k < 3; k++) {
                        int j1 = (j + 1) % 3, j2 = (j + 2) % 3, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
                        C[j][k] = A[j1][k1] * A[j2][k2] - A[j1][k2] * A[j2][k1];
                    }
                det = A[0][0] * C[0][0] + A[0][1] * C[0][1] + A[0][2] * C[0][2];
                for (int j = 0; j < 3; j++)
                    for (int k = 0; k < 3; k++) M[j][k] = C[j][k] / det;
            }
            /* update_basis! (Gradient): fill!(cvals, 0); cvals[k,dof,qp] += M[k,j] * refderiv[dof,j,qp] */
            for (int qp = 0; qp < nqp; qp++)
                for (int dof = 0; dof < nd; dof++)
                    for (int k = 0; k < dim; k++) {
                        double acc = 0.0;
                        for (int j = 0; j < dim; j++) acc += M[k][j] * refderivs[((int64_t)qp * nd + dof) * dim + j];
                        cvals[k][dof][qp] = acc;
                    }
            /* local stiffness, upper triangle (Example200:139-146) */
            const double vol = cellvol[cell];
            for (int j = 0; j < nd; j++)
                for (int k = j; k < nd; k++) {
                    double temp = 0.0;
                    for (int qp = 0; qp < nqp; qp++) {
                        double dot = 0.0;
                        for (int r = 0; r < dim; r++) dot += cvals[r][j][qp] * cvals[r][k][qp];
                        temp += w[qp] * dot;
                    }
                    Aloc[j][k] = temp * (mu * vol);
                }
            /* scatter (Example200:149-161) */
            const int32_t* cd = celldofs + cell * nd;
            for (int j = 0; j < nd; j++)
                for (int k = j; k < nd; k++) {
                    const double v = Aloc[j][k];
                    if (fabs(v) > drop_tol) {
                        missing += !csc_add(colptr, rowval, nzval, cd[j], cd[k], v, atomic);
                        if (k > j) missing += !csc_add(colptr, rowval, nzval, cd[k], cd[j], v, atomic);
                    }
                }
            /* right-hand side (Example200:165-178) */
            if (rhs && b) {
                for (int j = 0; j < nd; j++) {
                    double temp = 0.0;
                    for (int qp = 0; qp < nqp; qp++) {
                        double f = rhs[0];
                        for (int k = 0; k < dim; k++) {
                            double xk = x0[k]; /* eval_trafo!: x = A xref + b */
                            for (int i = 0; i < dim; i++) xk += A[k][i] * xref[qp * dim + i];
                            f += rhs[1 + k] * xk;
                        }
                        temp += w[qp] * refvals[qp * nd + j] * f;
                    }
                    const double v = temp * vol;
                    if (atomic) {
#pragma omp atomic
                        b[cd[j] - 1] += v;
                    } else {
                        b[cd[j] - 1] += v;
                    }
                }
            }
        }
    }
    return missing;
}
