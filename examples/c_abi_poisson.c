/* The C ABI used directly from C: Example200's assemble! for H1P1 on the unit square split into two triangles
 * (the same call sequence as julia/FEMBaseB200.jl and extendablefembase.jl_b200/assembly.py).
 *
 *   gcc -std=c11 -I include examples/c_abi_poisson.c -L extendablefembase.jl_b200 -lfemb -Wl,-rpath,$PWD/extendablefembase.jl_b200 -o c_abi_poisson
 *
 * On a box without a B200, femb_create fails with FEMB_ERR_CUDA and the program prints the library's message (there
 * is no CPU fallback).  Known answer: right-isosceles cells, mu = 1: A = [1 -.5 -.5 0; -.5 1 0 -.5;
 * -.5 0 1 -.5; 0 -.5 -.5 1] with the hypotenuse entries (1,4), (4,1) filtered out by Example200's 1e-15 test. */
#include <stdio.h>
#include <stdlib.h>

#include "femb.h"

#define CHECK(call)                                                             \
    do {                                                                        \
        int32_t st_ = (call);                                                   \
        if (st_ != FEMB_OK) {                                                   \
            fprintf(stderr, "%s -> %d: %s\n", #call, st_, femb_last_error(ctx)); \
            return st_ == FEMB_ERR_CUDA ? 77 : 1;                               \
        }                                                                       \
    } while (0)

int main(void) {
    femb_ctx* ctx = NULL;
    /* xgrid[Coordinates] (2 x 4), xgrid[CellNodes] (3 x 2), 1-based, column-major as Julia stores them */
    const double coords[8] = {0, 0, 1, 0, 0, 1, 1, 1};
    const int32_t cellnodes[6] = {1, 2, 4, 4, 3, 1};
    const double xref[2] = {1.0 / 3.0, 1.0 / 3.0}, w[1] = {1.0};           /* QuadratureRule{Float64,Triangle2D}(0) */
    const double refvals[3] = {1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0};           /* phi_j(xref), [qp][dof][comp] */
    const double refderivs[6] = {-1, -1, 1, 0, 0, 1};                      /* d phi_j / d xref, [qp][dof][comp][dim] */
    const double rhs[3] = {0.0, 1.0, -1.0};                                /* f(x) = x1 - x2 (Example200:38) */
    const int32_t space[1] = {0};
    const int32_t block_rows[1] = {0}, block_cols[1] = {0};
    int64_t nnz = 0;
    CHECK(femb_create(0, &ctx));
    CHECK(femb_mesh_set(ctx, 2, 4, coords, 3, 2, cellnodes, NULL));
    CHECK(femb_space_set(ctx, 0, FEMB_FAMILY_H1, 1, 4, 3, 3, FEMB_HANDLER_NONE, NULL, NULL, NULL));
    CHECK(femb_quadrature_set(ctx, 1, xref, w));
    CHECK(femb_evaluator_set(ctx, 0, 0, FEMB_OP_GRADIENT, NULL, refderivs));
    CHECK(femb_evaluator_set(ctx, 1, 0, FEMB_OP_IDENTITY, refvals, NULL));
    CHECK(femb_matrix_layout(ctx, 1, space, 1, space));
    CHECK(femb_pattern_track(ctx, 1));
    CHECK(femb_pattern_build(ctx, 1, block_rows, block_cols, 0, NULL, &nnz));
    CHECK(femb_matrix_zero(ctx));
    CHECK(femb_vector_zero(ctx));
    CHECK(femb_assemble_poisson(ctx, 0, 1, 1.0, FEMB_RHS_AFFINE, rhs, 1.0e-15));
    CHECK(femb_pattern_compress(ctx, &nnz)); /* what flush! leaves after Example200's value filter */
    int64_t colptr[5];
    int64_t* rowval = malloc(sizeof(int64_t) * (size_t)nnz);
    double* nzval = malloc(sizeof(double) * (size_t)nnz);
    double b[4];
    CHECK(femb_pattern_get(ctx, colptr, rowval));
    CHECK(femb_matrix_get(ctx, nzval));
    CHECK(femb_vector_get(ctx, b));
    printf("nnz = %lld\n", (long long)nnz);
    for (int j = 0; j < 4; j++)
        for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; p++) printf("A[%lld,%d] = % .3f\n", (long long)rowval[p], j + 1, nzval[p]);
    printf("b = % .4f % .4f % .4f % .4f\n", b[0], b[1], b[2], b[3]);
    free(rowval);
    free(nzval);
    CHECK(femb_destroy(ctx));
    return 0;
}
