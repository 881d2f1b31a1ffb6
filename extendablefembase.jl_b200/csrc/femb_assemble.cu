// Example200.assemble! entry point: dispatch between the fused P1 gather, the general H1 gather, the FP64 tensor-core
// kernel and the generic quadrature kernels (see femb_p1.cu, femb_h1.cu, femb_generic.cu).
#include "femb_kernels.cuh"

extern "C" int32_t femb_assemble_poisson(femb_ctx* ctx, int32_t eg, int32_t ei, double mu, int32_t rhs_kind, const double* data, double drop_tol) {
    FEMB_CHECK_CTX(ctx);
    FEMB_TRY(check_eval(ctx, eg));
    if (rhs_kind != FEMB_RHS_NONE) FEMB_TRY(check_eval(ctx, ei));
    FEMB_TRY(require_pattern(ctx));
    const FembEval& EG = ctx->evals[eg];
    if (EG.op != FEMB_OP_GRADIENT) return femb_fail(ctx, FEMB_ERR_ARG, "eval_grad must be a Gradient evaluator");
    if (rhs_kind != FEMB_RHS_NONE && (ctx->evals[ei].op != FEMB_OP_IDENTITY || ctx->evals[ei].space != EG.space || ctx->evals[ei].nqp != EG.nqp))
        return femb_fail(ctx, FEMB_ERR_ARG, "eval_id must be the Identity evaluator of the same space and rule");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RhsDev rhs;
    double* dev_f = nullptr;
    FEMB_TRY(make_rhs(ctx, rhs_kind, data, ctx->spaces[EG.space].ncomp, EG.nqp, rhs, &dev_f));
    const bool fused = ctx->fused_exchange && ctx->comm && !ctx->peers.empty() && ctx->stream_x;
    if (fused) {
        // the exchange sends the whole ghost slot / entry: accumulating over several assemblies would re-send earlier sums
        if (ctx->dirty_A || ctx->dirty_b) {
            cudaFree(dev_f);
            return femb_fail(ctx, FEMB_ERR_ARG, "fused exchange: call femb_matrix_zero and femb_vector_zero before each femb_assemble_poisson");
        }
        if (ctx->exchange_nnz != ctx->nnz) {
            cudaFree(dev_f);
            return femb_fail(ctx, FEMB_ERR_ARG, "fused exchange: the exchange plan was built for another pattern");
        }
        if (rhs_kind == FEMB_RHS_NONE) {  // no rhs is assembled: the vector part of the plan must not carry stale values
            int zs = femb_flush_zero(ctx, false, true);
            if (zs != FEMB_OK) { cudaFree(dev_f); return zs; }
        }
    }
    ctx->exchange_done = false;
    ctx->dirty_A = true;
    if (rhs_kind != FEMB_RHS_NONE) ctx->dirty_b = true;
    int st = femb_timer_begin(ctx);
    if (st == FEMB_OK) {
        if (ctx->scatter == FEMB_SCATTER_GATHER && ctx->gmap && ctx->gmap_space == EG.space && (!ctx->touched || ctx->max_collen <= 64) &&
            rhs_kind != FEMB_RHS_EX210 && (ctx->dim == 3 || EG.nqp <= FEMB_P1_MAXQP)) {
            st = launch_p1_gather(ctx, eg, ei, mu, rhs, drop_tol);
        } else if (ctx->scatter == FEMB_SCATTER_GATHER && rhs_kind != FEMB_RHS_EX210 && (st = launch_h1_gather(ctx, eg, ei, mu, rhs, drop_tol)) != FEMB_ERR_UNSUPPORTED) {
            // owner-computes gather for general scalar H1 elements (stiffness + rhs in one pass, write-once)
        } else {
            st = (ctx->scatter == FEMB_SCATTER_COLORED) ? FEMB_ERR_UNSUPPORTED : launch_stiffness_dmma(ctx, eg, mu, drop_tol);
            if (st == FEMB_ERR_UNSUPPORTED)  // elements with per-cell basis subsets, vector-valued spaces, ...: generic quadrature kernel
                st = launch_bilinear(ctx, eg, eg, FEMB_OP_GRADIENT, FEMB_OP_GRADIENT, 0, 0, mu, FEMB_BIL_SYMMETRIC_UPPER, drop_tol);
            if (st == FEMB_OK && rhs_kind != FEMB_RHS_NONE) st = launch_linear(ctx, ei, 0, 1.0, rhs);
        }
    }
    if (st == FEMB_OK && fused && !ctx->exchange_done) {
        // dispatch paths without an overlapped exchange (general H1 gather, tensor-core, generic kernels): same result, sequential
        st = femb_flush_zero(ctx, true, true);
        if (st == FEMB_OK) st = femb_exchange_on(ctx, ctx->stream);
    }
    if (st == FEMB_OK) st = femb_timer_end(ctx);
    cudaFree(dev_f);
    return st;
}
