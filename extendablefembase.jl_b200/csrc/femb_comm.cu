// Multi-GPU: interface-column exchange over NCCL (one process per GPU).  NCCL is bound at run time
// with dlopen so that single-GPU users do not need it.  (Implemented after the single-GPU path.)
#include "femb_internal.h"

extern "C" {
int32_t femb_comm_unique_id(char id_out[128]) { (void)id_out; return FEMB_ERR_UNSUPPORTED; }
int32_t femb_comm_init(femb_ctx* ctx, int32_t nranks, int32_t rank, const char id[128]) {
    (void)nranks; (void)rank; (void)id;
    return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "femb_comm_init: not built yet");
}
int32_t femb_comm_set_ghosts(femb_ctx* ctx, int32_t npeers, const int32_t* peers, const int64_t* ghost_ptr, const int64_t* ghost_local, const int64_t* ghost_remote) {
    (void)npeers; (void)peers; (void)ghost_ptr; (void)ghost_local; (void)ghost_remote;
    return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "femb_comm_set_ghosts: not built yet");
}
int32_t femb_exchange(femb_ctx* ctx) { return femb_fail(ctx, FEMB_ERR_UNSUPPORTED, "femb_exchange: not built yet"); }
int32_t femb_comm_destroy(femb_ctx* ctx) { (void)ctx; return FEMB_OK; }
}
