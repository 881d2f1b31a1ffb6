// Shared declarations of the numeric-phase translation units (femb_generic.cu, femb_p1.cu, femb_h1.cu, femb_ns.cu,
// femb_assemble.cu): device-side argument structs used across files and the host-side launchers.
#pragma once
#include <algorithm>

#include "femb_device.cuh"
#include "femb_internal.h"

#define FEMB_P1_MAXQP 8  // quadrature points the fused 2-D P1 kernel keeps in its parameter table

struct QuadDev {
    int nqp;
    double xref[FEMB_MAX_QP * 3];
    double w[FEMB_MAX_QP];
};

// scatter of one local entry through the cell->slot map (cell-parallel kernels)
__device__ __forceinline__ void scatter_add(double* nzval, uint8_t* touched, int32_t* errflag, int32_t slot, double v, int atomic) {
    if (slot < 0) {
        if (v != 0.0) atomicAdd(errflag, 1);
        return;
    }
    if (atomic) atomicAdd(nzval + slot, v);
    else nzval[slot] += v;
    if (touched) touched[slot] = 1;
}

// shared-memory accesses with an explicit 32-bit address (no generic-address arithmetic in the inner loops)
__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
// 0.0 when `skip` is set, else the shared-memory value (predicated load: no branch, no wavefront for skipped lanes)
__device__ __forceinline__ double lds64_unless(unsigned addr, unsigned skip) {
    double v;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@!p ld.shared.f64 %0, [%1];\n\t}" : "=d"(v) : "r"(addr), "r"(skip));
    return v;
}
__device__ __forceinline__ void lds128(unsigned addr, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void sts64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// femb_generic.cu
SpaceDev space_dev(const FembSpace& s);
QuadDev quad_dev(const FembEval& e, int dim);
int make_rhs(femb_ctx* ctx, int kind, const double* p, int ncomp, int nqp, RhsDev& r, double** dev_fvals);
const FembBlock* find_block(const femb_ctx* ctx, int brow, int bcol);
int check_eval(femb_ctx* ctx, int id);
int require_pattern(femb_ctx* ctx);
int launch_bilinear(femb_ctx* ctx, int er, int ec, int op_r, int op_c, int brow, int bcol, double factor, int flags, double drop_tol);
int launch_linear(femb_ctx* ctx, int ev, int brow, double factor, const RhsDev& rhs);
// femb_p1.cu
int launch_p1_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol);
// femb_h1.cu (both return FEMB_ERR_UNSUPPORTED, without an error message, when the path does not cover the space)
int launch_stiffness_dmma(femb_ctx* ctx, int eg, double mu, double drop_tol);
int launch_h1_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol);
int launch_h1_gather_vec(femb_ctx* ctx, int eg, double mu, int ei = -1, const RhsDev* rhs = nullptr, bool* rhs_done = nullptr);
// femb_p2.cu (FEMB_ERR_UNSUPPORTED, without an error message, when the closed form does not apply)
// host-to-host variant (femb_assemble_poisson_host): finished nzval / b ranges are downloaded chunk by chunk on stream_d2h
struct P2HostOut {
    double* nzval_out;
    double* b_out;
    int nchunks;
};
int launch_p2_gather(femb_ctx* ctx, int eg, int ei, double mu, const RhsDev& rhs, double drop_tol, const P2HostOut* host = nullptr);
int launch_p2_gather_vec(femb_ctx* ctx, int eg, double mu, int ei = -1, const RhsDev* rhs = nullptr, bool* rhs_done = nullptr);
// femb_hdiv.cu: HDIVRT0 on tetrahedra in closed form (FEMB_ERR_UNSUPPORTED, without an error message, when not covered)
int launch_rt0_mass_gather(femb_ctx* ctx, int er, int brow, int bcol, double factor);
int launch_bdm1_mass_gather(femb_ctx* ctx, int er, int brow, int bcol, double factor);
int launch_rt0_div(femb_ctx* ctx, int er, int ec, int brow, int bcol, double factor, int flags);
int launch_rt0_mixed(femb_ctx* ctx, int eid, int ediv, int eq, double fmass, double fdiv);
