// Shared device-side definitions: kernel parameters, epilogues, TMA/mbarrier helpers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pncov_math.cuh"

#define PN_MAX_P 8

// epilogue selector (uniform across a launch; switched at run time so that the
// (kind x D x ARD x columns) template space stays small)
enum PnEpi : int {
    PN_EPI_EXPQUAD = 0,  // exp(-S)
    PN_EPI_MATERN0 = 1,  // exp(-r)
    PN_EPI_MATERN1 = 2,  // (1 + r) exp(-r)
    PN_EPI_MATERN2 = 3,  // (1 + r + r^2/3) exp(-r)
    PN_EPI_MATERNP = 4,  // Horner of degree p
    PN_EPI_RATQUAD1 = 5, // 1 / b
    PN_EPI_RATQUAD2 = 6, // 1 / b^2
    PN_EPI_RATQUADG = 7, // b^-alpha
};

struct PnEpiParams {
    int epi;
    int p;
    double c_u;      // EXPQUAD: -log2(e) * s^2 (isotropic) or -log2(e) (ARD); MATERN: -log2(e)
    double c_s;      // MATERN: s (isotropic) or 1 (ARD), multiplies sqrt(S); RATQUAD: 1/(2 alpha l^2)
    double alpha;    // RATQUADG
    double coef[PN_MAX_P + 1];  // MATERNP polynomial coefficients, degree i at index i
};

// constant tables in global memory (one copy per translation unit, 6 KB); kernels stage
// them into shared memory at start-up
static __device__ const double pn_exp2_table_g[PN_EXP2_N] = PNCOV_EXP2_TABLE_INIT;
static __device__ const double pn_log2_table_g[2 * PN_LOG2_N] = PNCOV_LOG2_TABLE_INIT;

__device__ __forceinline__ void pn_stage_tables(double* tabE, double* tabL, int tid, int nthreads) {
    for (int i = tid; i < PN_EXP2_N; i += nthreads) tabE[i] = pn_exp2_table_g[i];
    for (int i = tid; i < 2 * PN_LOG2_N; i += nthreads) tabL[i] = pn_log2_table_g[i];
}

// ---------------------------------------------------------------------------------------
// FP64 epilogue: S = sum of squared (scaled) differences  ->  k
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double pn_epilogue(double S, const PnEpiParams& P,
                                              const double* __restrict__ tabE,
                                              const double* __restrict__ tabL) {
    switch (P.epi) {
        case PN_EPI_EXPQUAD:
            return pn_exp2(S * P.c_u, tabE);
        case PN_EPI_MATERN0: {
            double r = pn_sqrt(S) * P.c_s;
            return pn_exp2(r * P.c_u, tabE);
        }
        case PN_EPI_MATERN1: {
            double r = pn_sqrt(S) * P.c_s;
            double e = pn_exp2(r * P.c_u, tabE);
            return pn_fma(r, e, e);
        }
        case PN_EPI_MATERN2: {
            double r = pn_sqrt(S) * P.c_s;
            double e = pn_exp2(r * P.c_u, tabE);
            double poly = pn_fma(pn_fma(r, 1.0 / 3.0, 1.0), r, 1.0);
            return poly * e;
        }
        case PN_EPI_MATERNP: {
            double r = pn_sqrt(S) * P.c_s;
            double e = pn_exp2(r * P.c_u, tabE);
            double poly = P.coef[P.p];
            for (int i = P.p - 1; i >= 0; --i) poly = pn_fma(poly, r, P.coef[i]);
            return poly * e;
        }
        case PN_EPI_RATQUAD1:
            return pn_rcp(pn_fma(S, P.c_s, 1.0));
        case PN_EPI_RATQUAD2: {
            double y = pn_rcp(pn_fma(S, P.c_s, 1.0));
            return y * y;
        }
        default: {  // PN_EPI_RATQUADG
            double L = pn_log2_ge1(pn_fma(S, P.c_s, 1.0), tabL);
            return pn_exp2(-P.alpha * L, tabE);
        }
    }
}

// float32 inputs.  Distances are accumulated in float (as the reference does for
// float32 arrays); ExpQuad stays in float, the other families continue in double
// because the reference's float64 hyper-parameter scalars promote there (NumPy 2).
__device__ __forceinline__ float pn_epilogue_f32_expquad(float S, const PnEpiParams& P) {
    // S already carries the float32-rounded scale factors; c_u = -log2(e) * (s^2 or 1)
    return exp2f(S * (float)P.c_u);
}

// ---------------------------------------------------------------------------------------
// TMA (1-D bulk copy) + mbarrier
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pn_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void pn_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pn_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pn_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void pn_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pn_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pn_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PN_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PN_DONE_%=;\n"
        "bra PN_WAIT_%=;\n"
        "PN_DONE_%=:\n"
        "}\n" ::"r"(pn_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (UBLKCP), completion signalled on `bar`.  16-byte aligned
// addresses, size a multiple of 16 bytes.
__device__ __forceinline__ void pn_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(pn_smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(pn_smem_u32(bar))
        : "memory");
}

// streaming (evict-first) stores for write-once outputs
__device__ __forceinline__ void pn_store_cs2(double* p, double a, double b) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
}
__device__ __forceinline__ void pn_store_cs2(float* p, float a, float b) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(a, b));
}
__device__ __forceinline__ void pn_store_cs1(double* p, double a) { __stcs(p, a); }
__device__ __forceinline__ void pn_store_cs1(float* p, float a) { __stcs(p, a); }
