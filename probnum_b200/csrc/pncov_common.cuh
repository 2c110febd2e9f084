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
    // matrix-free product only: ExpQuad through exp(-|x|^2) exp(-|y|^2) exp(2 x.y); the kernel sees
    // u = -2 x.y straight from the DMMA accumulators and the two outer factors ride on v and y
    PN_EPI_EXPQUAD_FACT = 8,
    // dense Gram only: a short program over several isotropic leaves that share S (covariance arithmetic)
    PN_EPI_FUSED = 9,
    // dense symmetric Gram only: ProductMatern whose dimensions share the degree p = 0 .. 3
    // (_product_matern.py:103-122); the per-dimension parameters travel as PnPmProgram
    PN_EPI_PM0 = 10,
    PN_EPI_PM1 = 11,
    PN_EPI_PM2 = 12,
    PN_EPI_PM3 = 13,
    // PN_EPI_FUSED compiled for programs of at most 2 / 3 leaf evaluations (the unrolled slots of the full form are
    // 400 KB of code: instruction fetch, not arithmetic, then limits short programs)
    PN_EPI_FUSED2 = 14,
    PN_EPI_FUSED3 = 15,
};

struct PnEpiParams {
    int epi;
    int p;
    double c_u;      // EXPQUAD: -log2(e) * s^2 (isotropic) or -log2(e) (ARD); MATERN: -log2(e)
    double c_s;      // MATERN: s (isotropic) or 1 (ARD), multiplies sqrt(S); RATQUAD: 1/(2 alpha l^2)
    double c_su;     // MATERN0: c_s * c_u
    double alpha;    // RATQUADG
    double coef[PN_MAX_P + 1];  // MATERN: coefficients of the polynomial in r0 = sqrt(S): c_i * c_s^i, degree i at index i
    // polynomial coefficients of pn_exp2 / pn_log2, carried in the kernel-parameter constant
    // bank so that every use is a c[0x0][..] operand instead of two UMOVs (see profiles/r01)
    double e1, e2, e3;
    double f1, f2, f3, f4;  // degree-3 variant for the 256-entry replicated table
    double g1, g2, g3, g4, g5;
    // validity window of the fast epilogue in terms of the hi word of S (see PnFlags)
    unsigned s_lo_word, s_hi_word;
    // dense Gram only: the common covariance arithmetic `[o *] k [+ WhiteNoise]` folded into the compile-time
    // specialised kernels (pncov_set_program recognises it): K = K * post_scale (bit 0), then K += post_sigma_sq where
    // the two points are equal in every coordinate (bit 1) -- the operations, in the order, of the operand path
    int post_flags, post_pad_;
    double post_scale, post_sigma_sq;
};

// ---------------------------------------------------------------------------------------
// Covariance arithmetic in one launch (Scaled / Sum / Product nodes of
// /root/reference/src/probnum/randprocs/covfuncs/_arithmetic_fallbacks.py:55-56, :101-104, :163-166).
// All leaves are isotropic kernels on the same points, so they share S_ij; the program below runs per
// register tile over two values per entry, `acc` (the result so far) and `val` (the operand being
// built), with exactly the arithmetic the operand-by-operand path applies to whole matrices
// (combine_kernel in pncov_extra.cu): results are bit-identical to that path.
// ---------------------------------------------------------------------------------------
#define PN_FUSED_MAX_LEAVES 6   // leaf evaluations per program
#define PN_FUSED_MAX_STEPS 12
enum PnFusedOp : int {
    PN_F_LOAD = 0,      // val = leaf(S)
    PN_F_MULLEAF = 1,   // val = val * leaf(S)
    PN_F_SCALEVAL = 2,  // val = val * a
    PN_F_SET = 3,       // acc = val
    PN_F_LIN = 4,       // acc = fma(b, val, a * acc)
    PN_F_MUL = 5,       // acc = acc * val
    PN_F_SCALEACC = 6,  // acc = acc * a
};
struct PnFusedStep {
    int op, pad_;
    double a, b;
};
// The kernels see the program regrouped by leaf evaluation ("slot") in a canonical form that runs without copies
// and almost without branches: evaluate the leaf t, then
//     val = (mul ? val : 1) * t * s_val          (LOAD / MULLEAF, then an optional SCALEVAL)
//     acc = kind == MUL ? acc * val : fma(b, val, a * acc)    (SET is a = 0, b = 1; "no update" is a = 1, b = 0)
//     acc = acc * s_acc                          (an optional SCALEACC)
// Multiplications by 1.0 and additions of 0.0 are exact, so the bits are those of the step list.  Slots are
// unrolled at compile time: every epilogue constant is an immediate constant-bank operand exactly as in the
// single-kernel launches.
struct PnFusedSlot {
    PnEpiParams epi;
    double sigma_sq;  // WhiteNoise: sigma^2 where x0_i == x1_j (all coordinates), else 0
    int whitenoise;
    int mul;          // 0: val = leaf(S) (LOAD), 1: val = val * leaf(S) (MULLEAF)
    int kind;         // 1: acc = fma(b, val, a * acc); 2: acc = acc * val
    int n_post;       // host bookkeeping while the step list is regrouped
    double s_val, a, b, s_acc;
    PnFusedStep post[4];  // the slot's steps as given (recognition of `[o *] k [+ WhiteNoise]`)
};
struct PnFusedProgram {
    int n_slots;
    int has_wn, need_log, pad_;
    PnFusedSlot slots[PN_FUSED_MAX_LEAVES];
};
struct PnNoProgram {};
// ProductMatern, d <= 8: per dimension {scale, p, c_0 .. c_8, pad} (PN_PM_LD doubles, pncov_internal.h)
struct PnPmProgram {
    double pm[8][12];
};

inline void pn_fill_math_constants(PnEpiParams& P) {
    P.e1 = PNCOV_EXP2_E1; P.e2 = PNCOV_EXP2_E2; P.e3 = PNCOV_EXP2_E3;
    P.f1 = PNCOV_EXP2S_F1; P.f2 = PNCOV_EXP2S_F2; P.f3 = PNCOV_EXP2S_F3; P.f4 = PNCOV_EXP2S_F4;
    P.g1 = PNCOV_LOG2_G1; P.g2 = PNCOV_LOG2_G2; P.g3 = PNCOV_LOG2_G3; P.g4 = PNCOV_LOG2_G4;
    P.g5 = PNCOV_LOG2_G5;
}

// constant tables in global memory (one copy per translation unit, 6 KB); kernels stage
// them into shared memory at start-up
static __device__ __align__(128) const double pn_exp2_table_g[PN_EXP2_N] = PNCOV_EXP2_TABLE_INIT;
static __device__ __align__(128) const double pn_log2_table_g[2 * PN_LOG2_N] = PNCOV_LOG2_TABLE_INIT;

// 256-entry table replicated 16 times: entry j of copy c lives at [j * 16 + c], so the 16
// lanes of a half-warp (the unit a 64-bit shared load is served in) read 16 different
// 8-byte banks whatever their indices are: lookups are bank-conflict free (2 wavefronts per
// warp instead of ~6.6 for random indices into a single copy).  Costs one more DFMA (degree-3
// polynomial); used where shared-memory wavefronts, not the FP64 pipe, are the limiter.
static __device__ const double pn_exp2s_table_g[256] = PNCOV_EXP2S_TABLE_INIT;
#define PN_EXP2S_REP 16
#define PN_EXP2S_DOUBLES (256 * PN_EXP2S_REP)

__device__ __forceinline__ void pn_stage_tables_rep(double* tabE, double* tabL, int tid, int nthreads,
                                                    bool need_log) {
    for (int i = tid; i < PN_EXP2S_DOUBLES; i += nthreads) tabE[i] = pn_exp2s_table_g[i >> 4];
    if (need_log)
        for (int i = tid; i < 2 * PN_LOG2_N; i += nthreads) tabL[i] = pn_log2_table_g[i];
}

__device__ __forceinline__ void pn_stage_tables(double* tabE, double* tabL, int tid, int nthreads) {
    for (int i = tid; i < PN_EXP2_N; i += nthreads) tabE[i] = pn_exp2_table_g[i];
    for (int i = tid; i < 2 * PN_LOG2_N; i += nthreads) tabL[i] = pn_log2_table_g[i];
}

// ---------------------------------------------------------------------------------------
// Robust FP64 epilogue (any S >= 0 including 0, subnormal, huge, inf, nan): the slow path.
// ---------------------------------------------------------------------------------------
static __device__ __noinline__ double pn_epilogue(double S, const PnEpiParams& P,
                                           const double* __restrict__ tabE,
                                           const double* __restrict__ tabL) {
    switch (P.epi) {
        case PN_EPI_EXPQUAD:
            return pn_exp2(S * P.c_u, tabE);
        case PN_EPI_MATERN0:
            return pn_exp2(pn_sqrt(S) * P.c_su, tabE);
        case PN_EPI_MATERN1:
        case PN_EPI_MATERN2:
        case PN_EPI_MATERNP: {
            // coef[] holds the Horner coefficients in the variable r0 = sqrt(S): c_i * c_s^i
            double r0 = pn_sqrt(S);
            double e = pn_exp2(r0 * P.c_su, tabE);
            double poly = P.coef[P.p];
            for (int i = P.p - 1; i >= 0; --i) poly = pn_fma(poly, r0, P.coef[i]);
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

// ---------------------------------------------------------------------------------------
// Fast epilogue.  No per-entry branches: each call folds the hi word of S into a running
// min and max (two integer instructions) and the caller tests them ONCE per register tile
// against the window [s_lo_word, s_hi_word) computed on the host for this kernel; a
// flagged tile is redone with pn_epilogue.  The window excludes S == 0 / subnormal under a
// square root (the diagonal of a symmetric Gram), results that would underflow to
// subnormals, overflow of 1 + c S, and non-finite input.
// ---------------------------------------------------------------------------------------
struct PnFlags {
    unsigned lo, hi;
    __device__ __forceinline__ void reset() {
        lo = 0xffffffffu;
        hi = 0u;
    }
    __device__ __forceinline__ void see(double S) {
        const unsigned w = (unsigned)pn_hi(S);
        lo = min(lo, w);
        hi = max(hi, w);
    }
    __device__ __forceinline__ bool flagged(const PnEpiParams& P) const {
        return lo < P.s_lo_word || hi >= P.s_hi_word;
    }
};

// 2^(a*b) for |a*b| < 1021.  The product is never rounded: t = fma(a, b, magic) rounds it
// to the table grid, r = fma(a, b, -(t - magic)) is the exact remainder (one rounding).
__device__ __forceinline__ double pn_exp2_prod_core(double a, double b, const PnEpiParams& P,
                                                    const double* __restrict__ tab) {
    const double magic = PN_EXP2_MAGIC;
    double t = pn_fma(a, b, magic);
    int n = pn_lo(t);
    double r = pn_fma(a, b, magic - t);
    double q = pn_fma(P.e3, r, P.e2);
    q = pn_fma(q, r, P.e1);
    double T = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tab) +
                                                ((n << 3) & ((PN_EXP2_N - 1) << 3)));
    double res = pn_fma(T * r, q, T);
    // (n >> BITS) << 20 == n << (20 - BITS) with the low 20 bits cleared
    return pn_make(pn_hi(res) + ((n << (20 - PNCOV_EXP2_BITS)) & 0xfff00000), pn_lo(res));
}
// same with the replicated 256-entry table (TAB = 1): one more DFMA, no bank conflicts
__device__ __forceinline__ double pn_exp2_prod_core_rep(double a, double b, const PnEpiParams& P,
                                                        const double* __restrict__ tab_lane) {
    const double magic = 0x1.8p+44;  // ulp = 2^-8
    double t = pn_fma(a, b, magic);
    int n = pn_lo(t);
    double r = pn_fma(a, b, magic - t);
    double q = pn_fma(P.f4, r, P.f3);
    q = pn_fma(q, r, P.f2);
    q = pn_fma(q, r, P.f1);
    // tab_lane = table + (lane & 15): entry j of this lane's copy is tab_lane[j * 16]
    double T = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tab_lane) + ((n & 255) << 7));
    double res = pn_fma(T * r, q, T);
    return pn_make(pn_hi(res) + ((n << 12) & 0xfff00000), pn_lo(res));
}

__device__ __forceinline__ double pn_sqrt_core(double s) {
    double y = pn_rsqrt_seed(s);
    double g = s * y;
    // h = y / 2 on the integer pipe (exponent - 1; y is a normal number in the fast window)
    double h = pn_make(pn_hi(y) - 0x00100000, pn_lo(y));
    double e = pn_fma(-h, g, 0.5);
    g = pn_fma(g, e, g);
    double d = pn_fma(-g, g, s);
    return pn_fma(d, h, g);
}
__device__ __forceinline__ double pn_rcp_core(double b) {
    double y = pn_rcp_seed(b);
    double e = pn_fma(-b, y, 1.0);
    y = pn_fma(y, e, y);
    e = pn_fma(-b, y, 1.0);
    return pn_fma(y, e, y);
}
__device__ __forceinline__ double pn_log2_core(double b, const PnEpiParams& P,
                                               const double* __restrict__ tab) {
    int hi = pn_hi(b);
    int e = (hi >> 20) - 1023;
    double m = pn_make((hi & 0x000fffff) | 0x3ff00000, pn_lo(b));
    const double2 t = *reinterpret_cast<const double2*>(
        reinterpret_cast<const char*>(tab) + (((hi >> (20 - PNCOV_LOG2_BITS)) & (PN_LOG2_N - 1)) << 4));
    double r = pn_fma(m, t.x, -1.0);
    double q = pn_fma(P.g5, r, P.g4);
    q = pn_fma(q, r, P.g3);
    q = pn_fma(q, r, P.g2);
    q = pn_fma(q, r, P.g1);
    double ef = pn_make(0x43300000, (int)(0x80000000u ^ (unsigned)e)) - 4503601774854144.0;
    return pn_fma(q, r, ef + t.y);
}

// TAB = 0: tabE is the 2048-entry table; TAB = 1: tabE is this lane's view of the replicated
// 256-entry table (base + (lane & 15))
template <int EPI, int TAB = 0>
struct PnEpilogue {
    __device__ __forceinline__ static double exp2_prod(double a, double b, const PnEpiParams& P,
                                                       const double* __restrict__ tabE) {
        if constexpr (TAB == 0) return pn_exp2_prod_core(a, b, P, tabE);
        else return pn_exp2_prod_core_rep(a, b, P, tabE);
    }
    __device__ __forceinline__ static double fast(double S, const PnEpiParams& P,
                                                  const double* __restrict__ tabE,
                                                  const double* __restrict__ tabL, PnFlags& f) {
        f.see(S);
        if constexpr (EPI == PN_EPI_EXPQUAD) {
            return exp2_prod(S, P.c_u, P, tabE);
        } else if constexpr (EPI >= PN_EPI_MATERN0 && EPI <= PN_EPI_MATERNP) {
            double r0 = pn_sqrt_core(S);
            double e = exp2_prod(r0, P.c_su, P, tabE);
            // The scale c_s of r = c_s r0 is folded into the coefficients on the host
            // (coef[i] = c_i c_s^i, coef[0] = 1): one FP64 instruction less per entry.
            if constexpr (EPI == PN_EPI_MATERN0) {
                return e;
            } else if constexpr (EPI == PN_EPI_MATERN1) {
                return pn_fma(r0, P.coef[1], 1.0) * e;
            } else if constexpr (EPI == PN_EPI_MATERN2) {
                return pn_fma(pn_fma(r0, P.coef[2], P.coef[1]), r0, 1.0) * e;
            } else {
                double poly = P.coef[P.p];
                for (int i = P.p - 1; i >= 1; --i) poly = pn_fma(poly, r0, P.coef[i]);
                return pn_fma(poly, r0, 1.0) * e;
            }
        } else {
            double b = pn_fma(S, P.c_s, 1.0);
            if constexpr (EPI == PN_EPI_RATQUAD1) {
                return pn_rcp_core(b);
            } else if constexpr (EPI == PN_EPI_RATQUAD2) {
                double y = pn_rcp_core(b);
                return y * y;
            } else {
                return exp2_prod(pn_log2_core(b, P, tabL), -P.alpha, P, tabE);
            }
        }
    }
};

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

// Tables through the TMA engine: ONE thread issues the bulk copies, everybody waits on `bar`
// (phase 0) before the first epilogue.  Unlike pn_stage_tables no warp stalls on global loads
// in the CTA prologue (measured: 8 % of the stall samples of the symmetric Gram kernel).
__device__ __forceinline__ void pn_stage_tables_tma(double* tabE, double* tabL, uint64_t* bar, bool need_log) {
    const uint32_t bytes = 8u * PN_EXP2_N + (need_log ? 16u * PN_LOG2_N : 0u);
    pn_mbar_expect_tx(bar, bytes);
    pn_bulk_g2s(tabE, pn_exp2_table_g, 8u * PN_EXP2_N, bar);
    if (need_log) pn_bulk_g2s(tabL, pn_log2_table_g, 16u * PN_LOG2_N, bar);
}

// ---------------------------------------------------------------------------------------
// Order-independent accumulation for the symmetric-halving products.  Those kernels add the partial
// sums of many CTAs into one vector; floating-point atomics would make the result depend on the order
// in which the CTAs retire.  Instead every partial sum s is split exactly into two integers on a
// per-target power-of-two grid, s = hi 2^(E-52) + lo 2^(E-100) + (|.| <= 2^(E-101)), and the integers
// are added with 64-bit integer atomics: integer addition is associative, so the vector comes out bit
// for bit the same on every run -- and for every split of the row blocks over GPUs, whose integer
// planes are summed by an all-reduce.  2^E bounds the sum of the magnitudes of everything that can
// reach the target (n1 max|v| / f_t for the factorised ExpQuad form, n1 max|v| otherwise), so neither
// plane overflows; the resolution 2^(E-101) is far below the rounding of the partial sums themselves.
// Non-finite partial sums (NaN / Inf in v or in the points) are added as doubles to a third plane:
// their sums do not depend on the order either (IEEE: Inf + Inf = Inf, Inf - Inf = NaN, x + NaN = NaN).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double pn_pow2(int e) { return pn_make((e + 1023) << 20, 0); }
// E for a target with factor f (1 when the product is not factorised); bound0 = n1 * max|v|
__device__ __forceinline__ int pn_bin_exponent(double bound0, double f) {
    const int e = ((pn_hi(bound0) >> 20) & 0x7ff) - ((pn_hi(f) >> 20) & 0x7ff) + 2;
    return max(-920, min(1023, e));  // 2^(E-100) stays a normal number, 2^E a finite one
}
__device__ __forceinline__ void pn_bin_add(long long* hi_plane, long long* lo_plane, double* nf_plane, int64_t idx, double s,
                                           int E) {
    if (!(fabs(s) <= pn_pow2(E))) {  // NaN, Inf (or a bound violated by non-finite input)
        atomicAdd(nf_plane + idx, s);
        return;
    }
    const long long hi = __double2ll_rn(s * pn_pow2(52 - E));
    const double rem = fma(-(double)hi, pn_pow2(E - 52), s);  // exact: |hi| <= 2^52
    const long long lo = __double2ll_rn(rem * pn_pow2(100 - E));
    if (hi) atomicAdd(reinterpret_cast<unsigned long long*>(hi_plane + idx), (unsigned long long)hi);
    if (lo) atomicAdd(reinterpret_cast<unsigned long long*>(lo_plane + idx), (unsigned long long)lo);
}
__device__ __forceinline__ double pn_bin_value(long long hi, long long lo, double nf, int E) {
    const double y = fma((double)lo, pn_pow2(E - 100), (double)hi * pn_pow2(E - 52));
    return nf != 0.0 ? nf + y : y;
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
