// Device math for the covariance kernels (sm_100a).
//
// The FP64 pipe (64 lanes/SM) is the scarce resource of both hot kernels, so the
// transcendental functions are written to minimise FP64-pipe instructions and to
// push everything else (exponent manipulation, table indexing, special-case tests)
// onto the integer/LSU/MUFU pipes, which issue in the shadow of the 2-cycle DFMA.
//
//   pn_exp2   2^(a*b)                  2 DFMA + DADD (reduction, product never formed)
//                                      + 2 DFMA + DMUL + DFMA (2048-entry table, degree-2 q)
//   pn_sqrt   sqrt(s)                  MUFU.RSQ64H + 1 DMUL + 4 DFMA (+ integer halving)
//   pn_rcp    1/b, b >= 1              MUFU.RCP64H + 4 DFMA
//   pn_log2   log2(b), b >= 1          1 DFMA + 5 poly + 3 DADD        (+ LDS.128, int)
//
// All are accurate to ~1 ulp (measured in tests/test_device_math.py against
// mpmath through the host build of this header), i.e. well inside the 1e-12
// relative parity contract with the reference's NumPy results.
//
// The header also compiles for the host (PNCOV_HOST_MATH) so the algorithms can
// be validated without a GPU; MUFU seeds are emulated with 22-bit approximations.
#pragma once

#include <math.h>
#include <stdint.h>

#include "pncov_tables.h"

#if defined(__CUDACC__)
#define PN_HD __host__ __device__ __forceinline__
#else
#define PN_HD inline
#include <string.h>
#endif

// ------------------------------------------------------------------------------------
// bit access
// ------------------------------------------------------------------------------------
PN_HD int pn_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    int64_t b;
    memcpy(&b, &x, 8);
    return (int)(b >> 32);
#endif
}
PN_HD int pn_lo(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    int64_t b;
    memcpy(&b, &x, 8);
    return (int)(b & 0xffffffff);
#endif
}
PN_HD double pn_make(int hi, int lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    int64_t b = ((int64_t)hi << 32) | (uint32_t)lo;
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}

PN_HD double pn_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// MUFU seeds (about 22 good bits).  Host build: emulate by truncating the exact value.
PN_HD double pn_rsqrt_seed(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    double y = 1.0 / sqrt(x);
    return pn_make(pn_hi(y), pn_lo(y) & (int)0xC0000000);  // keep 22 mantissa bits like MUFU.RSQ64H
#endif
}
PN_HD double pn_rcp_seed(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    double y = 1.0 / x;
    return pn_make(pn_hi(y), pn_lo(y) & (int)0xC0000000);
#endif
}

// ------------------------------------------------------------------------------------
// exp2
// ------------------------------------------------------------------------------------
#define PN_EXP2_N (1 << PNCOV_EXP2_BITS)

// Fast path valid for |u| < 1022 (result normal).  tab = the 2^(j/2048) table (shared memory
// on the device).  Returns false if the caller must take the slow path.
PN_HD bool pn_exp2_in_fast_range(double u) {
    // |u| < 1022  <=>  (hi & 0x7fffffff) < hi(1022.0) = 0x408FF000; also rejects inf/nan
    return (unsigned)(pn_hi(u) & 0x7fffffff) < 0x408FF000u;
}

// magic constant whose ulp is 2^-PNCOV_EXP2_BITS: adding it rounds to the table grid
#define PN_EXP2_MAGIC (0x1.8p+52 / (double)PN_EXP2_N)

PN_HD double pn_exp2_fast(double u, const double* __restrict__ tab) {
    const double magic = PN_EXP2_MAGIC;
    double t = u + magic;
    int n = pn_lo(t);                // round(u * 2^BITS) as a 32-bit two's complement integer
    double kf = t - magic;           // exact
    double r = u - kf;               // exact, |r| <= 2^-(BITS+1)
    double q = pn_fma(PNCOV_EXP2_E3, r, PNCOV_EXP2_E2);
    q = pn_fma(q, r, PNCOV_EXP2_E1);
    double T = tab[n & (PN_EXP2_N - 1)];
    double res = pn_fma(T * r, q, T);  // T * 2^r, in [1, 2)
    int e = n >> PNCOV_EXP2_BITS;      // floor(n / 2^BITS)
    return pn_make(pn_hi(res) + e * 1048576, pn_lo(res));  // scale by 2^e in the exponent field
}

// full-range version: the underflow band -1075.5 < u <= -1022 (subnormal results) is
// evaluated as 2^(u+128) * 2^-128, where u + 128 is exact and the product rounds once
PN_HD double pn_exp2(double u, const double* __restrict__ tab) {
    if (pn_exp2_in_fast_range(u)) return pn_exp2_fast(u, tab);
    if (u != u) return u;
    if (u > 0.0) return exp2(u);       // overflow side: never reached by the covariance kernels
    if (u < -1075.5) return 0.0;       // also -inf
    return pn_exp2_fast(u + 128.0, tab) * 0x1p-128;
}

// ------------------------------------------------------------------------------------
// sqrt (s >= 0)
// ------------------------------------------------------------------------------------
PN_HD double pn_sqrt(double s) {
    // fast path for normal s well inside the exponent range: hi in [0x01000000, 0x7fe00000)
    unsigned h = (unsigned)pn_hi(s);
    if (h - 0x01000000u < 0x7ee00000u) {
        double y = pn_rsqrt_seed(s);
        double g = s * y;
        double hh = pn_make(pn_hi(y) - 0x00100000, pn_lo(y));  // y / 2 without an FP64 instruction
        double e = pn_fma(-hh, g, 0.5);
        g = pn_fma(g, e, g);
        // the final correction is ~2^-44 relative, so the 22-bit h is accurate enough for it
        double d = pn_fma(-g, g, s);
        return pn_fma(d, hh, g);
    }
    return sqrt(s);  // zero, subnormal, huge, inf, nan
}

// ------------------------------------------------------------------------------------
// reciprocal (1 <= b < 2^1000)
// ------------------------------------------------------------------------------------
PN_HD double pn_rcp(double b) {
    unsigned h = (unsigned)pn_hi(b);
    if (h - 0x00200000u < 0x7e000000u) {
        double y = pn_rcp_seed(b);
        double e = pn_fma(-b, y, 1.0);
        y = pn_fma(y, e, y);
        e = pn_fma(-b, y, 1.0);
        return pn_fma(y, e, y);
    }
    return 1.0 / b;
}

// ------------------------------------------------------------------------------------
// log2 (b >= 1, finite) ; tab = { rc_j, -log2(rc_j) } pairs
// ------------------------------------------------------------------------------------
#define PN_LOG2_N (1 << PNCOV_LOG2_BITS)

PN_HD double pn_log2_ge1(double b, const double* __restrict__ tab) {
    int hi = pn_hi(b);
    if ((unsigned)(hi - 0x3ff00000) < 0x40000000u - 0x00100000u) {  // 1 <= b < 2^1023
        int e = (hi >> 20) - 1023;
        int j = (hi >> (20 - PNCOV_LOG2_BITS)) & (PN_LOG2_N - 1);
        double m = pn_make((hi & 0x000fffff) | 0x3ff00000, pn_lo(b));  // [1, 2)
        double rc = tab[2 * j];
        double lc = tab[2 * j + 1];
        double r = pn_fma(m, rc, -1.0);
        double q = pn_fma(PNCOV_LOG2_G5, r, PNCOV_LOG2_G4);
        q = pn_fma(q, r, PNCOV_LOG2_G3);
        q = pn_fma(q, r, PNCOV_LOG2_G2);
        q = pn_fma(q, r, PNCOV_LOG2_G1);
        // e as a double without a conversion instruction on the FP64 pipe
        double ef = pn_make(0x43300000, (int)(0x80000000u ^ (unsigned)e)) - 4503601774854144.0;  // 2^52 + 2^31
        return pn_fma(q, r, ef + lc);
    }
    return log2(b);
}
