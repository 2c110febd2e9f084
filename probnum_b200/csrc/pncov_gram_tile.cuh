// Shared pieces of the dense Gram kernels (pncov_gram.cu: strip / fast / chunk kernels and the dispatcher;
// pncov_gram_sym.cu: the symmetric mirrored-store kernel with its program and ProductMatern variants).  Two
// translation units because the template space compiles for minutes otherwise.
#pragma once

#include <type_traits>

#include "pncov_internal.h"

namespace {


constexpr int G_THREADS = 256;
constexpr int G_TN = 2 * G_THREADS;  // columns per strip
constexpr int G_TM = 64;             // rows per pipelined x0 tile
constexpr int G_R = 4;               // rows per register tile
constexpr int G_DMAX = 16;           // strip kernel: all coordinates of a strip live in smem
constexpr int G_CHUNK_ROWS = 8;      // chunk kernel (d > G_DMAX): rows per CTA tile
constexpr int G_HDR_BYTES = 128 + 8 * (PN_EXP2_N + 2 * PN_LOG2_N);  // barriers + tables

template <typename T>
struct Pair;
template <>
struct Pair<double> {
    using type = double2;
};
template <>
struct Pair<float> {
    using type = float2;
};

// ---- epilogue over a register tile ------------------------------------------------------
template <int EPI, int N, typename T, typename TOut>
__device__ __forceinline__ void epi_tile(const T (&S)[N], TOut (&K)[N], const PnEpiParams& P,
                                         const double* tabE, const double* tabL) {
    if constexpr (std::is_same<TOut, float>::value) {
#pragma unroll
        for (int i = 0; i < N; ++i) K[i] = exp2f(S[i] * (float)P.c_u);  // float32 ExpQuad stays in float
    } else {
        PnFlags f;
        f.reset();
#pragma unroll
        for (int i = 0; i < N; ++i) K[i] = PnEpilogue<EPI>::fast((double)S[i], P, tabE, tabL, f);
        if (f.flagged(P)) {  // rare: diagonal tiles (sqrt of 0), subnormal results
            // only the entries outside the window are redone: which path an entry takes depends on its own S alone,
            // so K_ij and K_ji (equal S) get equal bits wherever they are evaluated
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const unsigned w = (unsigned)pn_hi((double)S[i]);
                if (w < P.s_lo_word || w >= P.s_hi_word) K[i] = pn_epilogue((double)S[i], P, tabE, tabL);
            }
        }
    }
}

template <int N, typename T, typename TOut>
__device__ __forceinline__ void epi_dispatch(const T (&S)[N], TOut (&K)[N], const PnEpiParams& P,
                                             const double* tabE, const double* tabL) {
    switch (P.epi) {  // uniform across the grid
        case PN_EPI_EXPQUAD: epi_tile<PN_EPI_EXPQUAD>(S, K, P, tabE, tabL); break;
        case PN_EPI_MATERN0: epi_tile<PN_EPI_MATERN0>(S, K, P, tabE, tabL); break;
        case PN_EPI_MATERN1: epi_tile<PN_EPI_MATERN1>(S, K, P, tabE, tabL); break;
        case PN_EPI_MATERN2: epi_tile<PN_EPI_MATERN2>(S, K, P, tabE, tabL); break;
        case PN_EPI_MATERNP: epi_tile<PN_EPI_MATERNP>(S, K, P, tabE, tabL); break;
        case PN_EPI_RATQUAD1: epi_tile<PN_EPI_RATQUAD1>(S, K, P, tabE, tabL); break;
        case PN_EPI_RATQUAD2: epi_tile<PN_EPI_RATQUAD2>(S, K, P, tabE, tabL); break;
        default: epi_tile<PN_EPI_RATQUADG>(S, K, P, tabE, tabL); break;
    }
}

// covariance arithmetic: the program of PnFusedProgram over one register tile (see pncov_common.cuh).
// `eq` has bit i set where the two points of entry i are equal in every coordinate (WhiteNoise leaves).
template <int N, int NSLOTS = PN_FUSED_MAX_LEAVES>
__device__ __forceinline__ void fused_tile(const double (&S)[N], double (&K)[N], const PnFusedProgram& F, unsigned eq,
                                           const double* tabE, const double* tabL) {
    double val[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        K[i] = 0.0;
        val[i] = 1.0;
    }
    // Canonical slots (PnFusedSlot), unrolled at compile time.  Only the leaf's kernel function is a run-time switch;
    // LOAD vs MULLEAF is a select of the multiplicand and the optional scalings are multiplications by 1.0, because
    // what limited the first forms of this kernel (profiles/README.md, round 2) was neither the FP64 pipe nor HBM
    // but register copies (`val = t`, `acc = val`) and the merges after run-time branches over whole register tiles:
    // a third of the executed instructions were moves.
#pragma unroll
    for (int l = 0; l < NSLOTS; ++l) {
        if (l < F.n_slots) {  // uniform across the grid
            const PnFusedSlot& L = F.slots[l];
            if (L.whitenoise == 2) {
                // a WhiteNoise summand, acc = fma(b, sigma^2 [x0 == x1], acc): acc + 0 is acc, so only entries whose
                // points are equal change (the diagonal, duplicated points)
                if (eq) {
#pragma unroll
                    for (int i = 0; i < N; ++i)
                        if ((eq >> i) & 1u) K[i] = pn_fma(L.b, L.sigma_sq, K[i]);
                }
                continue;
            }
            double t[N];
            if (L.whitenoise) {
#pragma unroll
                for (int i = 0; i < N; ++i) t[i] = ((eq >> i) & 1u) ? L.sigma_sq : 0.0;
            } else {
                epi_dispatch(S, t, L.epi, tabE, tabL);
            }
            const bool mul = L.mul != 0;
#pragma unroll
            for (int i = 0; i < N; ++i) val[i] = ((mul ? val[i] : 1.0) * t[i]) * L.s_val;
            if (L.kind == 2) {
#pragma unroll
                for (int i = 0; i < N; ++i) K[i] *= val[i];
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) K[i] = pn_fma(L.b, val[i], L.a * K[i]);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) K[i] *= L.s_acc;
        }
    }
}

// `[o *] k [+ WhiteNoise]` after the kernel function (PnEpiParams::post_*); `eq`: bit i set where entry i's points
// are equal in every coordinate
template <int N>
__device__ __forceinline__ void post_tile(double (&K)[N], const PnEpiParams& P, unsigned eq) {
    if (P.post_flags & 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) K[i] *= P.post_scale;
    }
    if (eq) {  // rare: the diagonal, duplicated points
#pragma unroll
        for (int i = 0; i < N; ++i)
            if ((eq >> i) & 1u) K[i] += P.post_sigma_sq;
    }
}

template <typename TOut>
__device__ __forceinline__ void store_pair(TOut* row, int64_t j0, int64_t n1, TOut a, TOut b,
                                           bool vec_ok) {
    if (j0 + 1 < n1) {
        if (vec_ok) {
            pn_store_cs2(row + j0, a, b);
        } else {
            pn_store_cs1(row + j0, a);
            pn_store_cs1(row + j0 + 1, b);
        }
    } else if (j0 < n1) {
        pn_store_cs1(row + j0, a);
    }
}


// tiling of the symmetric kernel (pncov_gram_sym.cu); the dispatcher needs the tile edge
constexpr int Y_T = 64;          // tile edge

}  // namespace

// pncov_gram_sym.cu: launches on the square 64-grid interior of a whole symmetric matrix (g.n0 == g.n1, multiples of 64)
// (_lo: d = 1 .. 4, _hi: d = 5 .. 8, one translation unit each)
cudaError_t pn_gram_sym_launch_lo(const PnGramLaunch& g, cudaStream_t stream);  // g.epi.epi selects the kernel function
cudaError_t pn_gram_sym_launch_hi(const PnGramLaunch& g, cudaStream_t stream);
cudaError_t pn_gram_sym_launch_fused_lo(const PnGramLaunch& g, const PnFusedProgram& F, cudaStream_t stream);
cudaError_t pn_gram_sym_launch_fused_hi(const PnGramLaunch& g, const PnFusedProgram& F, cudaStream_t stream);
cudaError_t pn_gram_sym_launch_pm_lo(const PnGramLaunch& g, const PnPmProgram& F, int p, cudaStream_t stream);
cudaError_t pn_gram_sym_launch_pm_hi(const PnGramLaunch& g, const PnPmProgram& F, int p, cudaStream_t stream);
inline cudaError_t pn_gram_sym_launch(const PnGramLaunch& g, cudaStream_t stream) {
    return g.d <= 4 ? pn_gram_sym_launch_lo(g, stream) : pn_gram_sym_launch_hi(g, stream);
}
inline cudaError_t pn_gram_sym_launch_fused(const PnGramLaunch& g, const PnFusedProgram& F, cudaStream_t stream) {
    return g.d <= 4 ? pn_gram_sym_launch_fused_lo(g, F, stream) : pn_gram_sym_launch_fused_hi(g, F, stream);
}
inline cudaError_t pn_gram_sym_launch_pm(const PnGramLaunch& g, const PnPmProgram& F, int p, cudaStream_t stream) {
    return g.d <= 4 ? pn_gram_sym_launch_pm_lo(g, F, p, stream) : pn_gram_sym_launch_pm_hi(g, F, p, stream);
}
