// Matrix-free product Y = K(X0, X1) @ V: K is evaluated tile by tile in registers and
// never stored.
//
// Replaces CovarianceLinearOperator._matmul
// (/root/reference/src/probnum/randprocs/covfuncs/_covariance_linear_operator.py:78-81),
// i.e. both the KeOps map-reduce (`keops_lazy_tensor @ x`, formulas built in
// _covariance_function.py:805-846, _exponentiated_quadratic.py:102-111,
// _matern.py:197-223) and the dense fallback `self.todense() @ x`.
//
// The work is O(N0 N1 d) FP64 operations on O((N0 + N1) d) bytes, so the FP64 pipe is the
// roofline.  Layout of the main kernel (d <= 8, float64):
//
//   * each thread owns MV_R rows of K; their points stay in registers for the whole kernel;
//   * column points and the matching rows of V stream through shared memory in tiles of
//     MV_TJ points, fetched by 1-D TMA bulk copies into a two-stage mbarrier pipeline, and
//     are read back as warp-wide broadcasts (every lane needs the same x1_j);
//   * per pair: d DADD + d DFMA (+ d DMUL with ARD scale factors), the epilogue of
//     pncov_math.cuh (8 FP64 ops for exp2), one DFMA per right-hand side;
//   * accumulation over j is sequential per row and per column split, partial sums of the
//     column splits are combined in a fixed order: results are run-to-run deterministic.
#include <type_traits>

#include "pncov_internal.h"

namespace {

constexpr int MV_THREADS = 128;
constexpr int MV_R = 2;     // rows per thread
constexpr int MV_TJ = 128;  // column points per pipeline stage
constexpr int MV_HDR_BYTES = 128 + 8 * (PN_EXP2_N + 2 * PN_LOG2_N);
constexpr int MV_HDR_REP_BYTES = 128 + 8 * (PN_EXP2S_DOUBLES + 2 * PN_LOG2_N);  // replicated exp2 table

template <int EPI, int N, int TAB = 0>
__device__ __forceinline__ void mv_epi_tile(double (&S)[N], const PnEpiParams& P, const double* tabE,
                                            const double* tabL) {
    PnFlags f;
    f.reset();
    double K[N];
#pragma unroll
    for (int i = 0; i < N; ++i) K[i] = PnEpilogue<EPI, TAB>::fast(S[i], P, tabE, tabL, f);
    if (f.flagged(P)) {  // rare: S == 0 under a sqrt (the diagonal), subnormal results
        // the robust path reads the 2048-entry table; TAB = 1 kernels leave it in global memory
        const double* slow_tab = TAB == 0 ? tabE : pn_exp2_table_g;
        const double* slow_log = TAB == 0 ? tabL : pn_log2_table_g;
#pragma unroll
        for (int i = 0; i < N; ++i) K[i] = pn_epilogue(S[i], P, slow_tab, slow_log);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) S[i] = K[i];
}

// epilogue of the norm-identity kernels: the factorised ExpQuad form needs no validity window (the
// host admits it only for |u| <= 2 R^2 <= 128, far inside the fast exp2 range)
template <int EPI, int N>
__device__ __forceinline__ void mm_epi_tile(double (&S)[N], const PnEpiParams& P, const double* tabE,
                                            const double* tabL) {
    if constexpr (EPI == PN_EPI_EXPQUAD_FACT) {
#pragma unroll
        for (int i = 0; i < N; ++i) S[i] = pn_exp2_prod_core(S[i], P.c_u, P, tabE);
    } else {
        mv_epi_tile<EPI>(S, P, tabE, tabL);
    }
}

template <int N>
__device__ __forceinline__ void mv_epi_dispatch(double (&S)[N], const PnEpiParams& P,
                                                const double* tabE, const double* tabL) {
    switch (P.epi) {
        case PN_EPI_EXPQUAD: mv_epi_tile<PN_EPI_EXPQUAD>(S, P, tabE, tabL); break;
        case PN_EPI_MATERN0: mv_epi_tile<PN_EPI_MATERN0>(S, P, tabE, tabL); break;
        case PN_EPI_MATERN1: mv_epi_tile<PN_EPI_MATERN1>(S, P, tabE, tabL); break;
        case PN_EPI_MATERN2: mv_epi_tile<PN_EPI_MATERN2>(S, P, tabE, tabL); break;
        case PN_EPI_MATERNP: mv_epi_tile<PN_EPI_MATERNP>(S, P, tabE, tabL); break;
        case PN_EPI_RATQUAD1: mv_epi_tile<PN_EPI_RATQUAD1>(S, P, tabE, tabL); break;
        case PN_EPI_RATQUAD2: mv_epi_tile<PN_EPI_RATQUAD2>(S, P, tabE, tabL); break;
        default: mv_epi_tile<PN_EPI_RATQUADG>(S, P, tabE, tabL); break;
    }
}

// ---- main kernel: float64, compile-time d in 1..8 -------------------------------------------
// IDENT = false: direct differences, S = sum_m ((x0_m - x1_m) s_m)^2 -- 2d (3d with ARD) FP64
//                instructions per pair, exact like the reference; every kernel family.
// IDENT = true : norm identity on centred, pre-scaled operands (a12 / b12, 12 doubles per
//                point): S = (|xt|^2 + |yt|^2) + sum_m xt_m (-2 yt_m) -- 1 + d instructions
//                per pair.  ExpQuad / RatQuad only, and only when the host-side bound on the
//                cancellation error holds (pncov_api.cu: mma_prepare).
// The same identity with the dot products on DMMA is matvec_mma_kernel below, which is the
// default on B200 (732 vs 621 Gpair/s at N = 2^20, d = 8; profiles/README.md); this DFMA form
// is selected with PNCOV_MATVEC_IDENT=dfma.
constexpr int MV_LD12 = 12;

template <int D, bool ARD, int KC, int EPI, bool IDENT>
__global__ void __launch_bounds__(MV_THREADS)
    matvec_kernel(const __grid_constant__ PnMatvecLaunch a, int tiles_per_split) {
    constexpr int XLD = IDENT ? MV_LD12 : D;  // doubles per column point in the streamed tile
    constexpr int TAB = 0;  // replicated conflict-free table (TAB = 1) measured slower: tools/mv_variants.cu
    constexpr int HDR = TAB ? MV_HDR_REP_BYTES : MV_HDR_BYTES;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + (TAB ? PN_EXP2S_DOUBLES : PN_EXP2_N);
    const double* tab_epi = TAB ? tabE + (tid & 15) : tabE;
    constexpr int STAGE_DOUBLES = MV_TJ * (XLD + KC);
    double* stage0 = reinterpret_cast<double*>(smem + HDR);

    const int64_t n_jt = a.n1pad / MV_TJ;
    const int64_t jt0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < n_jt) ? jt0 + tiles_per_split : n_jt;

    const double* x1g = static_cast<const double*>(IDENT ? a.b12 : a.x1_aos);
    const double* vg = static_cast<const double*>(a.v);
    constexpr uint32_t X_BYTES = MV_TJ * XLD * 8, V_BYTES = MV_TJ * KC * 8;

    if (tid == 0) {
        pn_mbar_init(&bars[0], 1);
        pn_mbar_init(&bars[1], 1);
        pn_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && jt0 < jt1) {
        pn_mbar_expect_tx(&bars[0], X_BYTES + V_BYTES);
        pn_bulk_g2s(stage0, x1g + jt0 * MV_TJ * XLD, X_BYTES, &bars[0]);
        pn_bulk_g2s(stage0 + MV_TJ * XLD, vg + jt0 * MV_TJ * KC, V_BYTES, &bars[0]);
    }
    if constexpr (TAB) pn_stage_tables_rep(tabE, tabL, tid, MV_THREADS, EPI == PN_EPI_RATQUADG);
    else pn_stage_tables(tabE, tabL, tid, MV_THREADS);

    // this thread's rows (the row point set is zero padded to a multiple of PN_PAD rows)
    const int64_t row0 = (int64_t)blockIdx.x * (MV_THREADS * MV_R) + tid;
    double xr[MV_R][D];
    double nxr[MV_R];
    {
        const double* x0g = static_cast<const double*>(IDENT ? a.a12 : a.x0_aos);
#pragma unroll
        for (int r = 0; r < MV_R; ++r) {
#pragma unroll
            for (int m = 0; m < D; ++m) xr[r][m] = x0g[(row0 + (int64_t)r * MV_THREADS) * XLD + m];
            nxr[r] = IDENT ? x0g[(row0 + (int64_t)r * MV_THREADS) * XLD + 8] : 0.0;
        }
    }
    double sc[D];
    if constexpr (ARD) {
#pragma unroll
        for (int m = 0; m < D; ++m) sc[m] = static_cast<const double*>(a.scale)[m];
    }
    double acc[MV_R][KC];
#pragma unroll
    for (int r = 0; r < MV_R; ++r)
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[r][k] = 0.0;
    __syncthreads();

    int it = 0;
    for (int64_t jt = jt0; jt < jt1; ++jt, ++it) {
        const int buf = it & 1;
        if (tid == 0 && jt + 1 < jt1) {
            double* dst = stage0 + (size_t)(buf ^ 1) * STAGE_DOUBLES;
            pn_mbar_expect_tx(&bars[buf ^ 1], X_BYTES + V_BYTES);
            pn_bulk_g2s(dst, x1g + (jt + 1) * MV_TJ * XLD, X_BYTES, &bars[buf ^ 1]);
            pn_bulk_g2s(dst + MV_TJ * XLD, vg + (jt + 1) * MV_TJ * KC, V_BYTES, &bars[buf ^ 1]);
        }
        pn_mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        const double* xt = stage0 + (size_t)buf * STAGE_DOUBLES;
        const double* vt = xt + MV_TJ * XLD;

#pragma unroll 1
        for (int jj = 0; jj < MV_TJ; jj += 2) {
            double xb[2][D];
            double nyb[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if constexpr (XLD % 2 == 0) {  // rows of the tile are 16-byte aligned
#pragma unroll
                    for (int m = 0; m + 1 < D; m += 2) {
                        const double2 t = *reinterpret_cast<const double2*>(xt + (jj + c) * XLD + m);
                        xb[c][m] = t.x;
                        xb[c][m + 1] = t.y;
                    }
                    if constexpr (D % 2 == 1) xb[c][D - 1] = xt[(jj + c) * XLD + D - 1];
                } else {
#pragma unroll
                    for (int m = 0; m < D; ++m) xb[c][m] = xt[(jj + c) * XLD + m];
                }
                nyb[c] = IDENT ? xt[(jj + c) * XLD + 8] : 0.0;
            }
            double S[2 * MV_R];
#pragma unroll
            for (int r = 0; r < MV_R; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    double s;
                    if constexpr (IDENT) {
                        s = nxr[r] + nyb[c];
#pragma unroll
                        for (int m = 0; m < D; ++m) s = pn_fma(xr[r][m], xb[c][m], s);
                    } else {
                        s = 0.0;
#pragma unroll
                        for (int m = 0; m < D; ++m) {
                            double df = xr[r][m] - xb[c][m];
                            if constexpr (ARD) df *= sc[m];
                            s = pn_fma(df, df, s);
                        }
                    }
                    S[2 * r + c] = s;
                }
            mv_epi_tile<EPI, 2 * MV_R, TAB>(S, a.epi, tab_epi, tabL);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const double vk = vt[(jj + c) * KC + k];
#pragma unroll
                    for (int r = 0; r < MV_R; ++r) acc[r][k] = pn_fma(S[2 * r + c], vk, acc[r][k]);
                }
        }
        __syncthreads();
    }

    double* dst = (a.jsplit > 1) ? static_cast<double*>(a.partial) + (size_t)blockIdx.y * a.n0 * KC
                                 : static_cast<double*>(a.y);
#pragma unroll
    for (int r = 0; r < MV_R; ++r) {
        const int64_t row = row0 + (int64_t)r * MV_THREADS;
        if (row < a.n0) {
#pragma unroll
            for (int k = 0; k < KC; ++k) dst[row * KC + k] = acc[r][k];
        }
    }
}

// ---- norm-identity kernel: S = |x|^2 + |y|^2 - 2 x.y with the inner products on DMMA ---------
// ExpQuad / RatQuad, float64, d <= 8, used only when the host-side bound on the cancellation
// error (pncov_api.cu: mma_admissible) keeps K within the parity contract.  On B200 the FP64
// tensor path shares the FP64 pipe with DFMA (profiles/r01_microbench.json), so DMMA buys
// FP64-pipe time only through the identity itself (d FMAs instead of d DADD + d DFMA per
// pair) plus ~8 % higher DMMA issue rate, and it frees issue slots: one DMMA.8x8x4 replaces
// eight DFMA warp instructions.
//
// Fragment layout of mma.m8n8k4.f64 (g = lane / 4, q = lane % 4):
//   A[g][q]  B[q][g]  C[g][2q], C[g][2q + 1]
// A warp owns MM_RG groups of 8 rows; per step of 8 columns every lane ends up with two
// entries S[g][2q], S[g][2q + 1] per row group, runs the epilogue on them and accumulates
// K * v for "its" row g; the four lanes that share a row are reduced with shuffles once,
// at the end of the kernel.
constexpr int MM_THREADS = 128;
constexpr int MM_RG = 4;                          // row groups of 8 per warp
constexpr int MM_ROWS = (MM_THREADS / 32) * MM_RG * 8;  // rows per CTA = 128
constexpr int MM_NS = 2;                          // 8-column steps per inner iteration (tools/mv_variants.cu)
// Operand layout and tile width by the number of k = 4 steps of the contraction (d <= 4 KSTEPS):
// a point occupies LD doubles [xt(4 KSTEPS, at least 8), |xt|^2, exp(-|xt|^2), 0, 0]; LD = 4 or 12
// (mod 16), which keeps the B-fragment loads bank-conflict free.  Wider operands get narrower column
// tiles so that the two pipeline stages stay around 25 KB.
template <int KSTEPS>
struct MmCfg {
    static constexpr int DPAD = KSTEPS < 2 ? 8 : 4 * KSTEPS;
    static constexpr int LD = DPAD + 4;
    static constexpr int NRM = DPAD;                      // index of |xt|^2 (the factor follows it)
    static constexpr int TJ = KSTEPS <= 2 ? 128 : (KSTEPS <= 4 ? 64 : 32);  // columns per pipeline stage
    static constexpr int MINB = KSTEPS <= 2 ? 4 : (KSTEPS <= 8 ? 2 : 1);  // CTAs per SM the symmetric kernel is compiled for
};

__device__ __forceinline__ void pn_dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int KSTEPS, int KC, int EPI>
__global__ void __launch_bounds__(MM_THREADS)
    matvec_mma_kernel(const __grid_constant__ PnMatvecLaunch a, int tiles_per_split) {
    using Cfg = MmCfg<KSTEPS>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    // eight right-hand sides: K V_J as a DMMA itself (see matvec_mma_symd_kernel), V in rows of 10 doubles
    constexpr bool ACC_MMA = KC == 8;
    constexpr int KCP = ACC_MMA ? 10 : KC;
    constexpr int NACC = ACC_MMA ? 2 : KC;
    constexpr int STAGE_DOUBLES = Cfg::TJ * (Cfg::LD + KCP);
    double* stage0 = reinterpret_cast<double*>(smem + MV_HDR_BYTES);

    const int64_t n_jt = a.n1pad / Cfg::TJ;
    const int64_t jt0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < n_jt) ? jt0 + tiles_per_split : n_jt;
    const double* bg = static_cast<const double*>(a.b12);
    const double* vg = static_cast<const double*>(a.v);
    constexpr uint32_t B_BYTES = Cfg::TJ * Cfg::LD * 8, V_BYTES = Cfg::TJ * KCP * 8;

    if (tid == 0) {
        pn_mbar_init(&bars[0], 1);
        pn_mbar_init(&bars[1], 1);
        pn_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && jt0 < jt1) {
        pn_mbar_expect_tx(&bars[0], B_BYTES + V_BYTES);
        pn_bulk_g2s(stage0, bg + jt0 * Cfg::TJ * Cfg::LD, B_BYTES, &bars[0]);
        pn_bulk_g2s(stage0 + Cfg::TJ * Cfg::LD, vg + jt0 * Cfg::TJ * KCP, V_BYTES, &bars[0]);
    }
    pn_stage_tables(tabE, tabL, tid, MM_THREADS);

    const int64_t row0 = (int64_t)blockIdx.x * MM_ROWS + warp * (MM_RG * 8);
    double af[MM_RG][KSTEPS], nx[MM_RG], acc[MM_RG][NACC];
    {
        const double* ag = static_cast<const double*>(a.a12);
#pragma unroll
        for (int rg = 0; rg < MM_RG; ++rg) {
            const double* pr = ag + (row0 + rg * 8 + g) * Cfg::LD;  // rows are padded to PN_PAD
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) af[rg][kk] = pr[4 * kk + q];
            nx[rg] = pr[Cfg::NRM];
#pragma unroll
            for (int k = 0; k < NACC; ++k) acc[rg][k] = 0.0;
        }
    }
    __syncthreads();

    int it = 0;
    for (int64_t jt = jt0; jt < jt1; ++jt, ++it) {
        const int buf = it & 1;
        if (tid == 0 && jt + 1 < jt1) {
            double* dst = stage0 + (size_t)(buf ^ 1) * STAGE_DOUBLES;
            pn_mbar_expect_tx(&bars[buf ^ 1], B_BYTES + V_BYTES);
            pn_bulk_g2s(dst, bg + (jt + 1) * Cfg::TJ * Cfg::LD, B_BYTES, &bars[buf ^ 1]);
            pn_bulk_g2s(dst + Cfg::TJ * Cfg::LD, vg + (jt + 1) * Cfg::TJ * KCP, V_BYTES, &bars[buf ^ 1]);
        }
        pn_mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        const double* bt = stage0 + (size_t)buf * STAGE_DOUBLES;
        const double* vt = bt + Cfg::TJ * Cfg::LD;

#pragma unroll 1
        for (int jj = 0; jj < Cfg::TJ; jj += 8 * MM_NS) {
            // MM_NS steps of 8 columns: all DMMAs first, then the epilogues, so that the
            // epilogue chains of one step overlap the DMMA latency of the next
            double S[MM_NS][2 * MM_RG];
#pragma unroll
            for (int s = 0; s < MM_NS; ++s) {
                const int j = jj + 8 * s;
                double bf[KSTEPS];
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) bf[kk] = bt[(j + g) * Cfg::LD + 4 * kk + q];
                const double ny0 = bt[(j + 2 * q) * Cfg::LD + Cfg::NRM];
                const double ny1 = bt[(j + 2 * q + 1) * Cfg::LD + Cfg::NRM];
#pragma unroll
                for (int rg = 0; rg < MM_RG; ++rg) {
                    double c0, c1;
                    if constexpr (EPI == PN_EPI_EXPQUAD_FACT) {
                        c0 = 0.0;
                        c1 = 0.0;
                    } else {
                        c0 = nx[rg] + ny0;
                        c1 = nx[rg] + ny1;
                    }
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk) pn_dmma884(c0, c1, af[rg][kk], bf[kk]);
                    S[s][2 * rg] = c0;
                    S[s][2 * rg + 1] = c1;
                }
            }
#pragma unroll
            for (int s = 0; s < MM_NS; ++s) {
                const int j = jj + 8 * s;
                mm_epi_tile<EPI>(S[s], a.epi, tabE, tabL);
                if constexpr (ACC_MMA) {
                    // K as the A operand (even / odd columns), V[column][right-hand side g] as B: two DMMAs per 8 x 8 block
                    const double be = vt[(j + 2 * q) * KCP + g], bo = vt[(j + 2 * q + 1) * KCP + g];
#pragma unroll
                    for (int rg = 0; rg < MM_RG; ++rg) {
                        pn_dmma884(acc[rg][0], acc[rg][1], S[s][2 * rg], be);
                        pn_dmma884(acc[rg][0], acc[rg][1], S[s][2 * rg + 1], bo);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        const double v0 = vt[(j + 2 * q) * KC + k];
                        const double v1 = vt[(j + 2 * q + 1) * KC + k];
#pragma unroll
                        for (int rg = 0; rg < MM_RG; ++rg) {
                            acc[rg][k] = pn_fma(S[s][2 * rg], v0, acc[rg][k]);
                            acc[rg][k] = pn_fma(S[s][2 * rg + 1], v1, acc[rg][k]);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    double* dst = (a.jsplit > 1) ? static_cast<double*>(a.partial) + (size_t)blockIdx.y * a.n0 * KC
                                 : static_cast<double*>(a.y);
#pragma unroll
    for (int rg = 0; rg < MM_RG; ++rg) {
        const int64_t row = row0 + rg * 8 + g;
        if constexpr (ACC_MMA) {  // lane (g, q) holds right-hand sides 2 q, 2 q + 1 of row g, complete
            if (row < a.n0) *reinterpret_cast<double2*>(dst + row * KC + 2 * q) = make_double2(acc[rg][0], acc[rg][1]);
        } else {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                double s = acc[rg][k];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (q == 0 && row < a.n0) dst[row * KC + k] = s;
            }
        }
    }
}

// ---- symmetric variant: x1 is x0, one right-hand side ------------------------------------------------
// K is symmetric, so an off-diagonal tile pair (I, J) / (J, I) needs its kernel values only once:
// the CTA of row block I walks the column tiles J >= I, accumulates y_I += K_IJ v_J in registers
// as above and ALSO y_J += K_IJ^T v_I: per 8-column step one more DFMA per pair, a shuffle
// reduction over the eight lanes that hold the same columns, a per-warp slot in shared memory,
// and once per tile one update of the integer planes per column (pn_bin_add, pncov_common.cuh: the
// partial sum split exactly into two integers and added with 64-bit integer atomics).  The exp2 /
// DMMA work per ordered pair halves: 19.25 instead of 35 FP64-pipe slots per two ordered pairs.  Row
// sums leave through the same planes at the end of the CTA.  Integer addition is associative, so the
// product is bit-for-bit reproducible although the CTAs retire in any order; the caller zeroes the
// planes and converts them afterwards (pn_launch_unbin).  Row blocks are dealt round-robin to
// `nparts` callers (multi-GPU: rank r takes blocks I = r mod nparts and the planes of all ranks are
// summed by an all-reduce: the same bits for every number of ranks).
template <int KSTEPS, int EPI>
__global__ void __launch_bounds__(MM_THREADS, MmCfg<KSTEPS>::MINB)
    matvec_mma_sym_kernel(const __grid_constant__ PnMatvecLaunch a, int tiles_per_split, int part, int nparts) {
    using Cfg = MmCfg<KSTEPS>;
    constexpr int TPB = MM_ROWS / Cfg::TJ;  // column tiles per row block: the diagonal spans TPB tiles
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    constexpr int STAGE_DOUBLES = Cfg::TJ * (Cfg::LD + 1);
    double* stage0 = reinterpret_cast<double*>(smem + MV_HDR_BYTES);
    double* colw = stage0 + 2 * STAGE_DOUBLES;  // [2][warps][Cfg::TJ] per-warp column sums of a tile

    static_assert(MM_ROWS % Cfg::TJ == 0, "a row block must span whole column tiles");
    const int64_t n_jt = a.n1pad / Cfg::TJ;
    const int64_t bx = (int64_t)blockIdx.x * nparts + part;  // this CTA's row block; diagonal tiles TPB bx .. TPB bx + TPB - 1
    if (bx * TPB >= n_jt) return;
    const int64_t jdiag_end = (bx + 1) * TPB;               // first tile right of the diagonal block
    int64_t jt0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < n_jt) ? jt0 + tiles_per_split : n_jt;
    if (jt0 < bx * TPB) jt0 = bx * TPB;
    if (jt0 >= jt1) return;
    const double* bg = static_cast<const double*>(a.b12);
    const double* vg = static_cast<const double*>(a.v);
    long long* const bhi = a.bins;
    long long* const blo = a.bins + a.n0;
    double* const bnf = reinterpret_cast<double*>(a.bins + 2 * a.n0);
    const double bound0 = (double)a.n1 * *a.vmax;
    constexpr bool FACT = EPI == PN_EPI_EXPQUAD_FACT;
    constexpr uint32_t B_BYTES = Cfg::TJ * Cfg::LD * 8, V_BYTES = Cfg::TJ * 8;

    if (tid == 0) {
        pn_mbar_init(&bars[0], 1);
        pn_mbar_init(&bars[1], 1);
        pn_mbar_init(&bars[2], 1);
        pn_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        pn_mbar_expect_tx(&bars[0], B_BYTES + V_BYTES);
        pn_bulk_g2s(stage0, bg + jt0 * Cfg::TJ * Cfg::LD, B_BYTES, &bars[0]);
        pn_bulk_g2s(stage0 + Cfg::TJ * Cfg::LD, vg + jt0 * Cfg::TJ, V_BYTES, &bars[0]);
        pn_stage_tables_tma(tabE, tabL, &bars[2], EPI == PN_EPI_RATQUADG);
    }

    const int64_t row0 = bx * MM_ROWS + warp * (MM_RG * 8);
    double af[MM_RG][KSTEPS], nx[MM_RG], vrow[MM_RG], acc[MM_RG];
    {
        const double* ag = static_cast<const double*>(a.a12);
#pragma unroll
        for (int rg = 0; rg < MM_RG; ++rg) {
            const int64_t row = row0 + rg * 8 + g;
            const double* pr = ag + row * Cfg::LD;  // rows are padded to PN_PAD
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) af[rg][kk] = pr[4 * kk + q];
            nx[rg] = pr[Cfg::NRM];
            vrow[rg] = vg[row];  // zero in the padding, so padded rows add nothing to column sums
            acc[rg] = 0.0;
        }
    }
    pn_mbar_wait(&bars[2], 0);

    int it = 0;
    for (int64_t jt = jt0; jt < jt1; ++jt, ++it) {
        const int buf = it & 1;
        if (tid == 0 && jt + 1 < jt1) {
            double* dst = stage0 + (size_t)(buf ^ 1) * STAGE_DOUBLES;
            pn_mbar_expect_tx(&bars[buf ^ 1], B_BYTES + V_BYTES);
            pn_bulk_g2s(dst, bg + (jt + 1) * Cfg::TJ * Cfg::LD, B_BYTES, &bars[buf ^ 1]);
            pn_bulk_g2s(dst + Cfg::TJ * Cfg::LD, vg + (jt + 1) * Cfg::TJ, V_BYTES, &bars[buf ^ 1]);
        }
        // column sums of the previous tile: one integer-plane update per column (the slot was completed before
        // the barrier that ended the previous iteration)
        if (it > 0 && jt - 1 >= jdiag_end && tid < Cfg::TJ) {
            const double* cw = colw + (buf ^ 1) * (MM_THREADS / 32) * Cfg::TJ;
            double s = cw[tid];
#pragma unroll
            for (int w = 1; w < MM_THREADS / 32; ++w) s += cw[w * Cfg::TJ + tid];
            const int64_t col = (jt - 1) * Cfg::TJ + tid;
            if (col < a.n1) pn_bin_add(bhi, blo, bnf, col, s, pn_bin_exponent(bound0, FACT ? bg[col * Cfg::LD + Cfg::NRM + 1] : 1.0));
        }
        pn_mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        const double* bt = stage0 + (size_t)buf * STAGE_DOUBLES;
        const double* vt = bt + Cfg::TJ * Cfg::LD;
        double* cw = colw + buf * (MM_THREADS / 32) * Cfg::TJ + warp * Cfg::TJ;
        const bool mirror = jt >= jdiag_end;  // uniform; the diagonal tiles hold both (i, j) and (j, i) themselves

#pragma unroll 1
        for (int jj = 0; jj < Cfg::TJ; jj += 8 * MM_NS) {
            double S[MM_NS][2 * MM_RG];
#pragma unroll
            for (int s = 0; s < MM_NS; ++s) {
                const int j = jj + 8 * s;
                double bf[KSTEPS];
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) bf[kk] = bt[(j + g) * Cfg::LD + 4 * kk + q];
                const double ny0 = bt[(j + 2 * q) * Cfg::LD + Cfg::NRM];
                const double ny1 = bt[(j + 2 * q + 1) * Cfg::LD + Cfg::NRM];
#pragma unroll
                for (int rg = 0; rg < MM_RG; ++rg) {
                    double c0, c1;
                    if constexpr (EPI == PN_EPI_EXPQUAD_FACT) {
                        c0 = 0.0;
                        c1 = 0.0;
                    } else {
                        c0 = nx[rg] + ny0;
                        c1 = nx[rg] + ny1;
                    }
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk) pn_dmma884(c0, c1, af[rg][kk], bf[kk]);
                    S[s][2 * rg] = c0;
                    S[s][2 * rg + 1] = c1;
                }
            }
            double tc[MM_NS][2];
#pragma unroll
            for (int s = 0; s < MM_NS; ++s) {
                const int j = jj + 8 * s;
                mm_epi_tile<EPI>(S[s], a.epi, tabE, tabL);
                const double2 v01 = *reinterpret_cast<const double2*>(vt + j + 2 * q);
                double t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int rg = 0; rg < MM_RG; ++rg) {
                    acc[rg] = pn_fma(S[s][2 * rg], v01.x, acc[rg]);
                    acc[rg] = pn_fma(S[s][2 * rg + 1], v01.y, acc[rg]);
                    t0 = pn_fma(S[s][2 * rg], vrow[rg], t0);
                    t1 = pn_fma(S[s][2 * rg + 1], vrow[rg], t1);
                }
                tc[s][0] = t0;
                tc[s][1] = t1;
            }
            if (mirror) {
                // Column sums over the 8 lanes g = 0..7 that hold the same columns, for both steps at
                // once as a butterfly: round 1 splits the two steps over g bit 2, round 2 the two
                // columns over g bit 1, round 3 finishes: 4 shuffles + 4 DADDs instead of 12 + 12.
                static_assert(MM_NS == 2, "the butterfly below folds exactly two 8-column steps");
                const bool hi = (g & 4) != 0, mid = (g & 2) != 0;
                double k0 = (hi ? tc[1][0] : tc[0][0]) + __shfl_xor_sync(0xffffffffu, hi ? tc[0][0] : tc[1][0], 16);
                double k1 = (hi ? tc[1][1] : tc[0][1]) + __shfl_xor_sync(0xffffffffu, hi ? tc[0][1] : tc[1][1], 16);
                double kk = (mid ? k1 : k0) + __shfl_xor_sync(0xffffffffu, mid ? k0 : k1, 8);
                kk += __shfl_xor_sync(0xffffffffu, kk, 4);
                if ((g & 1) == 0) cw[jj + (hi ? 8 : 0) + 2 * q + (mid ? 1 : 0)] = kk;
            }
        }
        __syncthreads();
    }
    // column sums of the last tile
    if (jt1 - 1 >= jdiag_end && tid < Cfg::TJ) {
        const double* cw = colw + ((it - 1) & 1) * (MM_THREADS / 32) * Cfg::TJ;
        double s = cw[tid];
#pragma unroll
        for (int w = 1; w < MM_THREADS / 32; ++w) s += cw[w * Cfg::TJ + tid];
        const int64_t col = (jt1 - 1) * Cfg::TJ + tid;
        if (col < a.n1) pn_bin_add(bhi, blo, bnf, col, s, pn_bin_exponent(bound0, FACT ? bg[col * Cfg::LD + Cfg::NRM + 1] : 1.0));
    }
    // row sums
#pragma unroll
    for (int rg = 0; rg < MM_RG; ++rg) {
        const int64_t row = row0 + rg * 8 + g;
        double s = acc[rg];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0 && row < a.n0)
            pn_bin_add(bhi, blo, bnf, row, s,
                       pn_bin_exponent(bound0, FACT ? static_cast<const double*>(a.a12)[row * Cfg::LD + Cfg::NRM + 1] : 1.0));
    }
}

// ---- symmetric variant for two to eight right-hand sides per pass: both accumulations on the tensor cores ----
// With eight columns the products K_IJ V_J and K_IJ^T V_I are small GEMMs themselves.  The kernel values of an 8 x 8
// block sit in the C-fragment layout of the distance DMMA (lane (g, q): row g, columns 2 q, 2 q + 1), which is an
// A operand as it stands (k-index = column, even and odd columns as two DMMAs): y_I += K V_J costs two DMMAs per
// block for all eight right-hand sides, accumulators two doubles per row group.  For K^T V_I the block is transposed
// through 512 bytes of shared memory per row group (one 16-byte store, two 8-byte loads per lane), the B operand
// -- V of this lane's rows -- lives in registers for the whole kernel, and the 8 x 8 result IS the tile's column-sum
// fragment: no shuffle butterfly (a DFMA form with a butterfly over the eight lanes of a column pair spent as many
// instructions on selects and shuffles as on arithmetic: 168 ms against 82 ms for eight columns at N = 262144, where
// one column takes 44 ms).  Column sums and row sums leave through the integer planes (pn_bin_add) as everywhere.
// Fewer than eight columns are padded with zeros: the tensor-core tile is eight wide either way.
constexpr int MD_KC = 8, MD_KCP = 10;  // columns per pass, doubles per packed row of V (conflict-free B-fragment loads)
constexpr int MD_TJ = 64;               // columns per tile: 16 KB of per-warp column-sum slots, three CTAs per SM
template <int KSTEPS, int EPI>
__global__ void __launch_bounds__(MM_THREADS, 3)
    matvec_mma_symd_kernel(const __grid_constant__ PnMatvecLaunch a, int tiles_per_split, int part, int nparts) {
    using Cfg = MmCfg<KSTEPS>;
    constexpr int TJ = MD_TJ, TPB = MM_ROWS / MD_TJ, LD = Cfg::LD, KC = MD_KC, KCP = MD_KCP;
    constexpr bool FACT = EPI == PN_EPI_EXPQUAD_FACT;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    constexpr int STAGE_DOUBLES = TJ * (LD + KCP);
    double* stage0 = reinterpret_cast<double*>(smem + MV_HDR_BYTES);
    double* colw = stage0 + 2 * STAGE_DOUBLES;                      // [warps][TJ][KC]
    double* tr = colw + (MM_THREADS / 32) * TJ * KC + warp * (MM_RG * 64);  // per warp: [MM_RG][8][8] blocks

    const int64_t n_jt = a.n1pad / TJ;
    const int64_t bx = (int64_t)blockIdx.x * nparts + part;  // this CTA's row block; diagonal tiles TPB bx .. TPB bx + TPB - 1
    if (bx * TPB >= n_jt) return;
    const int64_t jdiag_end = (bx + 1) * TPB;
    int64_t jt0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < n_jt) ? jt0 + tiles_per_split : n_jt;
    if (jt0 < bx * TPB) jt0 = bx * TPB;
    if (jt0 >= jt1) return;
    const double* bg = static_cast<const double*>(a.b12);
    const double* vg = static_cast<const double*>(a.v);  // (n1pad, KCP) packed
    const int64_t plane = a.n0 * KC;
    long long* const bhi = a.bins;
    long long* const blo = a.bins + plane;
    double* const bnf = reinterpret_cast<double*>(a.bins + 2 * plane);
    const double bound0 = (double)a.n1 * *a.vmax;
    constexpr uint32_t B_BYTES = TJ * LD * 8, V_BYTES = TJ * KCP * 8;

    if (tid == 0) {
        pn_mbar_init(&bars[0], 1);
        pn_mbar_init(&bars[1], 1);
        pn_mbar_init(&bars[2], 1);
        pn_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        pn_mbar_expect_tx(&bars[0], B_BYTES + V_BYTES);
        pn_bulk_g2s(stage0, bg + jt0 * TJ * LD, B_BYTES, &bars[0]);
        pn_bulk_g2s(stage0 + TJ * LD, vg + jt0 * TJ * KCP, V_BYTES, &bars[0]);
        pn_stage_tables_tma(tabE, tabL, &bars[2], EPI == PN_EPI_RATQUADG);
    }

    const int64_t row0 = bx * MM_ROWS + warp * (MM_RG * 8);
    double af[MM_RG][KSTEPS], nx[MM_RG], vrb[MM_RG][2], acc[MM_RG][2];
    {
        const double* ag = static_cast<const double*>(a.a12);
#pragma unroll
        for (int rg = 0; rg < MM_RG; ++rg) {
            const double* pr = ag + (row0 + rg * 8 + g) * LD;  // rows are padded to PN_PAD
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) af[rg][kk] = pr[4 * kk + q];
            nx[rg] = pr[Cfg::NRM];
            // B operand of the transposed product: V[row q (+ 4) of the group][right-hand side g]; zero in the padding
            vrb[rg][0] = vg[(row0 + rg * 8 + q) * KCP + g];
            vrb[rg][1] = vg[(row0 + rg * 8 + q + 4) * KCP + g];
            acc[rg][0] = 0.0;
            acc[rg][1] = 0.0;
        }
    }
    pn_mbar_wait(&bars[2], 0);

    int it = 0;
    for (int64_t jt = jt0; jt < jt1; ++jt, ++it) {
        const int buf = it & 1;
        if (tid == 0 && jt + 1 < jt1) {
            double* dst = stage0 + (size_t)(buf ^ 1) * STAGE_DOUBLES;
            pn_mbar_expect_tx(&bars[buf ^ 1], B_BYTES + V_BYTES);
            pn_bulk_g2s(dst, bg + (jt + 1) * TJ * LD, B_BYTES, &bars[buf ^ 1]);
            pn_bulk_g2s(dst + TJ * LD, vg + (jt + 1) * TJ * KCP, V_BYTES, &bars[buf ^ 1]);
        }
        // column sums of the previous tile (its slots were completed before the barrier that ended that iteration;
        // the barrier below keeps this tile's sums out of them until everybody has read)
        if (it > 0 && jt - 1 >= jdiag_end) {
            const double* cwp = colw;
            for (int e = tid; e < TJ * KC; e += MM_THREADS) {
                double s = cwp[e];
#pragma unroll
                for (int w = 1; w < MM_THREADS / 32; ++w) s += cwp[w * TJ * KC + e];
                const int64_t col = (jt - 1) * TJ + e / KC;
                if (col < a.n1)
                    pn_bin_add(bhi, blo, bnf, col * KC + e % KC, s, pn_bin_exponent(bound0, FACT ? bg[col * LD + Cfg::NRM + 1] : 1.0));
            }
        }
        if (it > 0) __syncthreads();
        pn_mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        const double* bt = stage0 + (size_t)buf * STAGE_DOUBLES;
        const double* vt = bt + TJ * LD;
        double* cw = colw + warp * TJ * KC;
        const bool mirror = jt >= jdiag_end;  // uniform; the diagonal tiles hold both (i, j) and (j, i) themselves

#pragma unroll 1
        for (int j = 0; j < TJ; j += 8) {
            double S[2 * MM_RG];
            {
                double bf[KSTEPS];
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) bf[kk] = bt[(j + g) * LD + 4 * kk + q];
                const double ny0 = bt[(j + 2 * q) * LD + Cfg::NRM];
                const double ny1 = bt[(j + 2 * q + 1) * LD + Cfg::NRM];
#pragma unroll
                for (int rg = 0; rg < MM_RG; ++rg) {
                    double c0, c1;
                    if constexpr (FACT) {
                        c0 = 0.0;
                        c1 = 0.0;
                    } else {
                        c0 = nx[rg] + ny0;
                        c1 = nx[rg] + ny1;
                    }
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk) pn_dmma884(c0, c1, af[rg][kk], bf[kk]);
                    S[2 * rg] = c0;
                    S[2 * rg + 1] = c1;
                }
            }
            mm_epi_tile<EPI>(S, a.epi, tabE, tabL);
            // y_I += K V_J: K as the A operand (even / odd columns), V[column][right-hand side g] as B
            const double be = vt[(j + 2 * q) * KCP + g], bo = vt[(j + 2 * q + 1) * KCP + g];
#pragma unroll
            for (int rg = 0; rg < MM_RG; ++rg) {
                pn_dmma884(acc[rg][0], acc[rg][1], S[2 * rg], be);
                pn_dmma884(acc[rg][0], acc[rg][1], S[2 * rg + 1], bo);
            }
            if (mirror) {
                // y_J += K^T V_I: the blocks transposed through shared memory, V of this warp's rows from registers
#pragma unroll
                for (int rg = 0; rg < MM_RG; ++rg)
                    *reinterpret_cast<double2*>(tr + rg * 64 + g * 8 + 2 * q) = make_double2(S[2 * rg], S[2 * rg + 1]);
                __syncwarp();
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int rg = 0; rg < MM_RG; ++rg) {
                    pn_dmma884(c0, c1, tr[rg * 64 + q * 8 + g], vrb[rg][0]);
                    pn_dmma884(c0, c1, tr[rg * 64 + (q + 4) * 8 + g], vrb[rg][1]);
                }
                *reinterpret_cast<double2*>(cw + (j + g) * KC + 2 * q) = make_double2(c0, c1);
                __syncwarp();
            }
        }
        __syncthreads();
    }
    // column sums of the last tile
    if (jt1 - 1 >= jdiag_end) {
        const double* cwp = colw;
        for (int e = tid; e < TJ * KC; e += MM_THREADS) {
            double s = cwp[e];
#pragma unroll
            for (int w = 1; w < MM_THREADS / 32; ++w) s += cwp[w * TJ * KC + e];
            const int64_t col = (jt1 - 1) * TJ + e / KC;
            if (col < a.n1)
                pn_bin_add(bhi, blo, bnf, col * KC + e % KC, s, pn_bin_exponent(bound0, FACT ? bg[col * LD + Cfg::NRM + 1] : 1.0));
        }
    }
    // row sums: lane (g, q) holds rows rg * 8 + g, right-hand sides 2 q and 2 q + 1, complete
#pragma unroll
    for (int rg = 0; rg < MM_RG; ++rg) {
        const int64_t row = row0 + rg * 8 + g;
        if (row < a.n0) {
            const int E = pn_bin_exponent(bound0, FACT ? static_cast<const double*>(a.a12)[row * LD + Cfg::NRM + 1] : 1.0);
            pn_bin_add(bhi, blo, bnf, row * KC + 2 * q, acc[rg][0], E);
            pn_bin_add(bhi, blo, bnf, row * KC + 2 * q + 1, acc[rg][1], E);
        }
    }
}

// ---- symmetric variant of the direct-difference kernel (x1 is x0, one right-hand side) ----------------
// Same halving for the kernels that cannot use the norm identity (Matern: sqrt(S) near 0; ExpQuad /
// RatQuad on clouds that are wide in lengthscale units): the CTA of row block I (256 rows = two
// column tiles) walks the column tiles J >= 2I, and for tiles right of its diagonal pair also
// accumulates y_J += K_IJ^T v_I.  A thread holds K for its 2 rows x 4 columns: the DFMAs with the row
// operand give four partial column sums, two halving exchanges fold the four columns onto the four
// quarter-warps and three more shuffle rounds finish the warp sums (0.75 DADD and 1.5 SHFL per pair; two
// columns per step cost 1.25 and 2.5: the kernel is co-bound by issue slots); warps meet in shared memory
// and one integer-plane update per column leaves the CTA per tile.  Matern-5/2, d = 8: 32 -> 17.75 FP64
// instructions per ordered pair.
constexpr int MS_ROWS = MV_THREADS * MV_R;
static_assert(MS_ROWS == 2 * MV_TJ, "a row block must span exactly two column tiles");

template <int D, bool ARD, int EPI>
__global__ void __launch_bounds__(MV_THREADS)
    matvec_sym_kernel(const __grid_constant__ PnMatvecLaunch a, int tiles_per_split, int part, int nparts) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    constexpr int STAGE_DOUBLES = MV_TJ * (D + 1);
    double* stage0 = reinterpret_cast<double*>(smem + MV_HDR_BYTES);
    double* colw = stage0 + 2 * STAGE_DOUBLES + 2;  // [2][warps][MV_TJ]; +2 keeps 16-byte alignment for odd D

    const int64_t n_jt = a.n1pad / MV_TJ;
    const int64_t bx = (int64_t)blockIdx.x * nparts + part;  // row block; its diagonal tiles are 2 bx, 2 bx + 1
    if (2 * bx >= n_jt) return;
    int64_t jt0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < n_jt) ? jt0 + tiles_per_split : n_jt;
    if (jt0 < 2 * bx) jt0 = 2 * bx;
    if (jt0 >= jt1) return;
    const double* x1g = static_cast<const double*>(a.x1_aos);
    const double* vg = static_cast<const double*>(a.v);
    long long* const bhi = a.bins;
    long long* const blo = a.bins + a.n0;
    double* const bnf = reinterpret_cast<double*>(a.bins + 2 * a.n0);
    const int binE = pn_bin_exponent((double)a.n1 * *a.vmax, 1.0);  // every kernel value is <= 1
    constexpr uint32_t X_BYTES = MV_TJ * D * 8, V_BYTES = MV_TJ * 8;

    if (tid == 0) {
        pn_mbar_init(&bars[0], 1);
        pn_mbar_init(&bars[1], 1);
        pn_mbar_init(&bars[2], 1);
        pn_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        pn_mbar_expect_tx(&bars[0], X_BYTES + V_BYTES);
        pn_bulk_g2s(stage0, x1g + jt0 * MV_TJ * D, X_BYTES, &bars[0]);
        pn_bulk_g2s(stage0 + MV_TJ * D, vg + jt0 * MV_TJ, V_BYTES, &bars[0]);
        pn_stage_tables_tma(tabE, tabL, &bars[2], EPI == PN_EPI_RATQUADG);
    }
    const int64_t row0 = bx * MS_ROWS + tid;
    double xr[MV_R][D], vrow[MV_R], acc[MV_R];
    {
        const double* x0g = static_cast<const double*>(a.x0_aos);
#pragma unroll
        for (int r = 0; r < MV_R; ++r) {
            const int64_t row = row0 + (int64_t)r * MV_THREADS;  // padded to PN_PAD rows
#pragma unroll
            for (int m = 0; m < D; ++m) xr[r][m] = x0g[row * D + m];
            vrow[r] = vg[row];  // zero in the padding
            acc[r] = 0.0;
        }
    }
    double sc[D];
    if constexpr (ARD) {
#pragma unroll
        for (int m = 0; m < D; ++m) sc[m] = static_cast<const double*>(a.scale)[m];
    }
    pn_mbar_wait(&bars[2], 0);

    int it = 0;
    for (int64_t jt = jt0; jt < jt1; ++jt, ++it) {
        const int buf = it & 1;
        if (tid == 0 && jt + 1 < jt1) {
            double* dst = stage0 + (size_t)(buf ^ 1) * STAGE_DOUBLES;
            pn_mbar_expect_tx(&bars[buf ^ 1], X_BYTES + V_BYTES);
            pn_bulk_g2s(dst, x1g + (jt + 1) * MV_TJ * D, X_BYTES, &bars[buf ^ 1]);
            pn_bulk_g2s(dst + MV_TJ * D, vg + (jt + 1) * MV_TJ, V_BYTES, &bars[buf ^ 1]);
        }
        if (it > 0 && jt - 1 > 2 * bx + 1) {  // column sums of the previous tile
            const double* cwp = colw + (buf ^ 1) * (MV_THREADS / 32) * MV_TJ;
            double s = cwp[tid];
#pragma unroll
            for (int w = 1; w < MV_THREADS / 32; ++w) s += cwp[w * MV_TJ + tid];
            const int64_t col = (jt - 1) * MV_TJ + tid;
            if (col < a.n1) pn_bin_add(bhi, blo, bnf, col, s, binE);
        }
        pn_mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        const double* xt = stage0 + (size_t)buf * STAGE_DOUBLES;
        const double* vt = xt + MV_TJ * D;
        double* cw = colw + buf * (MV_THREADS / 32) * MV_TJ + warp * MV_TJ;
        const bool mirror = jt > 2 * bx + 1;  // uniform

#pragma unroll 1
        for (int jj = 0; jj < MV_TJ; jj += 4) {
            double S[4 * MV_R];  // S[4 r + c]: row r, column jj + c
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double xb[D];
#pragma unroll
                for (int m = 0; m < D; ++m) xb[m] = xt[(jj + c) * D + m];
#pragma unroll
                for (int r = 0; r < MV_R; ++r) {
                    double s = 0.0;
#pragma unroll
                    for (int m = 0; m < D; ++m) {
                        double df = xr[r][m] - xb[m];
                        if constexpr (ARD) df *= sc[m];
                        s = pn_fma(df, df, s);
                    }
                    S[4 * r + c] = s;
                }
            }
            mv_epi_tile<EPI, 4 * MV_R>(S, a.epi, tabE, tabL);
            double tc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int r = 0; r < MV_R; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[r] = pn_fma(S[4 * r + c], vt[jj + c], acc[r]);
                    tc[c] = pn_fma(S[4 * r + c], vrow[r], tc[c]);
                }
            if (mirror) {
                // halving exchange: after the first step the lower half-warp holds columns jj, jj + 1 and the upper
                // jj + 2, jj + 3; after the second every quarter holds one column; three plain steps finish the sums
                const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
                const double a0 = (up16 ? tc[2] : tc[0]) + __shfl_xor_sync(0xffffffffu, up16 ? tc[0] : tc[2], 16);
                const double a1 = (up16 ? tc[3] : tc[1]) + __shfl_xor_sync(0xffffffffu, up16 ? tc[1] : tc[3], 16);
                double t = (up8 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, up8 ? a0 : a1, 8);
#pragma unroll
                for (int m = 4; m > 0; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
                if ((lane & 7) == 0) cw[jj + (lane >> 3)] = t;
            }
        }
        __syncthreads();
    }
    if (jt1 - 1 > 2 * bx + 1) {  // column sums of the last tile
        const double* cwp = colw + ((it - 1) & 1) * (MV_THREADS / 32) * MV_TJ;
        double s = cwp[tid];
#pragma unroll
        for (int w = 1; w < MV_THREADS / 32; ++w) s += cwp[w * MV_TJ + tid];
        const int64_t col = (jt1 - 1) * MV_TJ + tid;
        if (col < a.n1) pn_bin_add(bhi, blo, bnf, col, s, binE);
    }
#pragma unroll
    for (int r = 0; r < MV_R; ++r) {
        const int64_t row = row0 + (int64_t)r * MV_THREADS;
        if (row < a.n0) pn_bin_add(bhi, blo, bnf, row, acc[r], binE);
    }
}

// ---- generic kernel: any d <= 1024, float64 or float32 points ---------------------------------
// One row per thread, its point read through L1 each time; column points in coordinate-major
// shared-memory tiles, four columns per register tile.  Slower than the main kernel (extra
// L1 traffic); used for d > 8 and for float32.
constexpr int MG_THREADS = 128;

template <typename T, bool ARD, int KC>
__global__ void __launch_bounds__(MG_THREADS)
    matvec_generic_kernel(const __grid_constant__ PnMatvecLaunch a, const T* __restrict__ x1_soa,
                          int64_t ld1, int tj, int tiles_per_split) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int d = a.d;
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    T* sc = reinterpret_cast<T*>(smem + MV_HDR_BYTES);        // [d]
    T* xt = sc + ((d + 3) / 4) * 4;                            // [d][tj]
    double* vt = reinterpret_cast<double*>(xt + (size_t)d * tj);  // [tj][KC] (8-byte aligned: d*tj*4 % 8 == 0)

    pn_stage_tables(tabE, tabL, tid, MG_THREADS);
    for (int m = tid; m < d; m += MG_THREADS) sc[m] = ARD ? static_cast<const T*>(a.scale)[m] : T(1);

    const int64_t row = (int64_t)blockIdx.x * MG_THREADS + tid;
    const T* xrow = static_cast<const T*>(a.x0_aos) + row * d;  // padded rows are readable
    const T* vg = static_cast<const T*>(a.v);

    const int64_t n_jt = (a.n1pad + tj - 1) / tj;
    const int64_t jt0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < n_jt) ? jt0 + tiles_per_split : n_jt;

    double acc[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[k] = 0.0;

    for (int64_t jt = jt0; jt < jt1; ++jt) {
        const int64_t jbase = jt * tj;
        __syncthreads();
        for (int idx = tid; idx < d * tj; idx += MG_THREADS) {
            const int m = idx / tj, j = idx % tj;
            xt[idx] = (jbase + j < a.n1pad) ? x1_soa[(size_t)m * ld1 + jbase + j] : T(0);
        }
        for (int idx = tid; idx < tj * KC; idx += MG_THREADS) {
            const int64_t j = jbase + idx / KC;
            vt[idx] = (j < a.n1pad) ? (double)vg[j * KC + idx % KC] : 0.0;
        }
        __syncthreads();
        for (int jj = 0; jj < tj; jj += 4) {
            T S[4] = {T(0), T(0), T(0), T(0)};
            for (int m = 0; m < d; ++m) {
                const T xa = __ldg(xrow + m);
                T s = T(1);
                if constexpr (ARD) s = sc[m];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    T df = xa - xt[m * tj + jj + c];
                    if constexpr (ARD) df *= s;
                    if constexpr (std::is_same<T, double>::value) S[c] = pn_fma(df, df, S[c]);
                    else S[c] = fmaf(df, df, S[c]);
                }
            }
            double Kv[4];
            if constexpr (std::is_same<T, float>::value) {
                if (a.epi.epi == PN_EPI_EXPQUAD) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) Kv[c] = (double)exp2f(S[c] * (float)a.epi.c_u);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) Kv[c] = (double)S[c];
                    mv_epi_dispatch(Kv, a.epi, tabE, tabL);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) Kv[c] = S[c];
                mv_epi_dispatch(Kv, a.epi, tabE, tabL);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int k = 0; k < KC; ++k) acc[k] = pn_fma(Kv[c], vt[(jj + c) * KC + k], acc[k]);
        }
    }
    if (row < a.n0) {
        if (a.jsplit > 1) {
            double* dst = static_cast<double*>(a.partial) + ((size_t)blockIdx.y * a.n0 + row) * KC;
#pragma unroll
            for (int k = 0; k < KC; ++k) dst[k] = acc[k];
        } else {
            T* dst = static_cast<T*>(a.y) + row * KC;
#pragma unroll
            for (int k = 0; k < KC; ++k) dst[k] = (T)acc[k];
        }
    }
}

// ---- direct-difference kernel for 8 < d <= MR_DMAX (and float32 points) ---------------------------
// The generic kernel above re-reads its row point through L1 for every tile (32 scattered sectors
// per warp load: 8x below the FP64 roofline at d = 16..64).  Here the CTA's 128 row points sit in
// shared memory coordinate-major ([d][128]: consecutive lanes, conflict-free), column points
// stream in coordinate-major tiles of MR_TJ, and a thread evaluates its row against 8 columns at
// a time: per coordinate 1 + 4 shared loads (the column values as broadcast double2) for 16 FP64
// instructions.
constexpr int MR_THREADS = 128;
constexpr int MR_TJ = 64;
constexpr int MR_C = 8;      // columns per register tile
constexpr int MR_DMAX = 96;  // row block fits 96 KB of shared memory in float64

template <typename T, bool ARD, int KC>
__global__ void __launch_bounds__(MR_THREADS)
    matvec_rows_kernel(const __grid_constant__ PnMatvecLaunch a, const T* __restrict__ x0_soa, int64_t ld0,
                       const T* __restrict__ x1_soa, int64_t ld1, int tiles_per_split) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int d = a.d;
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    T* sc = reinterpret_cast<T*>(smem + MV_HDR_BYTES);                     // [d]
    T* xr = sc + ((d + 3) / 4) * 4;                                         // [d][MR_THREADS]
    T* xt = xr + (size_t)d * MR_THREADS;                                    // [d][MR_TJ]
    double* vt = reinterpret_cast<double*>(xt + (size_t)d * MR_TJ);         // [MR_TJ][KC]

    pn_stage_tables(tabE, tabL, tid, MR_THREADS);
    for (int m = tid; m < d; m += MR_THREADS) sc[m] = ARD ? static_cast<const T*>(a.scale)[m] : T(1);
    const int64_t row = (int64_t)blockIdx.x * MR_THREADS + tid;  // padded rows are readable (zeros)
    for (int m = 0; m < d; ++m) xr[m * MR_THREADS + tid] = x0_soa[(size_t)m * ld0 + row];
    const T* vg = static_cast<const T*>(a.v);

    const int64_t n_jt = (a.n1pad + MR_TJ - 1) / MR_TJ;
    const int64_t jt0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < n_jt) ? jt0 + tiles_per_split : n_jt;
    double acc[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[k] = 0.0;

    for (int64_t jt = jt0; jt < jt1; ++jt) {
        const int64_t jbase = jt * MR_TJ;
        __syncthreads();
        for (int idx = tid; idx < d * MR_TJ; idx += MR_THREADS) {
            const int m = idx / MR_TJ, j = idx % MR_TJ;
            xt[idx] = x1_soa[(size_t)m * ld1 + jbase + j];  // n1pad is a multiple of MR_TJ
        }
        for (int idx = tid; idx < MR_TJ * KC; idx += MR_THREADS) vt[idx] = (double)vg[jbase * KC + idx];
        __syncthreads();
        for (int jj = 0; jj < MR_TJ; jj += MR_C) {
            T S[MR_C];
#pragma unroll
            for (int c = 0; c < MR_C; ++c) S[c] = T(0);
            for (int m = 0; m < d; ++m) {
                const T xa = xr[m * MR_THREADS + tid];
                T s = T(1);
                if constexpr (ARD) s = sc[m];
#pragma unroll
                for (int c = 0; c < MR_C; ++c) {
                    T df = xa - xt[m * MR_TJ + jj + c];
                    if constexpr (ARD) df *= s;
                    if constexpr (std::is_same<T, double>::value) S[c] = pn_fma(df, df, S[c]);
                    else S[c] = fmaf(df, df, S[c]);
                }
            }
            double Kv[MR_C];
            if constexpr (std::is_same<T, float>::value) {
                if (a.epi.epi == PN_EPI_EXPQUAD) {
#pragma unroll
                    for (int c = 0; c < MR_C; ++c) Kv[c] = (double)exp2f(S[c] * (float)a.epi.c_u);
                } else {
#pragma unroll
                    for (int c = 0; c < MR_C; ++c) Kv[c] = (double)S[c];
                    mv_epi_dispatch(Kv, a.epi, tabE, tabL);
                }
            } else {
#pragma unroll
                for (int c = 0; c < MR_C; ++c) Kv[c] = S[c];
                mv_epi_dispatch(Kv, a.epi, tabE, tabL);
            }
#pragma unroll
            for (int c = 0; c < MR_C; ++c)
#pragma unroll
                for (int k = 0; k < KC; ++k) acc[k] = pn_fma(Kv[c], vt[(jj + c) * KC + k], acc[k]);
        }
    }
    if (row < a.n0) {
        if (a.jsplit > 1) {
            double* dst = static_cast<double*>(a.partial) + ((size_t)blockIdx.y * a.n0 + row) * KC;
#pragma unroll
            for (int k = 0; k < KC; ++k) dst[k] = acc[k];
        } else {
            T* dst = static_cast<T*>(a.y) + row * KC;
#pragma unroll
            for (int k = 0; k < KC; ++k) dst[k] = (T)acc[k];
        }
    }
}

// ---- symmetric variant of the row-staged kernel (x1 is x0, one right-hand side, float64) -------------
// Same halving as matvec_sym_kernel for 8 < d <= MR_DMAX: the CTA of row block I (128 rows = two
// column tiles) walks the column tiles J >= 2 I and adds y_J += K_IJ^T v_I for the tiles right of its
// diagonal pair.  The eight partial column sums of a register tile are folded over the warp by a
// butterfly (8 -> 4 -> 2 -> 1 values over lane bits 4, 3, 2, then two plain rounds: 9 shuffles +
// 9 DADDs per 8 columns), warps meet in shared memory, one integer-plane update per column and tile.
template <bool ARD>
__global__ void __launch_bounds__(MR_THREADS)
    matvec_rows_sym_kernel(const __grid_constant__ PnMatvecLaunch a, const double* __restrict__ x_soa, int64_t ld,
                           int tiles_per_split, int part, int nparts) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = a.d;
    constexpr int TPB = MR_THREADS / MR_TJ;  // column tiles per row block
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    double* sc = reinterpret_cast<double*>(smem + MV_HDR_BYTES);     // [d]
    double* xr = sc + ((d + 3) / 4) * 4;                               // [d][MR_THREADS]
    double* xt = xr + (size_t)d * MR_THREADS;                          // [d][MR_TJ]
    double* vt = xt + (size_t)d * MR_TJ;                               // [MR_TJ]
    double* colw = vt + MR_TJ;                                         // [2][warps][MR_TJ]

    const int64_t n_jt = a.n1pad / MR_TJ;
    const int64_t bx = (int64_t)blockIdx.x * nparts + part;
    if (bx * TPB >= n_jt) return;
    const int64_t jdiag_end = (bx + 1) * TPB;
    int64_t jt0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < n_jt) ? jt0 + tiles_per_split : n_jt;
    if (jt0 < bx * TPB) jt0 = bx * TPB;
    if (jt0 >= jt1) return;

    pn_stage_tables(tabE, tabL, tid, MR_THREADS);
    for (int m = tid; m < d; m += MR_THREADS) sc[m] = ARD ? static_cast<const double*>(a.scale)[m] : 1.0;
    const int64_t row = bx * MR_THREADS + tid;  // padded rows are readable (zeros)
    for (int m = 0; m < d; ++m) xr[m * MR_THREADS + tid] = x_soa[(size_t)m * ld + row];
    const double* vg = static_cast<const double*>(a.v);
    long long* const bhi = a.bins;
    long long* const blo = a.bins + a.n0;
    double* const bnf = reinterpret_cast<double*>(a.bins + 2 * a.n0);
    const int binE = pn_bin_exponent((double)a.n1 * *a.vmax, 1.0);  // every kernel value is <= 1
    const double vrow = vg[row];  // zero in the padding
    double acc = 0.0;

    int it = 0;
    for (int64_t jt = jt0; jt < jt1; ++jt, ++it) {
        const int buf = it & 1;
        const int64_t jbase = jt * MR_TJ;
        __syncthreads();  // previous tile: xt / vt consumed, its colw slot complete
        if (it > 0 && jt - 1 >= jdiag_end && tid < MR_TJ) {
            const double* cwp = colw + (buf ^ 1) * (MR_THREADS / 32) * MR_TJ;
            double s = cwp[tid];
#pragma unroll
            for (int w = 1; w < MR_THREADS / 32; ++w) s += cwp[w * MR_TJ + tid];
            const int64_t col = (jt - 1) * MR_TJ + tid;
            if (col < a.n1) pn_bin_add(bhi, blo, bnf, col, s, binE);
        }
        for (int idx = tid; idx < d * MR_TJ; idx += MR_THREADS) {
            const int m = idx / MR_TJ, j = idx % MR_TJ;
            xt[idx] = x_soa[(size_t)m * ld + jbase + j];
        }
        if (tid < MR_TJ) vt[tid] = vg[jbase + tid];
        __syncthreads();
        double* cw = colw + buf * (MR_THREADS / 32) * MR_TJ + warp * MR_TJ;
        const bool mirror = jt >= jdiag_end;  // uniform
        for (int jj = 0; jj < MR_TJ; jj += MR_C) {
            double S[MR_C];
#pragma unroll
            for (int c = 0; c < MR_C; ++c) S[c] = 0.0;
            for (int m = 0; m < d; ++m) {
                const double xa = xr[m * MR_THREADS + tid];
                double s = 1.0;
                if constexpr (ARD) s = sc[m];
#pragma unroll
                for (int c = 0; c < MR_C; ++c) {
                    double df = xa - xt[m * MR_TJ + jj + c];
                    if constexpr (ARD) df *= s;
                    S[c] = pn_fma(df, df, S[c]);
                }
            }
            mv_epi_dispatch(S, a.epi, tabE, tabL);
#pragma unroll
            for (int c = 0; c < MR_C; ++c) acc = pn_fma(S[c], vt[jj + c], acc);
            if (mirror) {
                static_assert(MR_C == 8, "the butterfly folds exactly eight columns");
                const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
                double t[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {  // keep columns 4 b4 + c, send the other half
                    const double keep = (b4 ? S[4 + c] : S[c]) * vrow, send = (b4 ? S[c] : S[4 + c]) * vrow;
                    t[c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
                double u0 = (b3 ? t[2] : t[0]) + __shfl_xor_sync(0xffffffffu, b3 ? t[0] : t[2], 8);
                double u1 = (b3 ? t[3] : t[1]) + __shfl_xor_sync(0xffffffffu, b3 ? t[1] : t[3], 8);
                double w = (b2 ? u1 : u0) + __shfl_xor_sync(0xffffffffu, b2 ? u0 : u1, 4);
                w += __shfl_xor_sync(0xffffffffu, w, 2);
                w += __shfl_xor_sync(0xffffffffu, w, 1);
                if ((lane & 3) == 0) cw[jj + (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0)] = w;
            }
        }
    }
    __syncthreads();
    if (jt1 - 1 >= jdiag_end && tid < MR_TJ) {  // column sums of the last tile
        const double* cwp = colw + ((it - 1) & 1) * (MR_THREADS / 32) * MR_TJ;
        double s = cwp[tid];
#pragma unroll
        for (int w = 1; w < MR_THREADS / 32; ++w) s += cwp[w * MR_TJ + tid];
        const int64_t col = (jt1 - 1) * MR_TJ + tid;
        if (col < a.n1) pn_bin_add(bhi, blo, bnf, col, s, binE);
    }
    if (row < a.n0) pn_bin_add(bhi, blo, bnf, row, acc, binE);
}

// max |v_j| as the bit pattern of a non-negative double (unsigned order = numeric order; NaN sorts above Inf)
__global__ void vmax_kernel(const double* __restrict__ v, int64_t ldv, int64_t n, int k, unsigned long long* __restrict__ out) {
    unsigned long long m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * k; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v[(i / k) * ldv + i % k]));
        m = b > m ? b : m;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// integer planes -> y (the same exponents the kernels used: pn_bin_exponent of the same inputs)
__global__ void unbin_kernel(const long long* __restrict__ bins, int64_t n0, int kc, int64_t n1, const double* __restrict__ vmax,
                             const double* __restrict__ fac, int fac_ld, int fac_idx, double* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // row * kc + column
    const int64_t plane = n0 * kc;
    if (i >= plane) return;
    const int E = pn_bin_exponent((double)n1 * *vmax, fac ? fac[(i / kc) * fac_ld + fac_idx] : 1.0);
    y[i] = pn_bin_value(bins[i], bins[plane + i], reinterpret_cast<const double*>(bins + 2 * plane)[i], E);
}

// partial[s][i] summed over s in increasing order (deterministic)
template <typename T>
__global__ void reduce_partials_kernel(const double* __restrict__ partial, T* __restrict__ y,
                                       int64_t n, int jsplit) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = partial[i];
    for (int k = 1; k < jsplit; ++k) s += partial[(size_t)k * n + i];
    y[i] = (T)s;
}

// Column splits: grid = row CTAs x jsplit.  Pick the split that fills whole waves of the
// machine (slots = SMs x resident CTAs of this instantiation); more, shorter CTAs also
// smooth the tail.  Each split keeps at least two column tiles.
static int choose_jsplit(int64_t row_ctas, int64_t n_jt, int slots) {
    int best = 1;
    double best_eff = -1.0;
    const int64_t max_split = n_jt / 2 < 1 ? 1 : (n_jt / 2 > 32 ? 32 : n_jt / 2);
    for (int js = 1; js <= max_split; ++js) {
        const int64_t total = row_ctas * js;
        const int64_t waves = (total + slots - 1) / slots;
        double eff = (double)total / (double)(waves * slots);
        if (waves >= 8) eff = 1.0;  // long queues balance dynamically; stop splitting further
        if (eff > best_eff + 1e-3) {
            best_eff = eff;
            best = js;
        }
    }
    return best;
}

template <int D, bool ARD, int KC, int EPI, bool IDENT = false>
cudaError_t launch_main(PnMatvecLaunch& a, int sm_count, bool plan_only, cudaStream_t stream) {
    auto kern = matvec_kernel<D, ARD, KC, EPI, IDENT>;
    const size_t smem = MV_HDR_BYTES + 2 * sizeof(double) * MV_TJ * ((IDENT ? MV_LD12 : D) + KC);
    const int64_t n_jt = a.n1pad / MV_TJ;
    const int64_t row_ctas = (a.n0 + MV_THREADS * MV_R - 1) / (MV_THREADS * MV_R);
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    if (plan_only) {
        int occ = 1;
        cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MV_THREADS, smem);
        if (err != cudaSuccess) return err;
        a.jsplit = choose_jsplit(row_ctas, n_jt, sm_count * (occ < 1 ? 1 : occ));
        return cudaSuccess;
    }
    const int64_t tps = (n_jt + a.jsplit - 1) / a.jsplit;
    dim3 grid((unsigned)row_ctas, (unsigned)a.jsplit);
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    kern<<<grid, MV_THREADS, smem, stream>>>(a, (int)tps);
    ++g_pn_launches;
    return cudaGetLastError();
}

template <int D, bool ARD, int KC>
cudaError_t launch_main_epi(PnMatvecLaunch& a, int sm, bool plan, cudaStream_t stream) {
    switch (a.epi.epi) {
        case PN_EPI_EXPQUAD: return launch_main<D, ARD, KC, PN_EPI_EXPQUAD>(a, sm, plan, stream);
        case PN_EPI_MATERN0: return launch_main<D, ARD, KC, PN_EPI_MATERN0>(a, sm, plan, stream);
        case PN_EPI_MATERN1: return launch_main<D, ARD, KC, PN_EPI_MATERN1>(a, sm, plan, stream);
        case PN_EPI_MATERN2: return launch_main<D, ARD, KC, PN_EPI_MATERN2>(a, sm, plan, stream);
        case PN_EPI_MATERNP: return launch_main<D, ARD, KC, PN_EPI_MATERNP>(a, sm, plan, stream);
        case PN_EPI_RATQUAD1: return launch_main<D, false, KC, PN_EPI_RATQUAD1>(a, sm, plan, stream);
        case PN_EPI_RATQUAD2: return launch_main<D, false, KC, PN_EPI_RATQUAD2>(a, sm, plan, stream);
        default: return launch_main<D, false, KC, PN_EPI_RATQUADG>(a, sm, plan, stream);
    }
}

template <int D, bool ARD>
cudaError_t launch_main_kc(PnMatvecLaunch& a, int sm, bool plan, cudaStream_t stream) {
    switch (a.kc) {
        case 1: return launch_main_epi<D, ARD, 1>(a, sm, plan, stream);
        case 4: return launch_main_epi<D, ARD, 4>(a, sm, plan, stream);
        default: return cudaErrorInvalidValue;
    }
}

// norm-identity (DFMA) instantiations: ExpQuad / RatQuad, scale folded into the operands
template <int D, int KC>
cudaError_t launch_ident_epi(PnMatvecLaunch& a, int sm, bool plan, cudaStream_t stream) {
    switch (a.epi.epi) {
        case PN_EPI_EXPQUAD: return launch_main<D, false, KC, PN_EPI_EXPQUAD, true>(a, sm, plan, stream);
        case PN_EPI_RATQUAD1: return launch_main<D, false, KC, PN_EPI_RATQUAD1, true>(a, sm, plan, stream);
        case PN_EPI_RATQUAD2: return launch_main<D, false, KC, PN_EPI_RATQUAD2, true>(a, sm, plan, stream);
        case PN_EPI_RATQUADG: return launch_main<D, false, KC, PN_EPI_RATQUADG, true>(a, sm, plan, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int D>
cudaError_t launch_ident_kc(PnMatvecLaunch& a, int sm, bool plan, cudaStream_t stream) {
    if (a.kc == 1) return launch_ident_epi<D, 1>(a, sm, plan, stream);
    if (a.kc == 4) return launch_ident_epi<D, 4>(a, sm, plan, stream);
    return cudaErrorInvalidValue;
}

cudaError_t launch_ident_d(PnMatvecLaunch& a, int sm, bool plan, cudaStream_t stream) {
    switch (a.d) {
        case 1: return launch_ident_kc<1>(a, sm, plan, stream);
        case 2: return launch_ident_kc<2>(a, sm, plan, stream);
        case 3: return launch_ident_kc<3>(a, sm, plan, stream);
        case 4: return launch_ident_kc<4>(a, sm, plan, stream);
        case 5: return launch_ident_kc<5>(a, sm, plan, stream);
        case 6: return launch_ident_kc<6>(a, sm, plan, stream);
        case 7: return launch_ident_kc<7>(a, sm, plan, stream);
        case 8: return launch_ident_kc<8>(a, sm, plan, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <bool ARD>
cudaError_t launch_main_d(PnMatvecLaunch& a, int sm, bool plan, cudaStream_t stream) {
    switch (a.d) {
        case 1: return launch_main_kc<1, ARD>(a, sm, plan, stream);
        case 2: return launch_main_kc<2, ARD>(a, sm, plan, stream);
        case 3: return launch_main_kc<3, ARD>(a, sm, plan, stream);
        case 4: return launch_main_kc<4, ARD>(a, sm, plan, stream);
        case 5: return launch_main_kc<5, ARD>(a, sm, plan, stream);
        case 6: return launch_main_kc<6, ARD>(a, sm, plan, stream);
        case 7: return launch_main_kc<7, ARD>(a, sm, plan, stream);
        case 8: return launch_main_kc<8, ARD>(a, sm, plan, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int KSTEPS, int KC, int EPI>
cudaError_t launch_mma(PnMatvecLaunch& a, int sm_count, bool plan_only, cudaStream_t stream) {
    using Cfg = MmCfg<KSTEPS>;
    auto kern = matvec_mma_kernel<KSTEPS, KC, EPI>;
    const size_t smem = MV_HDR_BYTES + 2 * sizeof(double) * Cfg::TJ * (Cfg::LD + (KC == 8 ? 10 : KC));
    const int64_t n_jt = a.n1pad / Cfg::TJ;
    const int64_t row_ctas = (a.n0 + MM_ROWS - 1) / MM_ROWS;
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    if (plan_only) {
        int occ = 1;
        cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MM_THREADS, smem);
        if (err != cudaSuccess) return err;
        a.jsplit = choose_jsplit(row_ctas, n_jt, sm_count * (occ < 1 ? 1 : occ));
        return cudaSuccess;
    }
    const int64_t tps = (n_jt + a.jsplit - 1) / a.jsplit;
    dim3 grid((unsigned)row_ctas, (unsigned)a.jsplit);
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    kern<<<grid, MM_THREADS, smem, stream>>>(a, (int)tps);
    ++g_pn_launches;
    return cudaGetLastError();
}

template <int KSTEPS, int KC>
cudaError_t launch_mma_epi(PnMatvecLaunch& a, int sm, bool plan, cudaStream_t stream) {
    switch (a.epi.epi) {
        case PN_EPI_EXPQUAD:
            if (a.fact) return launch_mma<KSTEPS, KC, PN_EPI_EXPQUAD_FACT>(a, sm, plan, stream);
            return launch_mma<KSTEPS, KC, PN_EPI_EXPQUAD>(a, sm, plan, stream);
        case PN_EPI_RATQUAD1: return launch_mma<KSTEPS, KC, PN_EPI_RATQUAD1>(a, sm, plan, stream);
        case PN_EPI_RATQUAD2: return launch_mma<KSTEPS, KC, PN_EPI_RATQUAD2>(a, sm, plan, stream);
        case PN_EPI_RATQUADG: return launch_mma<KSTEPS, KC, PN_EPI_RATQUADG>(a, sm, plan, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int D, bool ARD, int EPI>
cudaError_t launch_direct_sym(const PnMatvecLaunch& a, cudaStream_t stream) {
    auto kern = matvec_sym_kernel<D, ARD, EPI>;
    const size_t smem = MV_HDR_BYTES + 2 * sizeof(double) * MV_TJ * (D + 1) + 16 +
                        2 * sizeof(double) * (MV_THREADS / 32) * MV_TJ;
    const int64_t n_jt = a.n1pad / MV_TJ;
    const int64_t n_blocks = (n_jt + 1) / 2;
    const int nparts = a.sym_nparts > 0 ? a.sym_nparts : 1;
    const int64_t row_ctas = (n_blocks + nparts - 1) / nparts;
    int64_t tps = 256;
    while (tps > 8 && row_ctas * ((n_jt + tps - 1) / tps) < (int64_t)pn_sm_count() * 4 * 8) tps /= 2;
    const int64_t chunks = (n_jt + tps - 1) / tps;
    if (chunks > 65535) return cudaErrorInvalidConfiguration;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (!a.bins || !a.vmax) return cudaErrorInvalidValue;  // zeroed integer planes + max |v| come from the caller
    dim3 grid((unsigned)row_ctas, (unsigned)chunks);
    kern<<<grid, MV_THREADS, smem, stream>>>(a, (int)tps, a.sym_part, nparts);
    ++g_pn_launches;
    return cudaGetLastError();
}

template <int D, bool ARD>
cudaError_t launch_direct_sym_epi(const PnMatvecLaunch& a, cudaStream_t stream) {
    switch (a.epi.epi) {
        case PN_EPI_EXPQUAD: return launch_direct_sym<D, ARD, PN_EPI_EXPQUAD>(a, stream);
        case PN_EPI_MATERN0: return launch_direct_sym<D, ARD, PN_EPI_MATERN0>(a, stream);
        case PN_EPI_MATERN1: return launch_direct_sym<D, ARD, PN_EPI_MATERN1>(a, stream);
        case PN_EPI_MATERN2: return launch_direct_sym<D, ARD, PN_EPI_MATERN2>(a, stream);
        case PN_EPI_MATERNP: return launch_direct_sym<D, ARD, PN_EPI_MATERNP>(a, stream);
        case PN_EPI_RATQUAD1: return launch_direct_sym<D, false, PN_EPI_RATQUAD1>(a, stream);
        case PN_EPI_RATQUAD2: return launch_direct_sym<D, false, PN_EPI_RATQUAD2>(a, stream);
        default: return launch_direct_sym<D, false, PN_EPI_RATQUADG>(a, stream);
    }
}

template <bool ARD>
cudaError_t launch_direct_sym_d(const PnMatvecLaunch& a, cudaStream_t stream) {
    switch (a.d) {
        case 1: return launch_direct_sym_epi<1, ARD>(a, stream);
        case 2: return launch_direct_sym_epi<2, ARD>(a, stream);
        case 3: return launch_direct_sym_epi<3, ARD>(a, stream);
        case 4: return launch_direct_sym_epi<4, ARD>(a, stream);
        case 5: return launch_direct_sym_epi<5, ARD>(a, stream);
        case 6: return launch_direct_sym_epi<6, ARD>(a, stream);
        case 7: return launch_direct_sym_epi<7, ARD>(a, stream);
        case 8: return launch_direct_sym_epi<8, ARD>(a, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int KSTEPS, int EPI>
cudaError_t launch_mma_sym(const PnMatvecLaunch& a, cudaStream_t stream) {
    using Cfg = MmCfg<KSTEPS>;
    constexpr int TPB = MM_ROWS / Cfg::TJ;
    auto kern = matvec_mma_sym_kernel<KSTEPS, EPI>;
    const size_t smem = MV_HDR_BYTES + 2 * sizeof(double) * Cfg::TJ * (Cfg::LD + 1) +
                        2 * sizeof(double) * (MM_THREADS / 32) * Cfg::TJ;
    const int64_t n_jt = a.n1pad / Cfg::TJ;
    const int64_t n_blocks = (n_jt + TPB - 1) / TPB;
    const int nparts = a.sym_nparts > 0 ? a.sym_nparts : 1;
    const int64_t row_ctas = (n_blocks + nparts - 1) / nparts;
    // chunks of column tiles (in units of 128 columns): short enough to balance the triangle, long
    // enough to amortise the prologue
    int64_t tps = 256;
    while (tps > 8 && row_ctas * ((n_blocks + tps - 1) / tps) < (int64_t)pn_sm_count() * 4 * 8) tps /= 2;
    tps *= TPB;
    const int64_t chunks = (n_jt + tps - 1) / tps;
    if (chunks > 65535) return cudaErrorInvalidConfiguration;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (!a.bins || !a.vmax) return cudaErrorInvalidValue;  // zeroed integer planes + max |v| come from the caller
    dim3 grid((unsigned)row_ctas, (unsigned)chunks);
    kern<<<grid, MM_THREADS, smem, stream>>>(a, (int)tps, a.sym_part, nparts);
    ++g_pn_launches;
    return cudaGetLastError();
}

template <int KSTEPS, int EPI>
cudaError_t launch_mma_symd(const PnMatvecLaunch& a, cudaStream_t stream) {
    using Cfg = MmCfg<KSTEPS>;
    auto kern = matvec_mma_symd_kernel<KSTEPS, EPI>;
    constexpr int TPB = MM_ROWS / MD_TJ;
    const size_t smem = MV_HDR_BYTES + sizeof(double) * (2 * MD_TJ * (Cfg::LD + MD_KCP) + (MM_THREADS / 32) * MD_TJ * MD_KC +
                                                         (MM_THREADS / 32) * MM_RG * 64);
    const int64_t n_jt = a.n1pad / MD_TJ;
    const int64_t n_blocks = (n_jt + TPB - 1) / TPB;
    const int nparts = a.sym_nparts > 0 ? a.sym_nparts : 1;
    const int64_t row_ctas = (n_blocks + nparts - 1) / nparts;
    int64_t tps = 256;
    while (tps > 8 && row_ctas * ((n_blocks + tps - 1) / tps) < (int64_t)pn_sm_count() * 3 * 8) tps /= 2;
    tps *= TPB;
    const int64_t chunks = (n_jt + tps - 1) / tps;
    if (chunks > 65535) return cudaErrorInvalidConfiguration;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (!a.bins || !a.vmax) return cudaErrorInvalidValue;
    dim3 grid((unsigned)row_ctas, (unsigned)chunks);
    kern<<<grid, MM_THREADS, smem, stream>>>(a, (int)tps, a.sym_part, nparts);
    ++g_pn_launches;
    return cudaGetLastError();
}

template <int KSTEPS>
cudaError_t launch_mma_symd_epi(const PnMatvecLaunch& a, cudaStream_t stream) {
    switch (a.epi.epi) {
        case PN_EPI_EXPQUAD:
            if (a.fact) return launch_mma_symd<KSTEPS, PN_EPI_EXPQUAD_FACT>(a, stream);
            return launch_mma_symd<KSTEPS, PN_EPI_EXPQUAD>(a, stream);
        case PN_EPI_RATQUAD1: return launch_mma_symd<KSTEPS, PN_EPI_RATQUAD1>(a, stream);
        case PN_EPI_RATQUAD2: return launch_mma_symd<KSTEPS, PN_EPI_RATQUAD2>(a, stream);
        case PN_EPI_RATQUADG: return launch_mma_symd<KSTEPS, PN_EPI_RATQUADG>(a, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int KSTEPS>
cudaError_t launch_mma_sym_epi(const PnMatvecLaunch& a, cudaStream_t stream) {
    switch (a.epi.epi) {
        case PN_EPI_EXPQUAD:
            if (a.fact) return launch_mma_sym<KSTEPS, PN_EPI_EXPQUAD_FACT>(a, stream);
            return launch_mma_sym<KSTEPS, PN_EPI_EXPQUAD>(a, stream);
        case PN_EPI_RATQUAD1: return launch_mma_sym<KSTEPS, PN_EPI_RATQUAD1>(a, stream);
        case PN_EPI_RATQUAD2: return launch_mma_sym<KSTEPS, PN_EPI_RATQUAD2>(a, stream);
        case PN_EPI_RATQUADG: return launch_mma_sym<KSTEPS, PN_EPI_RATQUADG>(a, stream);
        default: return cudaErrorInvalidValue;
    }
}

// 8 < d <= 64 (KSTEPS 4, 8, 16): one right-hand side, ExpQuad in factorised form or RatQuad
template <int KSTEPS>
cudaError_t launch_mma_wide(PnMatvecLaunch& a, int sm, bool plan, cudaStream_t stream) {
    switch (a.epi.epi) {
        case PN_EPI_EXPQUAD: return launch_mma<KSTEPS, 1, PN_EPI_EXPQUAD_FACT>(a, sm, plan, stream);
        case PN_EPI_RATQUAD1: return launch_mma<KSTEPS, 1, PN_EPI_RATQUAD1>(a, sm, plan, stream);
        case PN_EPI_RATQUAD2: return launch_mma<KSTEPS, 1, PN_EPI_RATQUAD2>(a, sm, plan, stream);
        case PN_EPI_RATQUADG: return launch_mma<KSTEPS, 1, PN_EPI_RATQUADG>(a, sm, plan, stream);
        default: return cudaErrorInvalidValue;
    }
}
template <int KSTEPS>
cudaError_t launch_mma_sym_wide(const PnMatvecLaunch& a, cudaStream_t stream) {
    switch (a.epi.epi) {
        case PN_EPI_EXPQUAD: return launch_mma_sym<KSTEPS, PN_EPI_EXPQUAD_FACT>(a, stream);
        case PN_EPI_RATQUAD1: return launch_mma_sym<KSTEPS, PN_EPI_RATQUAD1>(a, stream);
        case PN_EPI_RATQUAD2: return launch_mma_sym<KSTEPS, PN_EPI_RATQUAD2>(a, stream);
        case PN_EPI_RATQUADG: return launch_mma_sym<KSTEPS, PN_EPI_RATQUADG>(a, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_mma_any(PnMatvecLaunch& a, int sm, bool plan, cudaStream_t stream) {
    if (a.d > 8) {  // the caller admits this only for kc == 1 (and ExpQuad only in factorised form)
        if (a.kc != 1) return cudaErrorInvalidValue;
        if (a.d <= 16) return launch_mma_wide<4>(a, sm, plan, stream);
        return a.d <= 32 ? launch_mma_wide<8>(a, sm, plan, stream) : launch_mma_wide<16>(a, sm, plan, stream);
    }
    const bool two = a.d > 4;
    if (a.kc == 1) return two ? launch_mma_epi<2, 1>(a, sm, plan, stream) : launch_mma_epi<1, 1>(a, sm, plan, stream);
    if (a.kc == 4) return two ? launch_mma_epi<2, 4>(a, sm, plan, stream) : launch_mma_epi<1, 4>(a, sm, plan, stream);
    // eight right-hand sides per pass: the kernel values (17 of the 25 FP64-pipe slots per pair) are evaluated once
    // for all eight columns instead of once per group of four
    if (a.kc == 8) return two ? launch_mma_epi<2, 8>(a, sm, plan, stream) : launch_mma_epi<1, 8>(a, sm, plan, stream);
    return cudaErrorInvalidValue;
}

int generic_tj(int d, size_t elem) {
    // coordinate-major tile of tj points: keep it under ~40 KB
    int tj = (int)((40 * 1024) / ((size_t)d * elem));
    tj = (tj / 4) * 4;
    if (tj > 128) tj = 128;
    return tj;
}

template <typename T, bool ARD, int KC>
cudaError_t launch_generic(const PnMatvecLaunch& a, const void* x1_soa, int64_t ld1, cudaStream_t stream) {
    const int tj = generic_tj(a.d, sizeof(T));
    if (tj < 4) return cudaErrorInvalidValue;
    const int64_t n_jt = (a.n1pad + tj - 1) / tj;
    const int64_t tps = (n_jt + a.jsplit - 1) / a.jsplit;
    const size_t smem = MV_HDR_BYTES + sizeof(T) * (((a.d + 3) / 4) * 4 + (size_t)a.d * tj) + 8 * (size_t)tj * KC + 16;
    auto kern = matvec_generic_kernel<T, ARD, KC>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    dim3 grid((unsigned)((a.n0 + MG_THREADS - 1) / MG_THREADS), (unsigned)a.jsplit);
    kern<<<grid, MG_THREADS, smem, stream>>>(a, static_cast<const T*>(x1_soa), ld1, tj, (int)tps);
    ++g_pn_launches;
    return cudaGetLastError();
}

template <typename T, bool ARD, int KC>
cudaError_t launch_rows(const PnMatvecLaunch& a, cudaStream_t stream) {
    const int64_t n_jt = (a.n1pad + MR_TJ - 1) / MR_TJ;
    const int64_t tps = (n_jt + a.jsplit - 1) / a.jsplit;
    const size_t smem = MV_HDR_BYTES + sizeof(T) * (((a.d + 3) / 4) * 4 + (size_t)a.d * (MR_THREADS + MR_TJ)) +
                        8 * (size_t)MR_TJ * KC + 16;
    auto kern = matvec_rows_kernel<T, ARD, KC>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    dim3 grid((unsigned)((a.n0 + MR_THREADS - 1) / MR_THREADS), (unsigned)a.jsplit);
    kern<<<grid, MR_THREADS, smem, stream>>>(a, static_cast<const T*>(a.x0_soa), a.ld0, static_cast<const T*>(a.x1_soa),
                                             a.ld1, (int)tps);
    ++g_pn_launches;
    return cudaGetLastError();
}

template <bool ARD>
cudaError_t launch_rows_sym(const PnMatvecLaunch& a, cudaStream_t stream) {
    constexpr int TPB = MR_THREADS / MR_TJ;
    const int64_t n_jt = a.n1pad / MR_TJ;
    const int64_t n_blocks = (n_jt + TPB - 1) / TPB;
    const int nparts = a.sym_nparts > 0 ? a.sym_nparts : 1;
    const int64_t row_ctas = (n_blocks + nparts - 1) / nparts;
    int64_t tps = 256;
    while (tps > 8 && row_ctas * ((n_blocks + tps - 1) / tps) < (int64_t)pn_sm_count() * 4 * 8) tps /= 2;
    tps *= TPB;
    const int64_t chunks = (n_jt + tps - 1) / tps;
    if (chunks > 65535) return cudaErrorInvalidConfiguration;
    const size_t smem = MV_HDR_BYTES + sizeof(double) * (((a.d + 3) / 4) * 4 + (size_t)a.d * (MR_THREADS + MR_TJ) + MR_TJ +
                                                         2 * (MR_THREADS / 32) * MR_TJ) + 16;
    auto kern = matvec_rows_sym_kernel<ARD>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (!a.bins || !a.vmax) return cudaErrorInvalidValue;  // zeroed integer planes + max |v| come from the caller
    dim3 grid((unsigned)row_ctas, (unsigned)chunks);
    kern<<<grid, MR_THREADS, smem, stream>>>(a, static_cast<const double*>(a.x0_soa), a.ld0, (int)tps, a.sym_part, nparts);
    ++g_pn_launches;
    return cudaGetLastError();
}

}  // namespace

static bool uses_main_kernel(const PnMatvecLaunch& a) { return !a.in_f32 && a.d <= 8; }
static bool uses_rows_kernel(const PnMatvecLaunch& a) { return a.d <= MR_DMAX && a.x0_soa != nullptr; }

// fills a.jsplit for this shape (scratch sizing happens in the caller)
cudaError_t pn_plan_matvec(PnMatvecLaunch& a, int sm_count) {
    if (a.n0 <= 0 || a.n1 <= 0) {
        a.jsplit = 1;
        return cudaSuccess;
    }
    if (a.sym && (a.use_mma == 2 || (a.use_mma == 0 && !a.in_f32 && a.d <= MR_DMAX))) {
        a.jsplit = 1;  // no partial buffers: the symmetric kernels accumulate in the integer planes
        return cudaSuccess;
    }
    if (a.use_mma == 2) return launch_mma_any(a, sm_count, true, nullptr);
    if (a.use_mma == 1) return launch_ident_d(a, sm_count, true, nullptr);
    if (uses_main_kernel(a)) {
        const bool ard = a.scale != nullptr;
        return ard ? launch_main_d<true>(a, sm_count, true, nullptr) : launch_main_d<false>(a, sm_count, true, nullptr);
    }
    if (uses_rows_kernel(a)) {
        const int64_t row_ctas = (a.n0 + MR_THREADS - 1) / MR_THREADS;
        a.jsplit = choose_jsplit(row_ctas, (a.n1pad + MR_TJ - 1) / MR_TJ, sm_count * 2);
        return cudaSuccess;
    }
    const int tj = generic_tj(a.d, a.in_f32 ? 4 : 8);
    if (tj < 4) return cudaErrorInvalidValue;
    const int64_t row_ctas = (a.n0 + MG_THREADS - 1) / MG_THREADS;
    a.jsplit = choose_jsplit(row_ctas, (a.n1pad + tj - 1) / tj, sm_count * 6);
    return cudaSuccess;
}

cudaError_t pn_launch_matvec(const PnMatvecLaunch& a_in, int sm_count, cudaStream_t stream) {
    PnMatvecLaunch a = a_in;
    if (a.n0 <= 0) return cudaSuccess;
    const void* x1_soa = a.x1_soa;
    const int64_t ld1 = a.ld1;
    const bool ard = a.scale != nullptr;
    cudaError_t err;
    if (a.use_mma == 2 && a.sym && a.kc > 1) {  // several right-hand sides: d <= 8 (checked by the caller), passes of eight
        if (a.kc != MD_KC) return cudaErrorInvalidValue;
        return a.d > 4 ? launch_mma_symd_epi<2>(a, stream) : launch_mma_symd_epi<1>(a, stream);
    }
    if (a.use_mma == 2 && a.sym) {
        if (a.d > 32) return launch_mma_sym_wide<16>(a, stream);
        if (a.d > 8) return a.d <= 16 ? launch_mma_sym_wide<4>(a, stream) : launch_mma_sym_wide<8>(a, stream);
        return a.d > 4 ? launch_mma_sym_epi<2>(a, stream) : launch_mma_sym_epi<1>(a, stream);
    } else if (a.sym && a.use_mma == 0 && uses_main_kernel(a)) {
        return ard ? launch_direct_sym_d<true>(a, stream) : launch_direct_sym_d<false>(a, stream);
    } else if (a.sym && a.use_mma == 0 && !a.in_f32 && uses_rows_kernel(a)) {
        return ard ? launch_rows_sym<true>(a, stream) : launch_rows_sym<false>(a, stream);
    } else if (a.use_mma == 2) {
        err = launch_mma_any(a, sm_count, false, stream);
    } else if (a.use_mma == 1) {
        err = launch_ident_d(a, sm_count, false, stream);
    } else if (uses_main_kernel(a)) {
        err = ard ? launch_main_d<true>(a, sm_count, false, stream) : launch_main_d<false>(a, sm_count, false, stream);
    } else if (uses_rows_kernel(a)) {
        if (!a.in_f32) {
            if (a.kc == 1) err = ard ? launch_rows<double, true, 1>(a, stream) : launch_rows<double, false, 1>(a, stream);
            else err = ard ? launch_rows<double, true, 4>(a, stream) : launch_rows<double, false, 4>(a, stream);
        } else {
            if (a.kc == 1) err = ard ? launch_rows<float, true, 1>(a, stream) : launch_rows<float, false, 1>(a, stream);
            else err = ard ? launch_rows<float, true, 4>(a, stream) : launch_rows<float, false, 4>(a, stream);
        }
    } else if (!a.in_f32) {
        if (a.kc == 1) err = ard ? launch_generic<double, true, 1>(a, x1_soa, ld1, stream) : launch_generic<double, false, 1>(a, x1_soa, ld1, stream);
        else err = ard ? launch_generic<double, true, 4>(a, x1_soa, ld1, stream) : launch_generic<double, false, 4>(a, x1_soa, ld1, stream);
    } else {
        if (a.kc == 1) err = ard ? launch_generic<float, true, 1>(a, x1_soa, ld1, stream) : launch_generic<float, false, 1>(a, x1_soa, ld1, stream);
        else err = ard ? launch_generic<float, true, 4>(a, x1_soa, ld1, stream) : launch_generic<float, false, 4>(a, x1_soa, ld1, stream);
    }
    if (err != cudaSuccess) return err;
    if (a.jsplit > 1) {
        const int64_t n = a.n0 * a.kc;
        const unsigned blocks = (unsigned)((n + 255) / 256);
        if (a.in_f32)
            reduce_partials_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const double*>(a.partial), static_cast<float*>(a.y), n, a.jsplit);
        else
            reduce_partials_kernel<double><<<blocks, 256, 0, stream>>>(static_cast<const double*>(a.partial), static_cast<double*>(a.y), n, a.jsplit);
        ++g_pn_launches;
        err = cudaGetLastError();
    }
    return err;
}

cudaError_t pn_launch_vmax(const double* v, int64_t ldv, int64_t n, int k, double* vmax, cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(vmax, 0, sizeof(double), stream);
    if (err != cudaSuccess || n <= 0 || k <= 0) return err;
    int64_t blocks = (n * k + 255) / 256;
    blocks = blocks > 1024 ? 1024 : blocks;
    vmax_kernel<<<(unsigned)blocks, 256, 0, stream>>>(v, ldv, n, k, reinterpret_cast<unsigned long long*>(vmax));
    ++g_pn_launches;
    return cudaGetLastError();
}

cudaError_t pn_launch_unbin(const long long* bins, int64_t n0, int kc, int64_t n1, const double* vmax, const double* fac,
                            int fac_ld, int fac_idx, double* y, cudaStream_t stream) {
    if (n0 <= 0) return cudaSuccess;
    unbin_kernel<<<(unsigned)((n0 * kc + 255) / 256), 256, 0, stream>>>(bins, n0, kc, n1, vmax, fac, fac_ld, fac_idx, y);
    ++g_pn_launches;
    return cudaGetLastError();
}
