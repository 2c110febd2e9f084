// Dense Gram matrix of ExpQuad / RatQuad for WIDE inputs (d > 16): squared distances from the norm
// identity |x|^2 + |y|^2 - 2 x.y with the inner products on FP64 tensor cores (mma.sync DMMA, the
// only FP64 tensor path of sm_100a), guarded so that the result keeps the reference's accuracy.
//
// Replaces IsotropicMixin._squared_euclidean_distances + ExpQuad / RatQuad._evaluate
// (/root/reference/src/probnum/randprocs/covfuncs/_covariance_function.py:757-781,
// _exponentiated_quadratic.py:89-100, _rational_quadratic.py:71-84) where the direct differences of
// the reference cost 2 d FP64 instructions per entry (gram_chunk_kernel: 5.25 ms at N = 16384,
// d = 100): the contraction over d is a GEMM, d/4 DMMA.8x8x4 per 64 entries.
//
// Accuracy ("the raw norm identity breaks the 1e-12 contract"):
//   * operands are centred on the mean of both point sets and carry the scale factors
//     (pncov_api.cu: gram_mma_prepare), xt = (x - c) s;  A = xt, B = -2 xt, nrm = |xt|^2;
//   * the accumulator starts at nrm_i + nrm_j and the DMMAs add the d products in index order, so
//     S_ij and S_ji are the same sequence of operations on the same numbers: a symmetric Gram is
//     symmetric bit for bit;
//   * the absolute error of S is at most gamma max(nrm_i, nrm_j), gamma = (4 d + 24) 2^-53.  An
//     entry is kept only if that error moves K by less than 2e-13 |K| (a fifth of the parity
//     contract): ExpQuad  S >= ln(gamma nrm / 2e-13),  RatQuad  1 + c S >= alpha c gamma nrm / 2e-13.
//     The thresholds are per POINT (monotone in nrm: the larger of the two applies to the pair),
//     stored as high words and compared with integer instructions; entries that fail -- near
//     duplicates, the diagonal of a symmetric Gram (always redone: exact 1.0 like the reference),
//     S rounded below zero -- are recomputed by direct differences from the ORIGINAL points with the
//     robust epilogue, exactly like gram_chunk_kernel.  Matern (sqrt(S) near 0) stays on the direct
//     kernels.
//
// Layout: the FP64 GEMM core of csrc/pncov_chol.cu (CTA = 4 warps = 128 x 64 tile, warp 32 x 64,
// 64 accumulators per lane, two-stage cp.async pipeline in K-chunks of 16, row stride 20 doubles).
// The exp2 / log2 tables arrive through the TMA engine into the operand stage that is idle during
// the last K-chunk, and each warp stages its results through the other idle stage (row stride 72
// doubles: conflict-free) so that every output row leaves as 512 contiguous bytes.
#include "pncov_internal.h"

namespace {

constexpr int GM_BM = 128, GM_BN = 64, GM_CK = 16, GM_LD = GM_CK + 4, GM_THREADS = 128;
constexpr int GM_SLD = 72;                                             // staging row stride (doubles)
constexpr int GM_STAGE_BYTES = (GM_BM + GM_BN) * GM_LD * 8;            // one operand stage: 30720
constexpr int GM_TAB_BYTES = 8 * PN_EXP2_N + 16 * PN_LOG2_N;           // 20480
constexpr int GM_SMEM_BYTES = 2 * GM_STAGE_BYTES + 128;                // + barrier
static_assert(GM_TAB_BYTES <= GM_STAGE_BYTES, "the tables fit the idle operand stage");
static_assert(4 * 16 * GM_SLD * 8 <= 2 * GM_STAGE_BYTES - GM_TAB_BYTES, "result staging fits beside the tables");

__device__ __forceinline__ void gm_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ void gm_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pn_smem_u32(smem)), "l"(gmem) : "memory");
}

// one entry by direct differences from the original points (the reference's arithmetic)
static __device__ __noinline__ double gm_redo(const PnGramMmaLaunch& p, int64_t row, int64_t col, const double* tabE,
                                              const double* tabL) {
    const double* a = p.x0_aos + row * p.d;
    const double* b = p.x1_aos + col * p.d;
    double S = 0.0;
    for (int m = 0; m < p.d; ++m) {
        double df = a[m] - b[m];
        if (p.scale) df *= p.scale[m];
        S = pn_fma(df, df, S);
    }
    return pn_epilogue(S, p.epi_direct, tabE, tabL);
}

template <int EPI>
__global__ void __launch_bounds__(GM_THREADS, 3) gram_mma_kernel(const __grid_constant__ PnGramMmaLaunch p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * GM_STAGE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int64_t row_base = (int64_t)blockIdx.y * GM_BM, col_base = (int64_t)blockIdx.x * GM_BN;
    const double* At = p.A + row_base * p.dpad;
    const double* Bt = p.B + col_base * p.dpad;

    if (tid == 0) {
        pn_mbar_init(bar, 1);
        pn_mbar_fence_init();
    }

    auto load_chunk = [&](int kc, int buf) {
        double* as = reinterpret_cast<double*>(smem + buf * GM_STAGE_BYTES);
        double* bs = as + GM_BM * GM_LD;
        constexpr int PIECES = GM_CK / 2;
#pragma unroll
        for (int i = 0; i < GM_BM * PIECES / GM_THREADS; ++i) {
            const int idx = tid + i * GM_THREADS;
            const int row = idx / PIECES, c2 = (idx % PIECES) * 2;
            gm_cp_async16(as + row * GM_LD + c2, At + (int64_t)row * p.dpad + kc * GM_CK + c2);
        }
#pragma unroll
        for (int i = 0; i < GM_BN * PIECES / GM_THREADS; ++i) {
            const int idx = tid + i * GM_THREADS;
            const int row = idx / PIECES, c2 = (idx % PIECES) * 2;
            gm_cp_async16(bs + row * GM_LD + c2, Bt + (int64_t)row * p.dpad + kc * GM_CK + c2);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // accumulators start at |xt_i|^2 + |xt_j|^2
    double acc[4][8][2];
    {
        double nx[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) nx[r] = p.nrmA[row_base + warp * 32 + r * 8 + g];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double2 ny = *reinterpret_cast<const double2*>(p.nrmB + col_base + c * 8 + 2 * q);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[r][c][0] = nx[r] + ny.x;
                acc[r][c][1] = nx[r] + ny.y;
            }
        }
    }

    const int nchunk = p.dpad / GM_CK;
    const int last_buf = (nchunk - 1) & 1;
    // tables go to the stage that is idle during the last chunk, results are staged beside them
    unsigned char* free_stage = smem + (last_buf ^ 1) * GM_STAGE_BYTES;
    double* tabE = reinterpret_cast<double*>(last_buf == 0 ? smem + 2 * GM_STAGE_BYTES - GM_TAB_BYTES : smem);
    double* tabL = tabE + PN_EXP2_N;
    double* stg = reinterpret_cast<double*>(last_buf == 0 ? smem : smem + GM_TAB_BYTES) + warp * (16 * GM_SLD);
    (void)free_stage;

    load_chunk(0, 0);
    __syncthreads();  // barrier initialised
    for (int kc = 0; kc < nchunk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nchunk) {
            load_chunk(kc + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (tid == 0) {  // the other stage is idle from here on: bring the tables in under the last chunk's DMMAs
                const uint32_t bytes = 8u * PN_EXP2_N + (EPI == PN_EPI_RATQUADG ? 16u * PN_LOG2_N : 0u);
                pn_mbar_expect_tx(bar, bytes);
                pn_bulk_g2s(tabE, pn_exp2_table_g, 8u * PN_EXP2_N, bar);
                if (EPI == PN_EPI_RATQUADG) pn_bulk_g2s(tabL, pn_log2_table_g, 16u * PN_LOG2_N, bar);
            }
        }
        __syncthreads();
        const double* as = reinterpret_cast<const double*>(smem + buf * GM_STAGE_BYTES) + (warp * 32 + g) * GM_LD + q;
        const double* bs = reinterpret_cast<const double*>(smem + buf * GM_STAGE_BYTES) + GM_BM * GM_LD + g * GM_LD + q;
#pragma unroll
        for (int k4 = 0; k4 < GM_CK / 4; ++k4) {
            double a[4], b[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = as[r * 8 * GM_LD + k4 * 4];
#pragma unroll
            for (int c = 0; c < 8; ++c) b[c] = bs[c * 8 * GM_LD + k4 * 4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) gm_dmma(acc[r][c][0], acc[r][c][1], a[r], b[c]);
        }
        __syncthreads();
    }
    pn_mbar_wait(bar, 0);

    // ---- epilogue: guard, kernel function, staged 512-byte row stores ---------------------------------
    const PnEpiParams& P = p.epi;
    const bool diag_tile = p.same_points && (p.row0 + row_base < col_base + GM_BN) && (col_base < p.row0 + row_base + GM_BM);
    const unsigned v_hi = P.s_hi_word;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * h + rr;
            const int64_t lrow = row_base + warp * 32 + r * 8 + g;  // row within this launch
            const int thr_r = p.thrA[lrow];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int64_t col = col_base + c * 8 + 2 * q;
                const int2 thr_c = *reinterpret_cast<const int2*>(p.thrB + col);
                const double s0 = acc[r][c][0], s1 = acc[r][c][1];
                double v0, v1, k0, k1;
                if constexpr (EPI == PN_EPI_EXPQUAD) {
                    v0 = s0;
                    v1 = s1;
                    k0 = pn_exp2_prod_core(s0, P.c_u, P, tabE);
                    k1 = pn_exp2_prod_core(s1, P.c_u, P, tabE);
                } else {
                    v0 = pn_fma(s0, P.c_s, 1.0);
                    v1 = pn_fma(s1, P.c_s, 1.0);
                    if constexpr (EPI == PN_EPI_RATQUAD1) {
                        k0 = pn_rcp_core(v0);
                        k1 = pn_rcp_core(v1);
                    } else if constexpr (EPI == PN_EPI_RATQUAD2) {
                        k0 = pn_rcp_core(v0);
                        k1 = pn_rcp_core(v1);
                        k0 *= k0;
                        k1 *= k1;
                    } else {
                        k0 = pn_exp2_prod_core(pn_log2_core(v0, P, tabL), -P.alpha, P, tabE);
                        k1 = pn_exp2_prod_core(pn_log2_core(v1, P, tabL), -P.alpha, P, tabE);
                    }
                }
                // keep the entry iff thr <= hi(v) < v_hi (signed hi words as unsigned differences: a negative
                // or NaN value fails as well)
                const int t0 = max(thr_r, thr_c.x), t1 = max(thr_r, thr_c.y);
                bool bad0 = (unsigned)(pn_hi(v0) - t0) >= v_hi - (unsigned)t0;
                bool bad1 = (unsigned)(pn_hi(v1) - t1) >= v_hi - (unsigned)t1;
                if (diag_tile) {
                    bad0 |= p.row0 + lrow == col;
                    bad1 |= p.row0 + lrow == col + 1;
                }
                if (bad0 | bad1) {  // rare: near duplicates, the diagonal, underflowing results
                    if (bad0 && lrow < p.n0 && col < p.n1) k0 = gm_redo(p, lrow, col, tabE, tabL);
                    if (bad1 && lrow < p.n0 && col + 1 < p.n1) k1 = gm_redo(p, lrow, col + 1, tabE, tabL);
                }
                *reinterpret_cast<double2*>(stg + (rr * 8 + g) * GM_SLD + c * 8 + 2 * q) = make_double2(k0, k1);
            }
        }
        __syncwarp();
        const int64_t lrow0 = row_base + warp * 32 + h * 16;
        const int64_t col = col_base + 2 * lane;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int64_t lrow = lrow0 + i;
            if (lrow >= p.n0) break;
            const double2 v = *reinterpret_cast<const double2*>(stg + i * GM_SLD + 2 * lane);
            double* dst = p.out + lrow * p.ldo + col;
            if (col + 1 < p.n1) {
                if (p.vec_ok) pn_store_cs2(dst, v.x, v.y);
                else {
                    pn_store_cs1(dst, v.x);
                    pn_store_cs1(dst + 1, v.y);
                }
            } else if (col < p.n1) {
                pn_store_cs1(dst, v.x);
            }
        }
        __syncwarp();
    }
}

template <int EPI>
cudaError_t launch_gm(const PnGramMmaLaunch& p, cudaStream_t stream) {
    auto kern = gram_mma_kernel<EPI>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES);
    if (err != cudaSuccess) return err;
    const int64_t gx = (p.n1 + GM_BN - 1) / GM_BN, gy = (p.n0 + GM_BM - 1) / GM_BM;
    if (gy > 65535) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)gx, (unsigned)gy);
    kern<<<grid, GM_THREADS, GM_SMEM_BYTES, stream>>>(p);
    ++g_pn_launches;
    return cudaGetLastError();
}

// centred + scaled GEMM operands of one point set: A = xt, B = -2 xt (npad x dpad, zero padded), nrm = |xt|^2 and the
// guard threshold of every point as a high word (see the header comment; `gam_over_tol` = gamma / 2e-13 for ExpQuad,
// alpha c gamma / 2e-13 for RatQuad)
__global__ void gm_build_operands_kernel(const double* __restrict__ aos, int64_t n, int64_t npad, int d, int dpad,
                                         const double* __restrict__ centre, const double* __restrict__ scale, double iso_scale,
                                         double gam_over_tol, int expquad, double* __restrict__ A, double* __restrict__ B,
                                         double* __restrict__ nrm, int* __restrict__ thr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    double s = 0.0;
    for (int m = 0; m < dpad; ++m) {
        double v = 0.0;
        if (i < n && m < d) v = (aos[i * d + m] - centre[m]) * (scale ? scale[m] : iso_scale);
        A[i * dpad + m] = v;
        B[i * dpad + m] = -2.0 * v;
        s = fma(v, v, s);
    }
    nrm[i] = s;
    // ExpQuad: S >= ln(w);  RatQuad: 1 + c S >= w;  w = gam_over_tol * nrm.  Rounded UP to the next high word.
    const double w = gam_over_tol * s;
    double t = expquad ? (w > 1.0 ? log(w) : 0.0) : w;
    int word = 0;
    if (t > 0.0) word = __double2hiint(t) + 1;
    thr[i] = (i < n) ? word : 0;
}

}  // namespace

cudaError_t pn_launch_gram_mma_operands(const double* aos, int64_t n, int64_t npad, int d, int dpad, const double* centre,
                                        const double* scale, double iso_scale, double gam_over_tol, int expquad, double* A,
                                        double* B, double* nrm, int* thr, cudaStream_t stream) {
    gm_build_operands_kernel<<<(unsigned)((npad + 127) / 128), 128, 0, stream>>>(aos, n, npad, d, dpad, centre, scale, iso_scale,
                                                                                  gam_over_tol, expquad, A, B, nrm, thr);
    ++g_pn_launches;
    return cudaGetLastError();
}

cudaError_t pn_launch_gram_mma(const PnGramMmaLaunch& p0, cudaStream_t stream) {
    if (p0.n0 <= 0 || p0.n1 <= 0) return cudaSuccess;
    PnGramMmaLaunch p = p0;
    p.vec_ok = ((p.ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(p.out) % 16) == 0) ? 1 : 0;
    switch (p.epi.epi) {
        case PN_EPI_EXPQUAD: return launch_gm<PN_EPI_EXPQUAD>(p, stream);
        case PN_EPI_RATQUAD1: return launch_gm<PN_EPI_RATQUAD1>(p, stream);
        case PN_EPI_RATQUAD2: return launch_gm<PN_EPI_RATQUAD2>(p, stream);
        case PN_EPI_RATQUADG: return launch_gm<PN_EPI_RATQUADG>(p, stream);
        default: return cudaErrorInvalidValue;
    }
}
