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
//     S rounded below zero or within 64 error bounds of zero (so that coincident points give exactly
//     1.0) -- are recomputed by direct differences from the ORIGINAL points with the robust epilogue,
//     exactly like gram_chunk_kernel.  Matern (sqrt(S) near 0) stays on the direct kernels.
//
// Layout: the FP64 GEMM core of csrc/pncov_chol.cu (CTA = 4 warps = 128 x 64 tile, warp 32 x 64,
// 64 accumulators per lane, two-stage cp.async pipeline in K-chunks of 16, row stride 20 doubles).
// The exp2 / log2 tables arrive through the TMA engine into the operand stage that is idle during
// the last K-chunk, and each warp stages its results through the other idle stage (row stride 72
// doubles: conflict-free) so that every output row leaves as 512 contiguous bytes.
#include "pncov_internal.h"

namespace {

constexpr int GM_BM = 128, GM_BN = 64, GM_CK = 16, GM_LD = GM_CK + 4, GM_THREADS = 128;
constexpr int GM_STAGE_BYTES = (GM_BM + GM_BN) * GM_LD * 8;            // one operand stage: 30720
constexpr int GM_TAB_BYTES = 8 * PN_EXP2_N + 16 * PN_LOG2_N;           // 20480
constexpr int GM_SMEM_BYTES = 2 * GM_STAGE_BYTES + 128;                // + barrier
static_assert(GM_TAB_BYTES <= GM_STAGE_BYTES, "the tables fit the idle operand stage");
static_assert(4 * 16 * 64 * 8 <= 2 * GM_STAGE_BYTES - GM_TAB_BYTES, "result staging fits beside the tables");

__device__ __forceinline__ void gm_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// one entry by direct differences from the original points (the reference's arithmetic)
static __device__ __noinline__ double gm_redo(const PnGramMmaLaunch& p, int64_t row, int64_t col, const double* tabE,
                                              const double* tabL) {
    const double* a = p.x0_aos + row * p.d;
    const double* b = p.x1_aos + col * p.d;
    double S = 0.0;
    int m = 0;
    for (; m + 4 <= p.d; m += 4) {  // four independent loads per operand in flight
        double df[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) df[u] = __ldg(a + m + u) - __ldg(b + m + u);
        if (p.scale) {
#pragma unroll
            for (int u = 0; u < 4; ++u) df[u] *= __ldg(p.scale + m + u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) S = pn_fma(df[u], df[u], S);
    }
    for (; m < p.d; ++m) {
        double df = __ldg(a + m) - __ldg(b + m);
        if (p.scale) df *= __ldg(p.scale + m);
        S = pn_fma(df, df, S);
    }
    return pn_epilogue(S, p.epi_direct, tabE, tabL);
}

// An entry the fast path did not accept.  Accuracy failures (value below the guard threshold of its two points,
// negative, the diagonal of a symmetric Gram) are recomputed from the original points; values that only left the
// window of the FAST epilogue (results at or below the underflow threshold) keep their S -- its relative error is
// as small as anywhere else -- and go through the robust epilogue, ExpQuad results far below the smallest
// subnormal (S > 750 in scaled units, the reference's exp gives 0.0 from 745.14 on) are 0 directly.
static __device__ __noinline__ double gm_fix(const PnGramMmaLaunch& p, int64_t row, int64_t col, double s, int hi, int thr,
                                             bool on_diag, const double* tabE, const double* tabL) {
    if (row >= p.n0 || col >= p.n1) return 0.0;  // padding
    // k(x, x): the reference's differences are exactly 0 and ExpQuad / RatQuad give exactly 1 -- unless the point has
    // non-finite coordinates (s is NaN then), which the direct form below turns into the reference's NaN
    if (on_diag && s == s) return 1.0;
    if (on_diag || hi < thr) return gm_redo(p, row, col, tabE, tabL);
    if (s != s) return gm_redo(p, row, col, tabE, tabL);  // NaN: let the direct form decide
    if (p.epi.epi == PN_EPI_EXPQUAD && s > 750.0) return 0.0;
    return pn_epilogue(s, p.epi, tabE, tabL);
}

template <int EPI>
__global__ void __launch_bounds__(GM_THREADS, 3) gram_mma_kernel(const __grid_constant__ PnGramMmaLaunch p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * GM_STAGE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    int64_t row_base, col_base;
    if (p.mirror) {
        // 1-D grid over the tiles on and above the diagonal blocks only: row block I holds the column tiles 2 I .. gx - 1,
        // c(I) = I (gx + 1) - I^2 tiles precede it
        const int64_t t = blockIdx.x, gx = p.mirror;
        int64_t I = (int64_t)(0.5 * ((double)(gx + 1) - sqrt((double)((gx + 1) * (gx + 1) - 4 * t))));
        I = I < 0 ? 0 : I;
        while (I * (gx + 1) - I * I > t) --I;
        while ((I + 1) * (gx + 1) - (I + 1) * (I + 1) <= t) ++I;
        row_base = I * GM_BM;
        col_base = (2 * I + (t - (I * (gx + 1) - I * I))) * GM_BN;
    } else {
        row_base = (int64_t)blockIdx.y * GM_BM;
        col_base = (int64_t)blockIdx.x * GM_BN;
    }
    const double* At = p.A + row_base * p.dpad;
    const double* Bt = p.B + col_base * p.dpad;

    if (tid == 0) {
        pn_mbar_init(bar, 1);
        pn_mbar_fence_init();
    }

    // every thread copies the same 16-byte pieces of each chunk: rows tid / 8 + 16 i, doubles 2 (tid % 8) .. + 1
    constexpr int PIECES = GM_CK / 2;
    const double* srcA = At + (int64_t)(tid / PIECES) * p.dpad + (tid % PIECES) * 2;
    const double* srcB = Bt + (int64_t)(tid / PIECES) * p.dpad + (tid % PIECES) * 2;
    const int64_t src_step = (int64_t)(GM_THREADS / PIECES) * p.dpad;
    const uint32_t dst0 = pn_smem_u32(smem) + (uint32_t)(((tid / PIECES) * GM_LD + (tid % PIECES) * 2) * 8);
    auto load_chunk = [&](int kc, int buf) {
        const uint32_t da = dst0 + (uint32_t)(buf * GM_STAGE_BYTES);
        const double* sa = srcA + kc * GM_CK;
#pragma unroll
        for (int i = 0; i < GM_BM * PIECES / GM_THREADS; ++i)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da + (uint32_t)(i * (GM_THREADS / PIECES) * GM_LD * 8)),
                         "l"(sa + i * src_step)
                         : "memory");
        const uint32_t db = da + (uint32_t)(GM_BM * GM_LD * 8);
        const double* sb = srcB + kc * GM_CK;
#pragma unroll
        for (int i = 0; i < GM_BN * PIECES / GM_THREADS; ++i)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(db + (uint32_t)(i * (GM_THREADS / PIECES) * GM_LD * 8)),
                         "l"(sb + i * src_step)
                         : "memory");
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
    double* stg = reinterpret_cast<double*>(last_buf == 0 ? smem : smem + GM_TAB_BYTES) + warp * (16 * 64);
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
        const int k4n = kc + 1 < nchunk ? GM_CK / 4 : p.k4_last;  // the last chunk may hold fewer than 16 coordinates
#pragma unroll
        for (int k4 = 0; k4 < GM_CK / 4; ++k4) {
            if (k4 >= k4n) break;
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

    // ---- epilogue ---------------------------------------------------------------------------------------------
    // The accumulator fragments (S) of 16 rows x 64 columns per warp go to shared memory, are read back one ROW per
    // step (two adjacent columns per lane: the guard thresholds of the columns stay in registers, the row's is warp
    // uniform, and no accumulator is live any more, so several rows are in flight per lane), pass the guard and the
    // kernel function and leave as 512 contiguous bytes per row.  Rows are 32 chunks of 16 bytes, chunk p of row i at
    // position p ^ sw(i), sw(i) = 4 (i & 1) | ((i >> 1) & 3): fragment stores and row reads are conflict free, the
    // column reads of the mirrored half are two-way.
    const PnEpiParams& P = p.epi;
    const bool diag_tile = p.same_points && (p.row0 + row_base < col_base + GM_BN) && (col_base < p.row0 + row_base + GM_BM);
    const bool mirror = p.mirror && col_base >= row_base + GM_BM;  // strictly above the diagonal block: K[j][i] = K[i][j]
    const unsigned v_hi = p.v_hi;
    const bool interior = p.vec_ok && row_base + GM_BM <= p.n0 && col_base + GM_BN <= p.n1;
    const int64_t col = col_base + 2 * lane;
    const int2 thr_c = *reinterpret_cast<const int2*>(p.thrB + col);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int i = rr * 8 + g;
            const int sw = ((i & 1) << 2) | ((i >> 1) & 3);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<double2*>(stg + i * 64 + (((4 * c + q) ^ sw) << 1)) =
                    make_double2(acc[2 * h + rr][c][0], acc[2 * h + rr][c][1]);
        }
        __syncwarp();
        const int64_t lrow0 = row_base + warp * 32 + h * 16;
        double* dst = p.out + lrow0 * p.ldo + col;
        const int* thr_rows = p.thrA + lrow0;
#pragma unroll 4
        for (int i = 0; i < 16; ++i, dst += p.ldo) {
            const int sw = ((i & 1) << 2) | ((i >> 1) & 3);
            double2* slot = reinterpret_cast<double2*>(stg + i * 64 + ((lane ^ sw) << 1));
            const double2 s = *slot;
            const int thr_r = thr_rows[i];
            double v0, v1, k0, k1;
            if constexpr (EPI == PN_EPI_EXPQUAD) {
                v0 = s.x;
                v1 = s.y;
                k0 = pn_exp2_prod_core(s.x, P.c_u, P, tabE);
                k1 = pn_exp2_prod_core(s.y, P.c_u, P, tabE);
            } else {
                v0 = pn_fma(s.x, P.c_s, 1.0);
                v1 = pn_fma(s.y, P.c_s, 1.0);
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
            // keep the entry iff thr <= hi(v) < v_hi: one unsigned comparison (a negative or NaN value fails too)
            const int t0 = max(thr_r, thr_c.x), t1 = max(thr_r, thr_c.y);
            const int h0 = pn_hi(v0), h1 = pn_hi(v1);
            bool bad0 = (unsigned)(h0 - t0) >= v_hi - (unsigned)t0;
            bool bad1 = (unsigned)(h1 - t1) >= v_hi - (unsigned)t1;
            if (diag_tile) {
                bad0 |= p.row0 + lrow0 + i == col;
                bad1 |= p.row0 + lrow0 + i == col + 1;
            }
            if (bad0 | bad1) {  // rare: near duplicates, the diagonal, results at the underflow threshold
                const int64_t lrow = lrow0 + i;  // row within this launch
                if (bad0) k0 = gm_fix(p, lrow, col, s.x, h0, t0, diag_tile && p.row0 + lrow == col, tabE, tabL);
                if (bad1) k1 = gm_fix(p, lrow, col + 1, s.y, h1, t1, diag_tile && p.row0 + lrow == col + 1, tabE, tabL);
            }
            if (mirror) *slot = make_double2(k0, k1);
            if (interior) {  // uniform per CTA: no bounds to check, 16-byte aligned rows
                pn_store_cs2(dst, k0, k1);
            } else if (lrow0 + i < p.n0) {
                if (col + 1 < p.n1 && p.vec_ok) pn_store_cs2(dst, k0, k1);
                else {
                    if (col < p.n1) pn_store_cs1(dst, k0);
                    if (col + 1 < p.n1) pn_store_cs1(dst + 1, k1);
                }
            }
        }
        __syncwarp();
        if (mirror) {  // uniform per CTA: the transposed 64 x 16 block, 128 contiguous bytes per mirrored row
            const int t = lane & 15;
            const int sw = ((t & 1) << 2) | ((t >> 1) & 3);
            const int64_t mcol = lrow0 + t;  // column of the mirrored entry
            const double* srow = stg + t * 64 + (lane >> 4);
            double* mdst = p.out + (col_base + (lane >> 4)) * p.ldo + mcol;
            const int64_t mstep = 2 * p.ldo;
#pragma unroll 8
            for (int jj = 0; jj < 32; ++jj, mdst += mstep) {
                const double v = srow[(jj ^ sw) << 1];
                if (interior || (col_base + 2 * jj + (lane >> 4) < p.n1 && mcol < p.n0)) pn_store_cs1(mdst, v);
            }
            __syncwarp();
        }
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
    if (p.mirror) {
        const int64_t tiles = gy * gx - gy * (gy - 1);
        if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        grid = dim3((unsigned)tiles, 1);
    }
    kern<<<grid, GM_THREADS, GM_SMEM_BYTES, stream>>>(p);
    ++g_pn_launches;
    return cudaGetLastError();
}

// centre of both point sets: one CTA per coordinate sums its row of the coordinate-major copies (zero padded, so
// the padding does not count) in a fixed order
__global__ void __launch_bounds__(256) gm_centre_kernel(const double* __restrict__ soa0, int64_t ld0,
                                                        const double* __restrict__ soa1, int64_t ld1, double ntot,
                                                        double* __restrict__ centre) {
    __shared__ double red[256];
    const int m = blockIdx.x;
    double s = 0.0;
    // non-finite coordinates do not move the centre: they poison only their own rows and columns, like the reference
    for (int64_t i = threadIdx.x; i < ld0; i += 256) {
        const double v = soa0[(size_t)m * ld0 + i];
        if (isfinite(v)) s += v;
    }
    if (soa1)
        for (int64_t i = threadIdx.x; i < ld1; i += 256) {
            const double v = soa1[(size_t)m * ld1 + i];
            if (isfinite(v)) s += v;
        }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) centre[m] = isfinite(red[0]) ? red[0] / ntot : 0.0;
}

// centred + scaled GEMM operands of one point set: A = xt, B = -2 xt (rows_alloc x dpad, zero padded), nrm = |xt|^2
// and the guard threshold of every point as a high word (see the header comment; `gam_over_tol` = gamma / 2e-13 for
// ExpQuad, alpha c gamma / 2e-13 for RatQuad).  One warp per point: coalesced along the coordinates.
__global__ void __launch_bounds__(256) gm_build_operands_kernel(const double* __restrict__ aos, int64_t n, int64_t rows_alloc,
                                                                int d, int dpad, const double* __restrict__ centre,
                                                                const double* __restrict__ scale, double iso_scale,
                                                                double gam_over_tol, double dup_factor, int expquad,
                                                                double* __restrict__ A,
                                                                double* __restrict__ B, double* __restrict__ nrm,
                                                                int* __restrict__ thr) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= rows_alloc) return;
    // |xt|^2 in index order (one lane would do; all lanes compute the same chain so that no shuffle tree
    // changes the order): the value the DMMA chain cancels against
    double s = 0.0;
    for (int m0 = 0; m0 < dpad; m0 += 32) {
        const int m = m0 + lane;
        double v = 0.0;
        if (i < n && m < d) v = (aos[i * d + m] - centre[m]) * (scale ? scale[m] : iso_scale);
        if (m < dpad) {
            A[i * dpad + m] = v;
            B[i * dpad + m] = -2.0 * v;
        }
        for (int l = 0; l < 32; ++l) {
            const double vl = __shfl_sync(0xffffffffu, v, l);
            s = fma(vl, vl, s);
        }
    }
    if (lane != 0) return;
    nrm[i] = s;
    // ExpQuad: S >= ln(w);  RatQuad: 1 + c S >= w;  w = gam_over_tol * nrm.  Rounded UP to the next high word.
    // Second guard: v below 64 error bounds of S (near-coincident points) is redone as well, so that coincident
    // points give exactly 1.0 like the reference's direct differences.
    const double w = gam_over_tol * s;
    double t = expquad ? fmax(w > 1.0 ? log(w) : 0.0, dup_factor * w) : fmax(w, 1.0 + dup_factor * w);
    int word = 0;
    if (t > 0.0) word = __double2hiint(t) + 1;
    if (!(w < 1e12)) word = 0x7ff00000;  // absurd or non-finite norms: every entry of this point is redone
    thr[i] = (i < n) ? word : 0;
}

}  // namespace

cudaError_t pn_launch_gram_mma_centre(const double* soa0, int64_t ld0, const double* soa1, int64_t ld1, double ntot, int d,
                                      double* centre, cudaStream_t stream) {
    gm_centre_kernel<<<(unsigned)d, 256, 0, stream>>>(soa0, ld0, soa1, ld1, ntot, centre);
    ++g_pn_launches;
    return cudaGetLastError();
}

cudaError_t pn_launch_gram_mma_operands(const double* aos, int64_t n, int64_t rows_alloc, int d, int dpad, const double* centre,
                                        const double* scale, double iso_scale, double gam_over_tol, double dup_factor,
                                        int expquad, double* A, double* B, double* nrm, int* thr, cudaStream_t stream) {
    gm_build_operands_kernel<<<(unsigned)((rows_alloc + 7) / 8), 256, 0, stream>>>(aos, n, rows_alloc, d, dpad, centre, scale,
                                                                                    iso_scale, gam_over_tol, dup_factor, expquad,
                                                                                    A, B, nrm, thr);
    ++g_pn_launches;
    return cudaGetLastError();
}

cudaError_t pn_launch_gram_mma(const PnGramMmaLaunch& p0, cudaStream_t stream) {
    if (p0.n0 <= 0 || p0.n1 <= 0) return cudaSuccess;
    PnGramMmaLaunch p = p0;
    p.vec_ok = ((p.ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(p.out) % 16) == 0) ? 1 : 0;
    // whole symmetric matrix: `mirror` carries the number of column tiles (> 0)
    p.mirror = (p.same_points && p.row0 == 0 && p.n0 == p.n1) ? (int)((p.n1 + GM_BN - 1) / GM_BN) : 0;
    p.k4_last = (p.d + 3) / 4 - (GM_CK / 4) * (p.dpad / GM_CK - 1);  // DMMA k-steps (of 4 coordinates) in the last chunk
    switch (p.epi.epi) {
        case PN_EPI_EXPQUAD: return launch_gm<PN_EPI_EXPQUAD>(p, stream);
        case PN_EPI_RATQUAD1: return launch_gm<PN_EPI_RATQUAD1>(p, stream);
        case PN_EPI_RATQUAD2: return launch_gm<PN_EPI_RATQUAD2>(p, stream);
        case PN_EPI_RATQUADG: return launch_gm<PN_EPI_RATQUADG>(p, stream);
        default: return cudaErrorInvalidValue;
    }
}
