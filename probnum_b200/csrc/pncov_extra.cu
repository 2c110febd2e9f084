// Gram matrices of the non-isotropic covariance functions and device-side covariance arithmetic
// (SURVEY.md 8f, row N3).  Paths relative to /root/reference/src/probnum/randprocs/covfuncs:
//
//   Linear        k = <x0, x1> + c                          _linear.py:44-50
//   Polynomial    k = (<x0, x1> + c)^q, q = 0, 1, 2, ...      _polynomial.py:50-56
//   WhiteNoise    k = sigma^2 [x0 == x1]                    _white_noise.py:38-55
//   ProductMatern k = prod_m Matern_{nu_m}(|x0_m - x1_m| / l_m)   _product_matern.py:103-122
//   Scaled / Sum / Product covariance functions             _arithmetic_fallbacks.py:21-180
//
// All of them are HBM-write-bound for small d, like the isotropic Gram kernel, so the layout is
// the same: a CTA of 256 threads owns a strip of 512 columns, a thread owns two adjacent columns
// and emits one 16-byte streaming store per row; the strip's column points are staged once in
// shared memory (coordinate-major, up to X_DC coordinates at a time), row points are read as
// warp broadcasts.  float64 only (the host side promotes float32 inputs, which is also what the
// reference's NumPy-2 arithmetic produces for these kernels).
#include "pncov_internal.h"

namespace {

constexpr int X_THREADS = 256;
constexpr int X_TN = 2 * X_THREADS;  // columns per strip
constexpr int X_R = 8;               // rows per register tile
constexpr int X_DC = 16;             // coordinates staged per chunk
constexpr int X_RS = 64;             // rows staged at a time when d <= X_DC

enum { X_DOT = 0, X_EQUAL = 1, X_PRODMATERN = 2 };

// (x + c)^q: q - 1 roundings for q <= 3, binary powering (O(log q) multiplications) above
__device__ __forceinline__ double x_powi(double b, int q) {
    if (q == 2) return b * b;
    if (q == 3) return b * b * b;
    double r = 1.0;
    while (q > 0) {
        if (q & 1) r *= b;
        b *= b;
        q >>= 1;
    }
    return r;
}

// univariate half-integer Matern polynomial in r = |x0_m - x1_m| * s_m with exp(-r) deferred;
// pm = {s_m, p_m, c_0 .. c_8}
__device__ __forceinline__ double x_matern_poly(double r, const double* pm) {
    const int p = (int)pm[1];
    double poly = pm[2 + p];
    for (int i = p - 1; i >= 0; --i) poly = pn_fma(poly, r, pm[2 + i]);
    return poly;
}

// one coordinate of the ProductMatern accumulation for a register tile, degree known at compile
// time: A += r, B *= poly_P(r); 4 + P FP64 instructions per entry
template <int P>
__device__ __forceinline__ void x_pm_dim(double (&A)[2 * X_R], double (&B)[2 * X_R], const double* __restrict__ xa,
                                         double2 xb, const double* __restrict__ pm) {
    const double s = pm[0];
    double c[P + 1];
#pragma unroll
    for (int i = 0; i <= P; ++i) c[i] = pm[2 + i];
#pragma unroll
    for (int r = 0; r < X_R; ++r) {
        // rounded product, then a plain addition (no FMA contraction): the reference's r_m and their sum
        const double ra = __dmul_rn(fabs(xa[r] - xb.x), s);
        const double rb = __dmul_rn(fabs(xa[r] - xb.y), s);
        A[2 * r] = __dadd_rn(A[2 * r], ra);
        A[2 * r + 1] = __dadd_rn(A[2 * r + 1], rb);
        if constexpr (P > 0) {
            double pa = c[P], pb = c[P];
#pragma unroll
            for (int i = P - 1; i >= 0; --i) {
                pa = pn_fma(pa, ra, c[i]);
                pb = pn_fma(pb, rb, c[i]);
            }
            B[2 * r] *= pa;
            B[2 * r + 1] *= pb;
        }
    }
}

// accumulate `dc` coordinates into the register tile; xa0 points at this tile's first row of the
// staged row coordinates (coordinate stride lda), x1c is the staged column chunk
template <int MODE>
__device__ __forceinline__ void x_accumulate(double (&A)[2 * X_R], double (&B)[2 * X_R], const double* __restrict__ xa0,
                                             int lda, const double* __restrict__ x1c, const double* __restrict__ pmc,
                                             int dc, int tid) {
    for (int m = 0; m < dc; ++m) {
        const double2 xb = *reinterpret_cast<const double2*>(x1c + (size_t)m * X_TN + 2 * tid);
        const double* xa = xa0 + m * lda;
        if constexpr (MODE == X_DOT) {
#pragma unroll
            for (int r = 0; r < X_R; ++r) {
                A[2 * r] = pn_fma(xa[r], xb.x, A[2 * r]);
                A[2 * r + 1] = pn_fma(xa[r], xb.y, A[2 * r + 1]);
            }
        } else if constexpr (MODE == X_EQUAL) {
#pragma unroll
            for (int r = 0; r < X_R; ++r) {
                A[2 * r] += (xa[r] == xb.x) ? 0.0 : 1.0;  // NaN != NaN, as in NumPy
                A[2 * r + 1] += (xa[r] == xb.y) ? 0.0 : 1.0;
            }
        } else {
            const double* pm = pmc + m * PN_PM_LD;
            switch ((int)pm[1]) {  // uniform across the CTA
                case 0: x_pm_dim<0>(A, B, xa, xb, pm); break;
                case 1: x_pm_dim<1>(A, B, xa, xb, pm); break;
                case 2: x_pm_dim<2>(A, B, xa, xb, pm); break;
                case 3: x_pm_dim<3>(A, B, xa, xb, pm); break;
                default:
#pragma unroll
                    for (int r = 0; r < X_R; ++r) {
                        const double ra = __dmul_rn(fabs(xa[r] - xb.x), pm[0]);
                        const double rb = __dmul_rn(fabs(xa[r] - xb.y), pm[0]);
                        A[2 * r] = __dadd_rn(A[2 * r], ra);
                        A[2 * r + 1] = __dadd_rn(A[2 * r + 1], rb);
                        B[2 * r] *= x_matern_poly(ra, pm);
                        B[2 * r + 1] *= x_matern_poly(rb, pm);
                    }
            }
        }
    }
}

template <int MODE>
__device__ __forceinline__ void x_finish_store(const double (&A)[2 * X_R], const double (&B)[2 * X_R],
                                               const PnExtraLaunch& g, const double* tabE, int64_t row0, int64_t j0) {
    double* out = static_cast<double*>(g.out);
#pragma unroll
    for (int r = 0; r < X_R; ++r) {
        double k0, k1;
        if constexpr (MODE == X_DOT) {
            k0 = A[2 * r] + g.constant;
            k1 = A[2 * r + 1] + g.constant;
            if (g.exponent != 1) {
                k0 = x_powi(k0, g.exponent);
                k1 = x_powi(k1, g.exponent);
            }
        } else if constexpr (MODE == X_EQUAL) {
            k0 = A[2 * r] == 0.0 ? g.constant : 0.0;
            k1 = A[2 * r + 1] == 0.0 ? g.constant : 0.0;
        } else {
            // exp(-sum_m r_m) once for the whole product; pn_exp2 takes its robust path for huge
            // or non-finite arguments
            k0 = B[2 * r] * pn_exp2(A[2 * r] * -1.4426950408889634074, tabE);
            k1 = B[2 * r + 1] * pn_exp2(A[2 * r + 1] * -1.4426950408889634074, tabE);
        }
        const int64_t row = row0 + r;
        if (row < g.n0) {
            double* p = out + row * g.ldo + j0;
            if (j0 + 1 < g.n1) {
                if (g.vec_ok) pn_store_cs2(p, k0, k1);
                else {
                    pn_store_cs1(p, k0);
                    pn_store_cs1(p + 1, k1);
                }
            } else if (j0 < g.n1) {
                pn_store_cs1(p, k0);
            }
        }
    }
}

// d <= X_DC: the strip's column coordinates are staged once per CTA, row coordinates X_RS rows at
// a time (one barrier pair per 64 rows); d > X_DC: coordinates stream through shared memory in
// chunks of X_DC and a CTA iteration covers one 8-row register tile.
template <int MODE>
__global__ void __launch_bounds__(X_THREADS)
    gram_extra_kernel(const __grid_constant__ PnExtraLaunch g, int rows_per_cta) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* tabE = reinterpret_cast<double*>(smem);                   // PRODMATERN only
    double* x1c = reinterpret_cast<double*>(smem + 8 * PN_EXP2_N);   // [X_DC][X_TN]
    double* x0c = x1c + X_DC * X_TN;                                  // [X_DC][X_RS]
    double* pmc = x0c + X_DC * X_RS;                                  // [X_DC][PN_PM_LD]
    const int tid = threadIdx.x;
    const int64_t col0 = (int64_t)blockIdx.x * X_TN;
    const int64_t j0 = col0 + 2 * tid;
    const double* x0g = static_cast<const double*>(g.x0_soa);
    const double* x1g = static_cast<const double*>(g.x1_soa);
    if constexpr (MODE == X_PRODMATERN) {
        for (int i = tid; i < PN_EXP2_N; i += X_THREADS) tabE[i] = pn_exp2_table_g[i];
    }
    const int64_t rbeg = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t rend = (rbeg + rows_per_cta < g.n0) ? rbeg + rows_per_cta : g.n0;

    if (g.d <= X_DC) {
        const int dc = g.d;
        for (int idx = tid; idx < dc * X_TN; idx += X_THREADS) {
            const int m = idx / X_TN, j = idx % X_TN;
            x1c[idx] = x1g[(size_t)m * g.ld1 + col0 + j];  // padded to PN_PAD columns
        }
        if constexpr (MODE == X_PRODMATERN) {
            for (int idx = tid; idx < dc * PN_PM_LD; idx += X_THREADS) pmc[idx] = static_cast<const double*>(g.pm)[idx];
        }
        for (int64_t s0 = rbeg; s0 < rend; s0 += X_RS) {
            __syncthreads();
            for (int idx = tid; idx < dc * X_RS; idx += X_THREADS) {
                const int m = idx / X_RS, r = idx % X_RS;
                x0c[idx] = (s0 + r < g.n0) ? x0g[(size_t)m * g.ld0 + s0 + r] : 0.0;
            }
            __syncthreads();
            const int nrows = (rend - s0 < X_RS) ? (int)(rend - s0) : X_RS;
            for (int rt = 0; rt < nrows; rt += X_R) {
                double A[2 * X_R], B[2 * X_R];
#pragma unroll
                for (int i = 0; i < 2 * X_R; ++i) {
                    A[i] = 0.0;
                    B[i] = 1.0;
                }
                x_accumulate<MODE>(A, B, x0c + rt, X_RS, x1c, pmc, dc, tid);
                x_finish_store<MODE>(A, B, g, tabE, s0 + rt, j0);
            }
        }
        return;
    }

    for (int64_t row0 = rbeg; row0 < rend; row0 += X_R) {
        double A[2 * X_R], B[2 * X_R];  // DOT: A = sum; EQUAL: A = #differences; PRODMATERN: A = sum r, B = prod poly
#pragma unroll
        for (int i = 0; i < 2 * X_R; ++i) {
            A[i] = 0.0;
            B[i] = 1.0;
        }
        for (int c0 = 0; c0 < g.d; c0 += X_DC) {
            const int dc = (g.d - c0 < X_DC) ? g.d - c0 : X_DC;
            __syncthreads();
            for (int idx = tid; idx < dc * X_TN; idx += X_THREADS) {
                const int m = idx / X_TN, j = idx % X_TN;
                x1c[idx] = x1g[(size_t)(c0 + m) * g.ld1 + col0 + j];
            }
            if constexpr (MODE == X_PRODMATERN) {
                for (int idx = tid; idx < dc * PN_PM_LD; idx += X_THREADS)
                    pmc[idx] = static_cast<const double*>(g.pm)[(size_t)c0 * PN_PM_LD + idx];
            }
            if (tid < dc * X_R) {
                const int m = tid / X_R, r = tid % X_R;
                x0c[m * X_RS + r] = (row0 + r < g.n0) ? x0g[(size_t)(c0 + m) * g.ld0 + row0 + r] : 0.0;
            }
            __syncthreads();
            x_accumulate<MODE>(A, B, x0c, X_RS, x1c, pmc, dc, tid);
        }
        x_finish_store<MODE>(A, B, g, tabE, row0, j0);
    }
}

// ---- Linear / Polynomial for d > X_DC: the inner products as an FP64 GEMM on DMMA ------------------
// K = X0 X1^T (+ c, ^q) is C = A B^T with K-contiguous operands, the "row.col" form of
// mma.sync.m8n8k4.f64.  Operands are zero-padded row-major copies (npad x dpad, dpad a multiple of
// XG_CK) built once per handle.  A CTA of 4 warps owns a 128 x 64 tile (a warp 32 x 64, 64
// accumulators per lane), operands stream through a two-stage cp.async pipeline in K-chunks of 16
// with a row stride of 20 doubles (conflict-free fragment loads) -- the GEMM of csrc/pncov_chol.cu
// with a different epilogue.  C_ij and C_ji accumulate the same products in the same order, so a
// symmetric Gram stays symmetric bit for bit.
constexpr int XG_BM = 128, XG_BN = 64, XG_CK = 16, XG_LD = XG_CK + 4, XG_THREADS = 128;

__device__ __forceinline__ void xg_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ void xg_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pn_smem_u32(smem)), "l"(gmem) : "memory");
}

struct XgLaunch {
    const double* A;  // (rows padded to 128, dpad) row-major
    const double* B;  // (cols padded to 128, dpad)
    int dpad;
    int64_t n0, n1;
    double* out;
    int64_t ldo;
    double constant;
    int exponent;
    int vec_ok;
};

__global__ void __launch_bounds__(XG_THREADS, 3) gram_dot_mma_kernel(const __grid_constant__ XgLaunch p) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* As = reinterpret_cast<double*>(smem);  // [2][XG_BM * XG_LD]
    double* Bs = As + 2 * XG_BM * XG_LD;           // [2][XG_BN * XG_LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const double* At = p.A + (int64_t)blockIdx.y * XG_BM * p.dpad;
    const double* Bt = p.B + (int64_t)blockIdx.x * XG_BN * p.dpad;

    auto load_chunk = [&](int kc, int buf) {
        double* as = As + buf * XG_BM * XG_LD;
        double* bs = Bs + buf * XG_BN * XG_LD;
        constexpr int PIECES = XG_CK / 2;
#pragma unroll
        for (int i = 0; i < XG_BM * PIECES / XG_THREADS; ++i) {
            const int idx = tid + i * XG_THREADS;
            const int row = idx / PIECES, c2 = (idx % PIECES) * 2;
            xg_cp_async16(as + row * XG_LD + c2, At + (int64_t)row * p.dpad + kc * XG_CK + c2);
        }
#pragma unroll
        for (int i = 0; i < XG_BN * PIECES / XG_THREADS; ++i) {
            const int idx = tid + i * XG_THREADS;
            const int row = idx / PIECES, c2 = (idx % PIECES) * 2;
            xg_cp_async16(bs + row * XG_LD + c2, Bt + (int64_t)row * p.dpad + kc * XG_CK + c2);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double acc[4][8][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;

    const int nchunk = p.dpad / XG_CK;
    load_chunk(0, 0);
    for (int kc = 0; kc < nchunk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nchunk) {
            load_chunk(kc + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* as = As + buf * XG_BM * XG_LD + (warp * 32 + g) * XG_LD + q;
        const double* bs = Bs + buf * XG_BN * XG_LD + g * XG_LD + q;
#pragma unroll
        for (int k4 = 0; k4 < XG_CK / 4; ++k4) {
            double a[4], b[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = as[r * 8 * XG_LD + k4 * 4];
#pragma unroll
            for (int c = 0; c < 8; ++c) b[c] = bs[c * 8 * XG_LD + k4 * 4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) xg_dmma(acc[r][c][0], acc[r][c][1], a[r], b[c]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t row = (int64_t)blockIdx.y * XG_BM + warp * 32 + r * 8 + g;
        if (row >= p.n0) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int64_t col = (int64_t)blockIdx.x * XG_BN + c * 8 + 2 * q;
            double k0 = acc[r][c][0] + p.constant, k1 = acc[r][c][1] + p.constant;
            if (p.exponent != 1) {
                k0 = x_powi(k0, p.exponent);
                k1 = x_powi(k1, p.exponent);
            }
            double* dst = p.out + row * p.ldo + col;
            if (col + 1 < p.n1) {
                if (p.vec_ok) pn_store_cs2(dst, k0, k1);
                else {
                    pn_store_cs1(dst, k0);
                    pn_store_cs1(dst + 1, k1);
                }
            } else if (col < p.n1) {
                pn_store_cs1(dst, k0);
            }
        }
    }
}

// (n, d) row-major -> (npad, dpad) row-major, zero padded
__global__ void pad_rows_kernel(const double* __restrict__ aos, int64_t n, int d, int64_t npad, int dpad,
                                double* __restrict__ outp) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npad * dpad) return;
    const int64_t i = idx / dpad;
    const int m = (int)(idx % dpad);
    outp[idx] = (i < n && m < d) ? aos[i * d + m] : 0.0;
}

// out[b] = k(x0[b], x1[b]); one thread per pair
template <int MODE>
__global__ void pairs_extra_kernel(const __grid_constant__ PnExtraLaunch g) {
    __shared__ double tabE[PN_EXP2_N];
    if constexpr (MODE == X_PRODMATERN) {
        for (int i = threadIdx.x; i < PN_EXP2_N; i += blockDim.x) tabE[i] = pn_exp2_table_g[i];
        __syncthreads();
    }
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.n0) return;
    const double* a = static_cast<const double*>(g.x0_soa) + b * g.d;  // here: (B, d) row-major
    const double* c = static_cast<const double*>(g.x1_soa) + b * g.d;
    double A = 0.0, B = 1.0;
    for (int m = 0; m < g.d; ++m) {
        if constexpr (MODE == X_DOT) A = pn_fma(a[m], c[m], A);
        else if constexpr (MODE == X_EQUAL) A += (a[m] == c[m]) ? 0.0 : 1.0;
        else {
            const double* pm = static_cast<const double*>(g.pm) + (size_t)m * PN_PM_LD;
            const double r = __dmul_rn(fabs(a[m] - c[m]), pm[0]);
            A = __dadd_rn(A, r);
            B *= x_matern_poly(r, pm);
        }
    }
    double k;
    if constexpr (MODE == X_DOT) k = g.exponent == 1 ? A + g.constant : x_powi(A + g.constant, g.exponent);
    else if constexpr (MODE == X_EQUAL) k = A == 0.0 ? g.constant : 0.0;
    else k = B * pn_exp2(A * -1.4426950408889634074, tabE);
    static_cast<double*>(g.out)[b] = k;
}

// dst = alpha * dst + beta * src (op 0; src may be null when beta == 0) or dst = dst * src (op 1)
template <int OP, bool HAS_SRC>
__global__ void __launch_bounds__(256) combine_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                                       int64_t n2, int64_t n, double alpha, double beta) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 a = reinterpret_cast<double2*>(dst)[i];
        if constexpr (OP == 0) {
            if constexpr (HAS_SRC) {
                const double2 b = __ldcs(reinterpret_cast<const double2*>(src) + i);
                a.x = pn_fma(beta, b.x, alpha * a.x);
                a.y = pn_fma(beta, b.y, alpha * a.y);
            } else {
                a.x *= alpha;
                a.y *= alpha;
            }
        } else {
            const double2 b = __ldcs(reinterpret_cast<const double2*>(src) + i);
            a.x *= b.x;
            a.y *= b.y;
        }
        reinterpret_cast<double2*>(dst)[i] = a;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) {  // odd tail element
        const int64_t i = n - 1;
        if constexpr (OP == 0) dst[i] = HAS_SRC ? pn_fma(beta, src[i], alpha * dst[i]) : alpha * dst[i];
        else dst[i] *= src[i];
    }
}

template <int MODE>
cudaError_t launch_gram_mode(const PnExtraLaunch& g, cudaStream_t stream) {
    const int64_t strips = (g.n1 + X_TN - 1) / X_TN;
    // ~16 CTAs per SM; whole row stages per CTA
    const int64_t unit = g.d <= X_DC ? X_RS : X_R;
    const int64_t n_units = (g.n0 + unit - 1) / unit;
    int64_t upc = (n_units * strips) / ((int64_t)pn_sm_count() * 16);
    upc = upc < 1 ? 1 : (upc > 64 ? 64 : upc);
    const int64_t gy = (n_units + upc - 1) / upc;
    if (gy > 65535) return cudaErrorInvalidConfiguration;
    const size_t smem = 8 * PN_EXP2_N + sizeof(double) * (X_DC * X_TN + X_DC * X_RS + X_DC * PN_PM_LD);
    auto kern = gram_extra_kernel<MODE>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    dim3 grid((unsigned)strips, (unsigned)gy);
    kern<<<grid, X_THREADS, smem, stream>>>(g, (int)(upc * unit));
    ++g_pn_launches;
    return cudaGetLastError();
}

}  // namespace

cudaError_t pn_launch_gram_extra(const PnExtraLaunch& g0, cudaStream_t stream) {
    if (g0.n0 <= 0 || g0.n1 <= 0) return cudaSuccess;
    PnExtraLaunch g = g0;
    g.vec_ok = ((g.ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(g.out) % 16) == 0) ? 1 : 0;
    switch (g.mode) {
        case X_DOT: return launch_gram_mode<X_DOT>(g, stream);
        case X_EQUAL: return launch_gram_mode<X_EQUAL>(g, stream);
        case X_PRODMATERN: return launch_gram_mode<X_PRODMATERN>(g, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t pn_launch_pairs_extra(const PnExtraLaunch& g, cudaStream_t stream) {
    if (g.n0 <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((g.n0 + 255) / 256);
    switch (g.mode) {
        case X_DOT: pairs_extra_kernel<X_DOT><<<blocks, 256, 0, stream>>>(g); break;
        case X_EQUAL: pairs_extra_kernel<X_EQUAL><<<blocks, 256, 0, stream>>>(g); break;
        case X_PRODMATERN: pairs_extra_kernel<X_PRODMATERN><<<blocks, 256, 0, stream>>>(g); break;
        default: return cudaErrorInvalidValue;
    }
    ++g_pn_launches;
    return cudaGetLastError();
}

cudaError_t pn_launch_combine(double* dst, const double* src, int64_t n, int op, double alpha, double beta,
                              cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if (reinterpret_cast<uintptr_t>(dst) % 16 || (src && reinterpret_cast<uintptr_t>(src) % 16))
        return cudaErrorMisalignedAddress;
    const int64_t n2 = n / 2;
    int64_t blocks = (n2 + 255) / 256;
    const int64_t max_blocks = (int64_t)pn_sm_count() * 32;
    blocks = blocks < 1 ? 1 : (blocks > max_blocks ? max_blocks : blocks);
    if (op == 0 && src)
        combine_kernel<0, true><<<(unsigned)blocks, 256, 0, stream>>>(dst, src, n2, n, alpha, beta);
    else if (op == 0)
        combine_kernel<0, false><<<(unsigned)blocks, 256, 0, stream>>>(dst, src, n2, n, alpha, beta);
    else if (op == 1 && src)
        combine_kernel<1, true><<<(unsigned)blocks, 256, 0, stream>>>(dst, src, n2, n, alpha, beta);
    else
        return cudaErrorInvalidValue;
    ++g_pn_launches;
    return cudaGetLastError();
}

cudaError_t pn_launch_pad_rows(const double* aos, int64_t n, int d, int64_t npad, int dpad, double* outp, cudaStream_t stream) {
    const int64_t total = npad * dpad;
    pad_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(aos, n, d, npad, dpad, outp);
    ++g_pn_launches;
    return cudaGetLastError();
}

// K = A B^T + c (^q) with A = rows [row0, row0 + n0) of the padded row operand, B = the padded column operand
cudaError_t pn_launch_gram_dot_mma(const double* A, const double* B, int dpad, int64_t n0, int64_t n1, double* out,
                                   int64_t ldo, double constant, int exponent, cudaStream_t stream) {
    if (n0 <= 0 || n1 <= 0) return cudaSuccess;
    XgLaunch p{};
    p.A = A;
    p.B = B;
    p.dpad = dpad;
    p.n0 = n0;
    p.n1 = n1;
    p.out = out;
    p.ldo = ldo;
    p.constant = constant;
    p.exponent = exponent;
    p.vec_ok = ((ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0) ? 1 : 0;
    const size_t smem = sizeof(double) * 2 * (XG_BM + XG_BN) * XG_LD;
    cudaError_t err = cudaFuncSetAttribute(gram_dot_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const int64_t gx = (n1 + XG_BN - 1) / XG_BN, gy = (n0 + XG_BM - 1) / XG_BM;
    if (gy > 65535) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)gx, (unsigned)gy);
    gram_dot_mma_kernel<<<grid, XG_THREADS, smem, stream>>>(p);
    ++g_pn_launches;
    return cudaGetLastError();
}

// ---- update chains of the matrix-based solver (SURVEY.md 8f, row N2 ii) ---------------------------------
// linalg/solvers/matrixbased.py:751-775 keeps the posterior means / covariance factors as a lazy sum
// of k rank-2 (u v^T + v u^T) or rank-1 (w u^T) LambdaLinearOperators; every product walks k Python
// lambdas.  Here the vectors are rows of two device-resident (capacity, n) matrices U and V and one
// product is two launches: all dot products (fixed split of n, partial sums to a scratch array), then
// y = sum_i U_i a_i + V_i b_i with the coefficients reduced from the partials in a fixed order.
namespace {
constexpr int R2_THREADS = 256;

// partial[((z * k) + i) * splits + s] = sum_{j in slice s} W_i[j] x[j],  W = (z == 0 ? U : V)
__global__ void __launch_bounds__(R2_THREADS) chain_dots_kernel(const double* __restrict__ U, const double* __restrict__ V,
                                                                int64_t ld, int64_t n, int k, const double* __restrict__ x,
                                                                double* __restrict__ partial, int splits) {
    __shared__ double red[R2_THREADS / 32];
    const int s = blockIdx.x, i = blockIdx.y, z = blockIdx.z;
    const double* w = (z == 0 ? U : V) + (int64_t)i * ld;
    const int64_t chunk = (n + splits - 1) / splits;
    const int64_t j0 = (int64_t)s * chunk, j1 = (j0 + chunk < n) ? j0 + chunk : n;
    double acc = 0.0;
    for (int64_t j = j0 + threadIdx.x; j < j1; j += R2_THREADS) acc = pn_fma(w[j], x[j], acc);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < R2_THREADS / 32; ++q) t += red[q];
        partial[((int64_t)z * k + i) * splits + s] = t;
    }
}

// mode 0: y = sum_i U_i (V_i . x) + V_i (U_i . x);  mode 1: y = sum_i V_i (U_i . x);  mode 2: y = sum_i U_i (V_i . x)
__global__ void __launch_bounds__(R2_THREADS) chain_apply_kernel(const double* __restrict__ U, const double* __restrict__ V,
                                                                 int64_t ld, int64_t n, int k, const double* __restrict__ partial,
                                                                 int splits, int mode, double* __restrict__ y) {
    extern __shared__ double coef[];  // [k] dots with U, [k] dots with V
    for (int t = threadIdx.x; t < 2 * k; t += R2_THREADS) {
        double sum = 0.0;
        for (int s = 0; s < splits; ++s) sum += partial[(int64_t)t * splits + s];
        coef[t] = sum;
    }
    __syncthreads();
    const int64_t j = ((int64_t)blockIdx.x * R2_THREADS + threadIdx.x) * 2;
    if (j >= n) return;
    const bool pair = j + 1 < n;
    double y0 = 0.0, y1 = 0.0;
    for (int i = 0; i < k; ++i) {
        const double du = coef[i], dv = coef[k + i];
        const double* u = U + (int64_t)i * ld + j;
        const double* v = V + (int64_t)i * ld + j;
        if (mode != 1) {  // U_i (V_i . x)
            y0 = pn_fma(u[0], dv, y0);
            if (pair) y1 = pn_fma(u[1], dv, y1);
        }
        if (mode != 2) {  // V_i (U_i . x)
            y0 = pn_fma(v[0], du, y0);
            if (pair) y1 = pn_fma(v[1], du, y1);
        }
    }
    y[j] = y0;
    if (pair) y[j + 1] = y1;
}
}  // namespace

// x, y: device vectors of length n; partial: scratch of 2 * k * splits doubles
cudaError_t pn_launch_chain_apply(const double* U, const double* V, int64_t ld, int64_t n, int k, int mode, const double* x,
                                  double* y, double* partial, int splits, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if (k <= 0) return cudaMemsetAsync(y, 0, (size_t)n * 8, stream);
    chain_dots_kernel<<<dim3((unsigned)splits, (unsigned)k, 2), R2_THREADS, 0, stream>>>(U, V, ld, n, k, x, partial, splits);
    ++g_pn_launches;
    const unsigned blocks = (unsigned)((n + 2 * R2_THREADS - 1) / (2 * R2_THREADS));
    const size_t coef_bytes = (size_t)2 * k * 8;  // above 48 KB (k > 3072) the launch needs the opt-in attribute
    if (coef_bytes > 48 * 1024) {
        if (coef_bytes > 200 * 1024) return cudaErrorInvalidValue;  // pncov_chain_create caps the capacity accordingly
        cudaError_t err = cudaFuncSetAttribute(chain_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coef_bytes);
        if (err != cudaSuccess) return err;
    }
    chain_apply_kernel<<<blocks, R2_THREADS, coef_bytes, stream>>>(U, V, ld, n, k, partial, splits, mode, y);
    ++g_pn_launches;
    return cudaGetLastError();
}
