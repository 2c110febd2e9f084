// C ABI of libpncov (see include/pncov.h): handle management, host<->device staging,
// argument validation and the small utility kernels (layout changes, packing, the
// elementwise "pairs" evaluation).  The two hot kernels live in pncov_gram.cu and
// pncov_matvec.cu.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sys/mman.h>

#include <algorithm>
#include <map>
#include <chrono>
#include <functional>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include "../../include/pncov.h"
#include "pncov_internal.h"

int64_t g_pn_launches = 0;
static int64_t g_pn_d2h_bytes = 0;  // bytes pncov_gram moved over PCIe (a symmetric result may cross as its upper blocks only)

int pn_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cached[dev] = sms;
    }
    return cached[dev];
}

namespace {

thread_local std::string t_err;

// Symmetric host return path: mirror on the host or copy everything?  Halving the PCIe bytes pays when the link
// is the limit (one GPU per host: 157 -> 119 ms per 8.6 GB matrix) and loses when the host's memory bandwidth
// is (eight ranks sharing one socket: the mirror reads and writes what the DMA engine would only write).  The
// library measures both on the process's first two large symmetric results and keeps the faster, re-probing the
// other one every 16th call.
struct MirrorPolicy {
    std::mutex m;
    double rate[2] = {0.0, 0.0};  // result bytes per second: [0] full copy, [1] mirrored
    int calls = 0;
    bool choose() {
        std::lock_guard<std::mutex> lk(m);
        ++calls;
        if (rate[1] == 0.0) return true;
        if (rate[0] == 0.0) return false;
        const bool best = rate[1] >= rate[0];
        return (calls % 16 == 0) ? !best : best;
    }
    void record(bool mirrored, double bytes, double seconds) {
        if (seconds <= 0) return;
        std::lock_guard<std::mutex> lk(m);
        rate[mirrored ? 1 : 0] = bytes / seconds;
    }
} g_mirror_policy;

int fail(int code, const std::string& msg) {
    t_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char* what) {
    t_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    cudaGetLastError();  // clear sticky-less error state
    return (int)e > 0 ? (int)e : 1;
}

#define PN_CUDA(expr)                                         \
    do {                                                      \
        cudaError_t pn_e_ = (expr);                           \
        if (pn_e_ != cudaSuccess) return fail_cuda(pn_e_, #expr); \
    } while (0)

constexpr double kLog2e = 1.4426950408889634074;

// Device scratch comes from CUDA's stream-ordered pool so that creating / destroying an
// operator does not pay cudaMalloc / cudaFree (hundreds of microseconds each, plus a device
// synchronisation) every time: small Gram calls are otherwise dominated by that overhead.
// The pool keeps up to 4 GiB cached per device.  Allocation is synchronous from the caller's
// point of view (the buffer may be used on any stream afterwards); frees are issued after the
// owning handle has drained its work.
cudaError_t pn_dev_alloc(void** p, size_t bytes, cudaStream_t s) {
    static bool configured[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = 16ull << 30;  // of 180 GB: factor buffers, staging blocks and operands of repeated calls are reused
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        configured[dev] = true;
    }
    e = cudaMallocAsync(p, bytes ? bytes : 1, s);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(s);
}
void pn_dev_free(void* p, cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
}

// ---- utility kernels ------------------------------------------------------------------------
// (n, d) row-major -> (d, ld) coordinate-major; pad columns [n, ld) are left as zero-filled
template <typename T>
__global__ void aos_to_soa_kernel(const T* __restrict__ aos, T* __restrict__ soa, int64_t n, int d,
                                  int64_t ld) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d) return;
    const int64_t i = idx / d;
    const int m = (int)(idx % d);
    soa[(size_t)m * ld + i] = aos[idx];
}

// V (rows, k) with leading dimension ldv -> packed (rows_pad, kc) chunk starting at column c0
template <typename T>
__global__ void pack_v_kernel(const T* __restrict__ v, int64_t ldv, int64_t rows, int64_t rows_pad,
                              int64_t k, int64_t c0, int kc, int vstride, T* __restrict__ vp) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows_pad * vstride) return;
    const int64_t j = idx / vstride;
    const int c = (int)(idx % vstride);
    vp[idx] = (j < rows && c < kc && c0 + c < k) ? v[j * ldv + c0 + c] : T(0);
}

template <typename T>
__global__ void unpack_y_kernel(const T* __restrict__ yp, int64_t rows, int kc, int64_t k, int64_t c0,
                                T* __restrict__ y, int64_t ldy) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * kc) return;
    const int64_t i = idx / kc;
    const int c = (int)(idx % kc);
    if (c0 + c < k) y[i * ldy + c0 + c] = yp[idx];
}

// factorised ExpQuad product: v'_j = f_j v_j on the way in, y_i = f_i y'_i on the way out, with
// f = slot 9 of the 12-double operand rows built by build_mma_operands_kernel
__global__ void pack_v_scaled_kernel(const double* __restrict__ v, int64_t ldv, int64_t rows, int64_t rows_pad,
                                     int64_t k, int64_t c0, int kc, int vstride, const double* __restrict__ op12, int ld, int fi,
                                     double* __restrict__ vp) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows_pad * vstride) return;
    const int64_t j = idx / vstride;
    const int c = (int)(idx % vstride);
    vp[idx] = (j < rows && c < kc && c0 + c < k) ? v[j * ldv + c0 + c] * op12[j * ld + fi] : 0.0;
}
__global__ void unpack_y_scaled_kernel(const double* __restrict__ yp, int64_t rows, int kc, int64_t k, int64_t c0,
                                       const double* __restrict__ op12, int ld, int fi, double* __restrict__ y,
                                       int64_t ldy) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * kc) return;
    const int64_t i = idx / kc;
    const int c = (int)(idx % kc);
    if (c0 + c < k) y[i * ldy + c0 + c] = yp[idx] * op12[i * ld + fi];
}

// per-coordinate sums of a point set, one block, fixed summation order (deterministic centre)
__global__ void coord_sums_kernel(const double* __restrict__ aos, int64_t n, int d, double* __restrict__ sums) {
    __shared__ double red[1024];
    for (int m = 0; m < d; ++m) {
        double s = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += aos[i * d + m];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int w = blockDim.x / 2; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) sums[m] = red[0];
        __syncthreads();
    }
}

// centred + scaled operands of the norm-identity matvec: xt = (x - c) * s, `ld` = dpad + 4 doubles per point
//   a[i] = [xt_0..xt_{dpad-1}, |xt|^2, f, 0, 0]      b[i] = [-2 xt_0..-2 xt_{dpad-1}, |xt|^2, f, 0, 0]
// with f = exp(-|xt|^2) for ExpQuad (the outer factors of the factorised product, PN_EPI_EXPQUAD_FACT)
// and 0 otherwise; coordinates m >= d and rows [n, npad) are zero.  max_norm receives max_i |xt_i|^2
// (bit pattern of a non-negative double, so an integer atomicMax is exact and order independent).
__global__ void build_mma_operands_kernel(const double* __restrict__ aos, int64_t n, int64_t npad, int d, int dpad,
                                          const double* __restrict__ centre, const double* __restrict__ scale,
                                          double iso_scale, double* __restrict__ a12, double* __restrict__ b12,
                                          unsigned long long* __restrict__ max_norm, int expquad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    const int ld = dpad + 4;
    double* pa = a12 + i * ld;
    double* pb = b12 + i * ld;
    double nrm = 0.0;
    for (int m = 0; m < dpad; ++m) {
        double v = 0.0;
        if (i < n && m < d) v = (aos[i * d + m] - centre[m]) * (scale ? scale[m] : iso_scale);
        pa[m] = v;
        pb[m] = -2.0 * v;
        nrm = fma(v, v, nrm);
    }
    pa[dpad] = nrm;
    pb[dpad] = nrm;
    const double ef = expquad ? exp(-nrm) : 0.0;
    pa[dpad + 1] = ef;
    pb[dpad + 1] = ef;
    pa[dpad + 2] = pa[dpad + 3] = 0.0;
    pb[dpad + 2] = pb[dpad + 3] = 0.0;
    if (i < n) atomicMax(max_norm, (unsigned long long)__double_as_longlong(nrm));
}

// out[b] = k(x0[b], x1[b]); one thread per pair (O(B d) work, not a hot path)
template <typename T, typename TOut>
__global__ void pairs_kernel(const __grid_constant__ PnPairsLaunch p) {
    __shared__ double tabE[PN_EXP2_N];
    __shared__ double tabL[2 * PN_LOG2_N];
    pn_stage_tables(tabE, tabL, threadIdx.x, blockDim.x);
    __syncthreads();
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.b) return;
    const T* a = static_cast<const T*>(p.x0) + b * p.d;
    const T* c = static_cast<const T*>(p.x1) + b * p.d;
    const T* sc = static_cast<const T*>(p.scale);
    T S = T(0);
    for (int m = 0; m < p.d; ++m) {
        T df = a[m] - c[m];
        if (sc) df *= sc[m];
        S += df * df;
    }
    TOut* out = static_cast<TOut*>(p.out);
    if constexpr (sizeof(TOut) == 4) {
        out[b] = exp2f(S * (float)p.epi.c_u);
    } else {
        out[b] = pn_epilogue((double)S, p.epi, tabE, tabL);
    }
}

// ---- kernel description -----------------------------------------------------------------------
struct KernelDesc {
    int kind = 0, dtype = 0, d = 0;
    bool ard = false;
    std::vector<double> scale;  // per-dimension scale factors when ard
    PnEpiParams epi{};
    // non-isotropic families (pncov_extra.cu): mode >= 0 selects them
    int extra_mode = -1;
    double constant = 0.0;   // Linear / Polynomial: c; WhiteNoise: sigma^2
    int exponent = 1;        // Polynomial
    std::vector<double> pm;  // ProductMatern: (d, PN_PM_LD) per-dimension {scale, p, c_0..c_8}
    bool gram_f32() const { return dtype == PNCOV_F32 && kind == PNCOV_EXPQUAD; }
    size_t elem() const { return dtype == PNCOV_F32 ? 4 : 8; }
    size_t gram_elem() const { return gram_f32() ? 4 : 8; }
};

unsigned long long gcd_u(unsigned long long a, unsigned long long b) {
    while (b) {
        unsigned long long t = a % b;
        a = b;
        b = t;
    }
    return a;
}

// exact rational recursion of Matern.half_integer_coefficients (_matern.py:225-267);
// coef[i] multiplies r^i
bool matern_coefficients(int p, double* coef) {
    unsigned long long num = 1, den = 1;
    coef[0] = 1.0;
    int idx = 1;
    for (int i = p - 1; i >= 0; --i, ++idx) {
        unsigned long long a = 2ULL * (unsigned long long)(i + 1);
        unsigned long long b = (unsigned long long)(p + i + 1) * (unsigned long long)(p - i);
        unsigned long long g = gcd_u(a, b);
        a /= g;
        b /= g;
        g = gcd_u(num, b);
        num /= g;
        b /= g;
        g = gcd_u(a, den);
        a /= g;
        den /= g;
        if (num > (1ULL << 53) / a || den > (1ULL << 53) / b) return false;
        num *= a;
        den *= b;
        coef[idx] = (double)num / (double)den;  // both exact in double: one correctly rounded division
    }
    return true;
}

unsigned hi_word(double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    return (unsigned)(b >> 32);
}

// Window of S (as hi words) in which PnEpilogue::fast is valid; see pncov_common.cuh.
void fill_window(PnEpiParams& P) {
    const double u_max = 1021.0;      // |exponent argument| bound: results stay normal
    double s_max = ldexp(1.0, 1000);  // generic cap
    double s_min = 0.0;
    switch (P.epi) {
        case PN_EPI_EXPQUAD:
            s_max = fmin(s_max, u_max / fabs(P.c_u));
            break;
        case PN_EPI_MATERN0:
        case PN_EPI_MATERN1:
        case PN_EPI_MATERN2:
        case PN_EPI_MATERNP: {
            const double r_max = u_max / fabs(P.c_su);
            s_max = fmin(s_max, r_max * r_max);
            s_min = ldexp(1.0, -1000);  // sqrt_core needs a normal, non-zero argument
            break;
        }
        default: {
            double b_log2_max = 990.0;  // rcp_core / log2_core need b < 2^994
            if (P.epi == PN_EPI_RATQUADG) b_log2_max = fmin(b_log2_max, u_max / P.alpha);
            // b = 1 + c S < 2^b_log2_max
            const double b_max = exp2(b_log2_max);
            s_max = fmin(s_max, (b_max - 1.0) / P.c_s * (1.0 - 1e-12));
            break;
        }
    }
    P.s_hi_word = hi_word(s_max);                       // flag when hi(S) >= hi(s_max)
    P.s_lo_word = s_min > 0.0 ? hi_word(s_min) + 1u : 0u;  // flag when hi(S) < this
}

int build_desc(KernelDesc& k, int kind, int dtype, int64_t d, const double* ls, int64_t nls, double nu,
               double alpha) {
    if (kind < PNCOV_EXPQUAD || kind > PNCOV_PRODUCT_MATERN) return fail(PNCOV_EINVAL, "unknown kernel kind");
    if (dtype != PNCOV_F64 && dtype != PNCOV_F32) return fail(PNCOV_EINVAL, "unknown dtype");
    if (d < 1 || d > (1 << 20)) return fail(PNCOV_EINVAL, "input dimension d must be in [1, 2^20]");
    if (kind >= PNCOV_LINEAR) {
        // _linear.py:44-50, _polynomial.py:50-56, _white_noise.py:38-55, _product_matern.py:60-122
        if (dtype != PNCOV_F64) return fail(PNCOV_EINVAL, "Linear / Polynomial / WhiteNoise / ProductMatern take float64 points");
        k.kind = kind;
        k.dtype = dtype;
        k.d = (int)d;
        k.ard = false;
        k.epi = PnEpiParams{};
        k.constant = alpha;
        k.exponent = 1;
        if (kind == PNCOV_LINEAR) {
            k.extra_mode = 0;
        } else if (kind == PNCOV_POLYNOMIAL) {
            if (!(nu >= 0 && nu <= 64 && nu == floor(nu)))
                return fail(PNCOV_ENOTIMPL, "Polynomial exponent must be an integer in [0, 64] on the device path");
            k.extra_mode = 0;
            k.exponent = (int)nu;
        } else if (kind == PNCOV_WHITENOISE) {
            k.extra_mode = 1;
        } else {
            if (!ls || nls != 2 * d) return fail(PNCOV_EINVAL, "ProductMatern takes d lengthscales followed by d smoothness values");
            k.extra_mode = 2;
            k.pm.assign((size_t)d * PN_PM_LD, 0.0);
            for (int64_t m = 0; m < d; ++m) {
                const double l = ls[m], nu_m = ls[d + m];
                if (!(l > 0)) return fail(PNCOV_EINVAL, "lengthscales must be positive");
                if (!(nu_m > 0)) return fail(PNCOV_EINVAL, "nu must be positive");
                const double pm_ = nu_m - 0.5;
                if (pm_ != floor(pm_) || pm_ > PNCOV_MAX_MATERN_P)
                    return fail(PNCOV_ENOTIMPL, "ProductMatern needs half-integer nu <= PNCOV_MAX_MATERN_P + 1/2 in every dimension");
                double* row = &k.pm[(size_t)m * PN_PM_LD];
                row[0] = sqrt(2 * nu_m) / l;  // _matern.py:171-173
                row[1] = pm_;
                if (!matern_coefficients((int)pm_, row + 2)) return fail(PNCOV_ENOTIMPL, "Matern coefficients overflow");
            }
        }
        return 0;
    }
    if (!ls || (nls != 1 && nls != d)) return fail(PNCOV_EINVAL, "n_lengthscales must be 1 or d");
    for (int64_t i = 0; i < nls; ++i)
        if (!(ls[i] > 0)) return fail(PNCOV_EINVAL, "lengthscales must be positive");
    k.kind = kind;
    k.dtype = dtype;
    k.d = (int)d;
    k.ard = nls > 1;
    k.epi = PnEpiParams{};
    pn_fill_math_constants(k.epi);
    double base = 1.0;  // numerator of the scale factor
    if (kind == PNCOV_EXPQUAD) {
        base = sqrt(0.5);  // _exponentiated_quadratic.py:85-87
        k.epi.epi = PN_EPI_EXPQUAD;
    } else if (kind == PNCOV_MATERN) {
        if (!(nu > 0)) return fail(PNCOV_EINVAL, "nu must be positive");
        const double pm = nu - 0.5;
        if (pm != floor(pm)) return fail(PNCOV_ENOTIMPL, "Matern with non-half-integer nu (Bessel branch) is outside the accelerated scope");
        if (pm > PNCOV_MAX_MATERN_P) return fail(PNCOV_ENOTIMPL, "Matern degree p = nu - 1/2 above PNCOV_MAX_MATERN_P");
        const int p = (int)pm;
        base = sqrt(2 * nu);  // _matern.py:171-173
        k.epi.p = p;
        k.epi.epi = p == 0 ? PN_EPI_MATERN0 : p == 1 ? PN_EPI_MATERN1 : p == 2 ? PN_EPI_MATERN2 : PN_EPI_MATERNP;
        if (!matern_coefficients(p, k.epi.coef)) return fail(PNCOV_ENOTIMPL, "Matern coefficients overflow");
    } else {
        if (nls != 1) return fail(PNCOV_EINVAL, "RatQuad takes a single lengthscale");
        if (!(alpha > 0)) return fail(PNCOV_EINVAL, "alpha must be positive");
        k.epi.c_s = 1.0 / (2.0 * alpha * (ls[0] * ls[0]));  // _rational_quadratic.py:78-84
        k.epi.alpha = alpha;
        k.epi.epi = alpha == 1.0 ? PN_EPI_RATQUAD1 : alpha == 2.0 ? PN_EPI_RATQUAD2 : PN_EPI_RATQUADG;
        k.ard = false;
        fill_window(k.epi);
        return 0;
    }
    if (k.ard) {
        k.scale.resize(d);
        for (int64_t i = 0; i < d; ++i) k.scale[i] = base / ls[i];
        k.epi.c_u = -kLog2e;
        k.epi.c_s = 1.0;
    } else {
        const double s = base / ls[0];
        if (kind == PNCOV_EXPQUAD) {
            k.epi.c_u = -kLog2e * (s * s);
            k.epi.c_s = 1.0;
        } else {
            k.epi.c_u = -kLog2e;
            k.epi.c_s = s;
        }
    }
    k.epi.c_su = k.epi.c_s * k.epi.c_u;
    if (kind == PNCOV_MATERN) {
        // the device evaluates the polynomial in r0 = sqrt(S) (r = c_s r0): fold c_s^i into c_i
        long double pw = 1.0L;
        for (int i = 1; i <= k.epi.p; ++i) {
            pw *= (long double)k.epi.c_s;
            k.epi.coef[i] = (double)((long double)k.epi.coef[i] * pw);
        }
    }
    fill_window(k.epi);
    return 0;
}

int64_t pad_up(int64_t n) { return ((n + PN_PAD - 1) / PN_PAD) * PN_PAD; }

struct Points {
    int64_t n = 0, npad = 0;
    void* aos = nullptr;  // (npad, d)
    void* soa = nullptr;  // (d, npad)
};

}  // namespace

struct pncov_handle {
    int device = 0;
    int sm_count = 0;  // set at creation from the device
    KernelDesc k;
    Points p0, p1;
    bool symmetric = false;  // x1 was NULL: p1 aliases p0
    void* scale_dev = nullptr;
    void* pm_dev = nullptr;  // ProductMatern per-dimension parameters
    void* dot_ops[2] = {nullptr, nullptr};  // Linear / Polynomial, d > 16: zero-padded row-major operands (x0, x1)
    int dot_dpad = 0;
    // ExpQuad / RatQuad, d > 16: operands of the guarded norm-identity DMMA Gram (pncov_gram_mma.cu), per point set
    int gm_dpad = 0;  // > 0: that kernel serves the dense Gram of this handle
    void* gm_A[2] = {nullptr, nullptr};
    void* gm_B[2] = {nullptr, nullptr};
    void* gm_nrm[2] = {nullptr, nullptr};
    void* gm_thr[2] = {nullptr, nullptr};
    PnEpiParams gm_epi{};
    unsigned gm_v_hi = 0;
    // covariance arithmetic in one launch (pncov_set_program): the dense Gram of this handle evaluates `prog`
    bool has_prog = false;
    PnFusedProgram prog{};
    bool prog_simple = false;  // the program is `[o *] k [+ WhiteNoise]`: prog_epi on the specialised kernels
    PnEpiParams prog_epi{};
    // kernel-selection switches of the environment, read ONCE when the handle is created
    int env_ident = -1;   // PNCOV_MATVEC_IDENT: -1 unset, 0 "off", 1 "dfma"
    int env_fact = -1;    // PNCOV_MATVEC_FACT: 0 "off"
    int env_sym = -1;     // PNCOV_MATVEC_SYM: 0 "off", 1 "on"
    int env_mirror = -1;  // PNCOV_GRAM_MIRROR: 0 "off", 1 "on" (always), unset: measured policy
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    // scratch (grown on demand)
    void* vpack = nullptr;
    size_t vpack_bytes = 0;
    void* ypack = nullptr;
    size_t ypack_bytes = 0;
    void* partial = nullptr;
    size_t partial_bytes = 0;
    void* bins = nullptr;  // symmetric-halving products: integer planes (2 n0 int64 + n0 doubles) and max |v|
    size_t bins_bytes = 0;
    void* vmax_dev = nullptr;
    void* stage = nullptr;  // device staging for host entry points
    size_t stage_bytes = 0;
    bool used_foreign_stream = false;  // a *_device entry point ran on a caller-owned stream
    // norm-identity / DMMA matvec operands (built on the first product; see mma_prepare)
    int mma_state = 0;           // 0 = not examined, 1 = ready, -1 = not admissible for this data
    int mma_dpad = 8;            // coordinates per operand row (d padded to 8, 16, 32 or 64)
    bool last_sym = false;       // the last product ran the symmetric-halving kernel
    bool last_fact = false;      // ... in the factorised ExpQuad form (row factors applied when the planes are finished)
    int last_use_mma = -1;       // kernel family of the last product (-1: none yet)
    // per-handle kernel selection (pncov_set_option); -1 = default (environment / automatic)
    int opt_identity = -1, opt_symmetric = -1, opt_factorised = -1;
    void* a12[2] = {nullptr, nullptr};  // per point set (0: x0, 1: x1); aliases when symmetric
    void* b12[2] = {nullptr, nullptr};
    PnEpiParams mma_epi{};
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int ensure(void** buf, size_t* have, size_t need, cudaStream_t s) {
    if (*have >= need) return 0;
    if (*buf) {
        // earlier launches that used the old buffer ran on the caller's stream, not on `s`
        PN_CUDA(cudaDeviceSynchronize());
        pn_dev_free(*buf, s);
    }
    *buf = nullptr;
    *have = 0;
    PN_CUDA(pn_dev_alloc(buf, need, s));
    *have = need;
    return 0;
}

int upload_points(pncov_handle* h, Points& P, const void* x, int64_t n, bool on_device) {
    const size_t es = h->k.elem();
    const int d = h->k.d;
    P.n = n;
    P.npad = pad_up(n > 0 ? n : 1);
    const size_t bytes = (size_t)P.npad * d * es;
    PN_CUDA(pn_dev_alloc(&P.aos, bytes, h->stream));
    PN_CUDA(pn_dev_alloc(&P.soa, bytes, h->stream));
    PN_CUDA(cudaMemsetAsync(P.aos, 0, bytes, h->stream));
    PN_CUDA(cudaMemsetAsync(P.soa, 0, bytes, h->stream));
    if (n > 0) {
        PN_CUDA(cudaMemcpyAsync(P.aos, x, (size_t)n * d * es, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        const int64_t total = n * d;
        const unsigned blocks = (unsigned)((total + 255) / 256);
        if (es == 8)
            aos_to_soa_kernel<double><<<blocks, 256, 0, h->stream>>>((const double*)P.aos, (double*)P.soa, n, d, P.npad);
        else
            aos_to_soa_kernel<float><<<blocks, 256, 0, h->stream>>>((const float*)P.aos, (float*)P.soa, n, d, P.npad);
        ++g_pn_launches;
        PN_CUDA(cudaGetLastError());
    }
    return 0;
}

void free_points(Points& P, cudaStream_t s) {
    pn_dev_free(P.aos, s);
    pn_dev_free(P.soa, s);
    P = Points{};
}

int upload_scale(const KernelDesc& k, void** dev, cudaStream_t stream) {
    *dev = nullptr;
    if (!k.ard) return 0;
    PN_CUDA(pn_dev_alloc(dev, k.scale.size() * k.elem(), stream));
    if (k.dtype == PNCOV_F64) {
        PN_CUDA(cudaMemcpyAsync(*dev, k.scale.data(), k.scale.size() * 8, cudaMemcpyHostToDevice, stream));
        PN_CUDA(cudaStreamSynchronize(stream));
    } else {
        std::vector<float> s32(k.scale.begin(), k.scale.end());
        PN_CUDA(cudaMemcpyAsync(*dev, s32.data(), s32.size() * 4, cudaMemcpyHostToDevice, stream));
        PN_CUDA(cudaStreamSynchronize(stream));
    }
    return 0;
}

PnGramLaunch gram_launch(const pncov_handle* h, bool transpose, int64_t row0, int64_t rows, void* out, int64_t ld) {
    const Points& R = transpose ? h->p1 : h->p0;  // rows of the output
    const Points& C = transpose ? h->p0 : h->p1;
    const size_t es = h->k.elem();
    PnGramLaunch g{};
    g.x0_soa = static_cast<const char*>(R.soa) + (size_t)row0 * es;
    g.x1_soa = C.soa;
    g.ld0 = R.npad;
    g.ld1 = C.npad;
    g.n0 = rows;
    g.n1 = C.n;
    g.scale = h->scale_dev;
    g.d = h->k.d;
    g.out = out;
    g.ldo = ld;
    g.in_f32 = h->k.dtype == PNCOV_F32;
    g.out_f32 = h->k.gram_f32();
    g.symmetric = (h->symmetric && row0 == 0 && rows == R.n) ? 1 : 0;
    g.sm_count = h->sm_count;
    g.epi = h->k.epi;
    return g;
}

int upload_pm(const KernelDesc& k, void** dev, cudaStream_t stream) {
    *dev = nullptr;
    if (k.pm.empty()) return 0;
    PN_CUDA(pn_dev_alloc(dev, k.pm.size() * 8, stream));
    PN_CUDA(cudaMemcpyAsync(*dev, k.pm.data(), k.pm.size() * 8, cudaMemcpyHostToDevice, stream));
    PN_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

// one Gram launch (rows [row0, row0 + rows) of K or K^T) of whichever family the handle holds
cudaError_t launch_gram_any(const pncov_handle* h, bool transpose, int64_t row0, int64_t rows, void* out, int64_t ld,
                            cudaStream_t stream) {
    const Points& R = transpose ? h->p1 : h->p0;
    const Points& C = transpose ? h->p0 : h->p1;
    if (h->has_prog && h->prog_simple) {
        PnGramLaunch g = gram_launch(h, transpose, row0, rows, out, ld);
        g.epi = h->prog_epi;
        return pn_launch_gram(g, stream);
    }
    if (h->has_prog) return pn_launch_gram_fused(gram_launch(h, transpose, row0, rows, out, ld), h->prog, stream);
    if (h->gm_dpad > 0) {  // ExpQuad / RatQuad with wide inputs: guarded norm identity on DMMA
        const int r = transpose ? 1 : 0, c = transpose ? 0 : 1;
        PnGramMmaLaunch p{};
        p.A = static_cast<const double*>(h->gm_A[r]) + (size_t)row0 * h->gm_dpad;
        p.B = static_cast<const double*>(h->gm_B[c]);
        p.nrmA = static_cast<const double*>(h->gm_nrm[r]) + row0;
        p.nrmB = static_cast<const double*>(h->gm_nrm[c]);
        p.thrA = static_cast<const int*>(h->gm_thr[r]) + row0;
        p.thrB = static_cast<const int*>(h->gm_thr[c]);
        p.x0_aos = static_cast<const double*>(R.aos) + (size_t)row0 * h->k.d;
        p.x1_aos = static_cast<const double*>(C.aos);
        p.scale = static_cast<const double*>(h->scale_dev);
        p.d = h->k.d;
        p.dpad = h->gm_dpad;
        p.n0 = rows;
        p.n1 = C.n;
        p.row0 = row0;
        p.same_points = h->symmetric ? 1 : 0;
        p.out = static_cast<double*>(out);
        p.ldo = ld;
        p.v_hi = h->gm_v_hi;
        p.epi = h->gm_epi;
        p.epi_direct = h->k.epi;
        return pn_launch_gram_mma(p, stream);
    }
    if (h->k.extra_mode < 0) return pn_launch_gram(gram_launch(h, transpose, row0, rows, out, ld), stream);
    if (h->dot_dpad > 0) {  // Linear / Polynomial with wide inputs: FP64 GEMM on DMMA
        const double* A = static_cast<const double*>(h->dot_ops[transpose ? 1 : 0]) + (size_t)row0 * h->dot_dpad;
        const double* B = static_cast<const double*>(h->dot_ops[transpose ? 0 : 1]);
        return pn_launch_gram_dot_mma(A, B, h->dot_dpad, rows, C.n, static_cast<double*>(out), ld, h->k.constant,
                                      h->k.exponent, stream);
    }
    PnExtraLaunch g{};
    g.mode = h->k.extra_mode;
    g.x0_soa = static_cast<const char*>(R.soa) + (size_t)row0 * 8;
    g.x1_soa = C.soa;
    g.ld0 = R.npad;
    g.ld1 = C.npad;
    g.n0 = rows;
    g.n1 = C.n;
    g.d = h->k.d;
    g.pm = h->pm_dev;
    g.constant = h->k.constant;
    g.exponent = h->k.exponent;
    g.out = out;
    g.ldo = ld;
    if (h->k.extra_mode == 2 && h->symmetric && row0 == 0 && rows == R.n && h->k.d <= 8) {
        // whole symmetric ProductMatern Gram: upper triangle + mirrored stores on the 64 grid, general kernel on the edges
        int64_t ni = 0;
        cudaError_t err = pn_launch_gram_pm_sym(gram_launch(h, transpose, row0, rows, out, ld), h->k.pm.data(), stream, &ni);
        if (err != cudaSuccess) return err;
        if (ni > 0) {
            if (ni == g.n1) return cudaSuccess;
            PnExtraLaunch e = g;  // right edge: all rows, remaining columns
            e.x1_soa = static_cast<const char*>(C.soa) + (size_t)ni * 8;
            e.n1 = g.n1 - ni;
            e.out = static_cast<double*>(out) + ni;
            if ((err = pn_launch_gram_extra(e, stream)) != cudaSuccess) return err;
            e = g;  // bottom edge: remaining rows, interior columns
            e.x0_soa = static_cast<const char*>(R.soa) + (size_t)ni * 8;
            e.n0 = g.n0 - ni;
            e.n1 = ni;
            e.out = static_cast<double*>(out) + ni * ld;
            return pn_launch_gram_extra(e, stream);
        }
    }
    return pn_launch_gram_extra(g, stream);
}

// Operands of the guarded norm-identity Gram (pncov_gram_mma.cu) for ExpQuad / RatQuad on wide inputs.  Unlike the
// matrix-free product (mma_prepare below) nothing is decided per handle: the guard is per ENTRY inside the kernel.
constexpr int64_t kGmRowSlack = 128;  // a launch may start at any row: its last 128-row tile may read past npad
int gram_mma_prepare(pncov_handle* h) {
    const KernelDesc& k = h->k;
    if (k.dtype != PNCOV_F64 || k.d <= 16 || k.d > 4096 || (k.kind != PNCOV_EXPQUAD && k.kind != PNCOV_RATQUAD)) return 0;
    if (h->p0.n == 0 || h->p1.n == 0) return 0;
    {
        const char* env = getenv("PNCOV_GRAM_IDENT");  // "off": direct differences at every d (read once per handle)
        if (env && strcmp(env, "off") == 0) return 0;
    }
    cudaStream_t stream = h->stream;
    const int d = k.d, dpad = ((d + 15) / 16) * 16;
    const bool expq = k.kind == PNCOV_EXPQUAD;
    double* centre = nullptr;
    PN_CUDA(pn_dev_alloc(reinterpret_cast<void**>(&centre), (size_t)d * sizeof(double), stream));
    const double ntot = (double)h->p0.n + (h->symmetric ? 0.0 : (double)h->p1.n);
    cudaError_t e = pn_launch_gram_mma_centre((const double*)h->p0.soa, h->p0.npad, h->symmetric ? nullptr : (const double*)h->p1.soa,
                                              h->p1.npad, ntot, d, centre, stream);
    // ExpQuad: operands carry the scale factors, the exponent constant becomes -log2(e); RatQuad: unscaled
    double iso = 1.0;
    if (expq && !k.ard) iso = sqrt(-k.epi.c_u / kLog2e);
    const double* scale_dev = (expq && k.ard) ? (const double*)h->scale_dev : nullptr;
    const double gamma = (4.0 * d + 24.0) * ldexp(1.0, -53);
    const double gam_over_tol = (expq ? 1.0 : k.epi.alpha * k.epi.c_s) * gamma / 2e-13;
    const double dup_factor = 64.0 * 2e-13 / (expq ? 1.0 : k.epi.alpha);  // 64 gamma nrm (times c for RatQuad) = dup_factor * w
    const int nsets = h->symmetric ? 1 : 2;
    for (int sidx = 0; sidx < nsets && e == cudaSuccess; ++sidx) {
        const Points& P = sidx == 0 ? h->p0 : h->p1;
        const int64_t rows_alloc = P.npad + kGmRowSlack;
        if (e == cudaSuccess) e = pn_dev_alloc(&h->gm_A[sidx], (size_t)rows_alloc * dpad * 8, stream);
        if (e == cudaSuccess) e = pn_dev_alloc(&h->gm_B[sidx], (size_t)rows_alloc * dpad * 8, stream);
        if (e == cudaSuccess) e = pn_dev_alloc(&h->gm_nrm[sidx], (size_t)rows_alloc * 8, stream);
        if (e == cudaSuccess) e = pn_dev_alloc(&h->gm_thr[sidx], (size_t)rows_alloc * 4, stream);
        if (e == cudaSuccess)
            e = pn_launch_gram_mma_operands((const double*)P.aos, P.n, rows_alloc, d, dpad, centre, scale_dev, iso, gam_over_tol,
                                            dup_factor, expq ? 1 : 0, (double*)h->gm_A[sidx], (double*)h->gm_B[sidx], (double*)h->gm_nrm[sidx],
                                            (int*)h->gm_thr[sidx], stream);
    }
    pn_dev_free(centre, stream);
    if (e != cudaSuccess) return fail_cuda(e, "gram_mma_prepare");
    if (h->symmetric) {
        h->gm_A[1] = h->gm_A[0];
        h->gm_B[1] = h->gm_B[0];
        h->gm_nrm[1] = h->gm_nrm[0];
        h->gm_thr[1] = h->gm_thr[0];
    }
    h->gm_epi = k.epi;
    if (expq) {
        h->gm_epi.c_u = -kLog2e;
        h->gm_epi.c_s = 1.0;
        h->gm_epi.c_su = -kLog2e;
        fill_window(h->gm_epi);
        h->gm_v_hi = h->gm_epi.s_hi_word;  // v = S
    } else {
        // v = 1 + c S: pn_rcp_core / pn_log2_core need v < 2^990 (and alpha log2 v < 1021 for the general form)
        double b_log2_max = 990.0;
        if (k.epi.epi == PN_EPI_RATQUADG) b_log2_max = fmin(b_log2_max, 1021.0 / k.epi.alpha);
        h->gm_v_hi = hi_word(exp2(b_log2_max));
    }
    h->gm_dpad = dpad;
    return 0;
}

// Decide whether the norm-identity kernel may serve this handle and build its operands.
//
// With xt = (x - c) s the identity S = |xt_i|^2 + |yt_j|^2 - 2 xt_i.yt_j carries an absolute
// error of at most about (4 d + 24) 2^-53 R^2, R^2 = max |xt|^2 (rounding of the centred and
// scaled coordinates, of the two norms, of the d products and of the sums), where the
// direct differences of the reference are exact to a few ulp of S itself.  For ExpQuad the
// relative error of K equals the absolute error of S (in the scaled units); for RatQuad it
// is bounded by alpha c |dS| = |dS| / (2 l^2).  The identity is used only if that bound stays
// under 2e-13, a fifth of the 1e-12 parity contract; otherwise (clouds that are wide in units
// of the lengthscale) the direct-difference kernel runs.  Matern needs sqrt(S), whose error
// is unbounded as S -> 0, and always uses direct differences.
int mma_prepare(pncov_handle* h, cudaStream_t stream) {
    if (h->mma_state != 0) return 0;
    h->mma_state = -1;
    const KernelDesc& k = h->k;
    if (k.dtype != PNCOV_F64 || k.d > 64 || (k.kind != PNCOV_EXPQUAD && k.kind != PNCOV_RATQUAD)) return 0;
    if (h->p0.n == 0 || h->p1.n == 0) return 0;
    if (h->opt_identity < 0 && h->env_ident == 0) return 0;  // PNCOV_MATVEC_IDENT=off: always direct differences
    const int d = k.d;
    double* sums = nullptr;
    const int dpad = d <= 8 ? 8 : (d <= 16 ? 16 : (d <= 32 ? 32 : 64));
    h->mma_dpad = dpad;
    PN_CUDA(pn_dev_alloc(reinterpret_cast<void**>(&sums), 128 * sizeof(double) + sizeof(unsigned long long), stream));
    unsigned long long* max_norm = reinterpret_cast<unsigned long long*>(sums + 128);
    auto fail_free = [&](int rc) {
        pn_dev_free(sums, stream);
        return rc;
    };
    coord_sums_kernel<<<1, 1024, 0, stream>>>((const double*)h->p0.aos, h->p0.n, d, sums);
    coord_sums_kernel<<<1, 1024, 0, stream>>>((const double*)h->p1.aos, h->p1.n, d, sums + 64);
    g_pn_launches += 2;
    double host_sums[128];
    cudaError_t e = cudaMemcpyAsync(host_sums, sums, sizeof(host_sums), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return fail_free(fail_cuda(e, "mma_prepare"));
    double centre[64] = {0};
    const double ntot = (double)h->p0.n + (h->symmetric ? 0.0 : (double)h->p1.n);
    for (int m = 0; m < d; ++m)
        centre[m] = (host_sums[m] + (h->symmetric ? 0.0 : host_sums[64 + m])) / ntot;
    e = cudaMemcpyAsync(sums, centre, sizeof(centre), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(max_norm, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return fail_free(fail_cuda(e, "mma_prepare"));
    // ExpQuad: operands carry the scale factors, the exponent constant becomes -log2(e);
    // RatQuad: unscaled (its c multiplies S in the epilogue)
    const bool expq = k.kind == PNCOV_EXPQUAD;
    double iso = 1.0;
    if (expq && !k.ard) iso = sqrt(-k.epi.c_u / kLog2e);  // s, from c_u = -log2(e) s^2
    const double* scale_dev = (expq && k.ard) ? (const double*)h->scale_dev : nullptr;
    const int nsets = h->symmetric ? 1 : 2;
    for (int sidx = 0; sidx < nsets; ++sidx) {
        const Points& P = sidx == 0 ? h->p0 : h->p1;
        const size_t bytes = (size_t)P.npad * (dpad + 4) * sizeof(double);
        e = pn_dev_alloc(&h->a12[sidx], bytes, stream);
        if (e == cudaSuccess) e = pn_dev_alloc(&h->b12[sidx], bytes, stream);
        if (e != cudaSuccess) return fail_free(fail_cuda(e, "mma_prepare"));
        build_mma_operands_kernel<<<(unsigned)((P.npad + 255) / 256), 256, 0, stream>>>(
            (const double*)P.aos, P.n, P.npad, d, dpad, sums, scale_dev, iso, (double*)h->a12[sidx],
            (double*)h->b12[sidx], max_norm, expq ? 1 : 0);
        ++g_pn_launches;
    }
    if (h->symmetric) {
        h->a12[1] = h->a12[0];
        h->b12[1] = h->b12[0];
    }
    unsigned long long bits = 0;
    e = cudaMemcpyAsync(&bits, max_norm, sizeof(bits), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    pn_dev_free(sums, stream);
    if (e != cudaSuccess) return fail_cuda(e, "mma_prepare");
    double r2;
    memcpy(&r2, &bits, 8);
    const double sens = expq ? 1.0 : k.epi.alpha * k.epi.c_s;  // d(log K)/dS bound
    const double bound = (4.0 * d + 24.0) * ldexp(1.0, -53) * r2 * sens;
    if (!(bound <= 2e-13)) return 0;  // stays at -1: direct differences
    h->mma_epi = k.epi;
    if (expq) {
        h->mma_epi.c_u = -kLog2e;
        h->mma_epi.c_s = 1.0;
        h->mma_epi.c_su = -kLog2e;
        fill_window(h->mma_epi);
    }
    h->mma_state = 1;
    return 0;
}

// sym_nparts > 0 asks for the symmetric-halving kernel restricted to row blocks sym_part, sym_part +
// sym_nparts, ... (the caller sums the partial vectors of all parts); it fails with PNCOV_ENOTIMPL
// when that kernel does not apply to this handle.  sym_nparts == 0: whole product, best kernel.
// integer planes of a symmetric-halving product -> columns c0 .. c0 + kc - 1 of y (row factors of the factorised
// ExpQuad form applied)
int sym_finish(pncov_handle* h, const void* bins, void* y, int64_t ldy, cudaStream_t stream, int kc = 1, int64_t k = 1,
               int64_t c0 = 0) {
    const int64_t n = h->p0.n;
    int rc;
    if ((rc = ensure(&h->ypack, &h->ypack_bytes, (size_t)n * kc * 8, h->stream))) return rc;
    const double* fac = h->last_fact ? (const double*)h->a12[0] : nullptr;
    PN_CUDA(pn_launch_unbin((const long long*)bins, n, kc, n, (const double*)h->vmax_dev, fac, h->mma_dpad + 4, h->mma_dpad + 1,
                            (double*)h->ypack, stream));
    const unsigned blocks = (unsigned)((n * kc + 255) / 256);
    if (h->last_fact)
        unpack_y_scaled_kernel<<<blocks, 256, 0, stream>>>((const double*)h->ypack, n, kc, k, c0, fac, h->mma_dpad + 4,
                                                           h->mma_dpad + 1, (double*)y, ldy);
    else
        unpack_y_kernel<double><<<blocks, 256, 0, stream>>>((const double*)h->ypack, n, kc, k, c0, (double*)y, ldy);
    ++g_pn_launches;
    PN_CUDA(cudaGetLastError());
    return 0;
}

// `bins_out`: with sym_nparts > 0, leave the integer planes of the partial product there (3 n0 words of 8 bytes,
// overwritten) instead of finishing it into y
int matmul_device_impl(pncov_handle* h, const void* v, int64_t k, int64_t ldv, void* y, int64_t ldy,
                       bool transpose, cudaStream_t stream, int sym_part = 0, int sym_nparts = 0, void* bins_out = nullptr) {
    const Points& R = transpose ? h->p1 : h->p0;
    const Points& C = transpose ? h->p0 : h->p1;
    const size_t es = h->k.elem();
    if (k <= 0 || R.n <= 0) return 0;
    if (C.n <= 0) {  // empty sum: Y = 0
        PN_CUDA(cudaMemset2DAsync(y, (size_t)ldy * es, 0, (size_t)k * es, (size_t)R.n, stream));
        return 0;
    }
    if (h->has_prog) return fail(PNCOV_ENOTIMPL, "a handle with a program (pncov_set_program) serves the dense Gram only");
    if (h->k.extra_mode >= 0)
        return fail(PNCOV_ENOTIMPL, "the matrix-free product covers ExpQuad, Matern and RatQuad; use the dense Gram for this kernel");
    if (h->k.d > 1024) return fail(PNCOV_ENOTIMPL, "matrix-free product supports d <= 1024");
    int kc_max = (k == 1) ? 1 : 4;
    PnMatvecLaunch m{};
    m.x0_aos = R.aos;
    m.x1_aos = C.aos;
    m.x1_soa = C.soa;
    m.ld1 = C.npad;
    m.x0_soa = R.soa;
    m.ld0 = R.npad;
    m.n0 = R.n;
    m.n1 = C.n;
    m.n1pad = C.npad;
    m.scale = h->scale_dev;
    m.d = h->k.d;
    m.kc = kc_max;
    m.in_f32 = h->k.dtype == PNCOV_F32;
    m.epi = h->k.epi;
    if (h->opt_identity != 0) {
        int prc = mma_prepare(h, stream);
        if (prc) return prc;
    }
    const bool wide_ok = h->k.d <= 8 || kc_max == 1;  // the 8 < d <= 64 kernels take one right-hand side
    // five or more right-hand sides on the DMMA identity kernel (d <= 8): passes of eight columns
    if (h->mma_state == 1 && h->opt_identity != 0 && h->env_ident != 1 && h->k.d <= 8 && k > 4) {
        kc_max = 8;
        m.kc = 8;
    }
    if (h->mma_state == 1 && h->opt_identity != 0 && wide_ok) {
        // norm identity: dot products on DMMA by default (fastest measured), DFMA form opt-in
        m.use_mma = h->env_ident == 1 ? 1 : 2;
        m.a12 = h->a12[transpose ? 1 : 0];
        m.b12 = h->b12[transpose ? 0 : 1];
        m.epi = h->mma_epi;
        const bool fact_on = h->opt_factorised >= 0 ? h->opt_factorised != 0 : h->env_fact != 0;
        m.fact = (m.use_mma == 2 && h->k.kind == PNCOV_EXPQUAD && fact_on) ? 1 : 0;
        if (h->k.d > 8 && (m.use_mma != 2 || (h->k.kind == PNCOV_EXPQUAD && !m.fact))) {
            m.use_mma = 0;  // wide operands exist only as DMMA kernels, ExpQuad only in factorised form
            m.a12 = m.b12 = nullptr;
            m.epi = h->k.epi;
            m.fact = 0;
        }
    }
    {
        // symmetric halving: x1 is x0, one right-hand side, DMMA identity kernel admissible or direct differences.
        // Reproducible (integer planes); PNCOV_MATVEC_SYM=off selects the kernel that evaluates all N^2 pairs.
        const bool kernel_ok = m.use_mma == 2 || (m.use_mma == 0 && h->k.d <= 96);  // DMMA identity or direct differences
        const bool sym_on = h->opt_symmetric >= 0 ? h->opt_symmetric != 0 : h->env_sym != 0;
        bool sym_ok = h->symmetric && kernel_ok && kc_max == 1 && es == 8 && R.n >= 2048 && sym_on;
        if (sym_nparts > 0 && !sym_ok && k == 1)
            return fail(PNCOV_ENOTIMPL, "symmetric partial product needs x1 = x0, one column, float64, d <= 96 and n >= 2048");
        // several right-hand sides on the DMMA identity kernel (d <= 8): symmetric passes of eight columns (zero padded)
        if (h->symmetric && m.use_mma == 2 && h->k.d <= 8 && k > 1 && es == 8 && R.n >= 2048 && sym_on &&
            (sym_nparts == 0 || (bins_out && k <= 8))) {
            sym_ok = true;
            kc_max = 8;
            m.kc = kc_max;
        }
        if (sym_nparts > 0 && !sym_ok)
            return fail(PNCOV_ENOTIMPL, "symmetric partial products with 2 .. 8 columns need x1 = x0, float64, d <= 8, n >= 2048 and "
                                        "an admissible norm-identity kernel");
        m.sym = sym_ok ? 1 : 0;
        m.sym_part = sym_part;
        m.sym_nparts = sym_nparts > 0 ? sym_nparts : 1;
        h->last_sym = m.sym != 0;
        h->last_use_mma = m.use_mma;
        h->last_fact = m.sym != 0 && m.fact != 0;
    }
    PN_CUDA(pn_plan_matvec(m, h->sm_count));
    const int jsplit = m.jsplit;
    int rc;
    void* bins = nullptr;
    if (m.sym) {
        // order-independent accumulation (pn_bin_add): max |v| and zeroed integer planes
        if (!h->vmax_dev) PN_CUDA(pn_dev_alloc(&h->vmax_dev, sizeof(double), h->stream));
        bins = bins_out;
        if (!bins) {
            if ((rc = ensure(&h->bins, &h->bins_bytes, (size_t)R.n * kc_max * 24, h->stream))) return rc;
            bins = h->bins;
        }
        PN_CUDA(pn_launch_vmax((const double*)v, ldv, C.n, (int)k, (double*)h->vmax_dev, stream));
        m.bins = (long long*)bins;
        m.vmax = (const double*)h->vmax_dev;
    }
    if ((rc = ensure(&h->vpack, &h->vpack_bytes, (size_t)C.npad * (kc_max + 2) * es, h->stream))) return rc;
    if ((rc = ensure(&h->ypack, &h->ypack_bytes, (size_t)R.n * kc_max * es, h->stream))) return rc;
    if (jsplit > 1 && (rc = ensure(&h->partial, &h->partial_bytes, (size_t)jsplit * R.n * kc_max * 8, h->stream))) return rc;

    for (int64_t c0 = 0; c0 < k; c0 += kc_max) {
        const int kc = kc_max;
        {
            const int vstride = kc == 8 ? 10 : kc;  // the eight-column DMMA kernels read V in padded rows (conflict-free fragments)
            m.vstride = vstride;
            const int64_t total = C.npad * vstride;
            const unsigned blocks = (unsigned)((total + 255) / 256);
            if (m.fact)
                pack_v_scaled_kernel<<<blocks, 256, 0, stream>>>((const double*)v, ldv, C.n, C.npad, k, c0, kc, vstride,
                                                                 (const double*)m.b12, h->mma_dpad + 4, h->mma_dpad + 1,
                                                                 (double*)h->vpack);
            else if (es == 8)
                pack_v_kernel<double><<<blocks, 256, 0, stream>>>((const double*)v, ldv, C.n, C.npad, k, c0, kc, vstride, (double*)h->vpack);
            else
                pack_v_kernel<float><<<blocks, 256, 0, stream>>>((const float*)v, ldv, C.n, C.npad, k, c0, kc, vstride, (float*)h->vpack);
            ++g_pn_launches;
            PN_CUDA(cudaGetLastError());
        }
        m.v = h->vpack;
        m.y = h->ypack;
        m.kc = kc;
        m.partial = h->partial;
        m.jsplit = jsplit;
        if (m.sym) PN_CUDA(cudaMemsetAsync(bins, 0, (size_t)R.n * kc * 24, stream));
        PN_CUDA(pn_launch_matvec(m, h->sm_count, stream));
        if (m.sym) {
            if (!bins_out && (rc = sym_finish(h, bins, y, ldy, stream, kc, k, c0))) return rc;
            continue;
        }
        {
            const int64_t total = R.n * kc;
            const unsigned blocks = (unsigned)((total + 255) / 256);
            if (m.fact)
                unpack_y_scaled_kernel<<<blocks, 256, 0, stream>>>((const double*)h->ypack, R.n, kc, k, c0,
                                                                   (const double*)m.a12, h->mma_dpad + 4, h->mma_dpad + 1,
                                                                   (double*)y, ldy);
            else if (es == 8)
                unpack_y_kernel<double><<<blocks, 256, 0, stream>>>((const double*)h->ypack, R.n, kc, k, c0, (double*)y, ldy);
            else
                unpack_y_kernel<float><<<blocks, 256, 0, stream>>>((const float*)h->ypack, R.n, kc, k, c0, (float*)y, ldy);
            ++g_pn_launches;
            PN_CUDA(cudaGetLastError());
        }
    }
    return 0;
}

}  // namespace

extern "C" {

PNCOV_API int pncov_version(void) { return PNCOV_VERSION; }

PNCOV_API const char* pncov_last_error(void) { return t_err.c_str(); }

PNCOV_API int pncov_device_count(int* count) {
    if (!count) return fail(PNCOV_EINVAL, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail_cuda(e, "cudaGetDeviceCount");
    }
    return 0;
}

// Large page-locked blocks: an anonymous mapping advised towards transparent huge pages, populated by all host
// worker threads at once and then registered with the driver.  cudaHostAlloc populates and pins 4 KB pages from one
// thread (0.2 s per GB on this pool's hosts: 3.5 s before the first 2 x 8.6 GB result); huge pages also take the
// TLB pressure off the host-side mirroring of symmetric results (pncov_host.cu).  PNCOV_HOST_ALLOC=cuda keeps
// cudaHostAlloc for everything.
namespace {
std::mutex g_host_map_mutex;
std::map<void*, size_t> g_host_maps;  // blocks that came from mmap + cudaHostRegister
}  // namespace

PNCOV_API int pncov_host_alloc(void** ptr, int64_t nbytes) {
    if (!ptr || nbytes < 0) return fail(PNCOV_EINVAL, "bad arguments");
    *ptr = nullptr;
    if (nbytes == 0) return 0;
    static const bool use_mmap = !(getenv("PNCOV_HOST_ALLOC") && strcmp(getenv("PNCOV_HOST_ALLOC"), "cuda") == 0);
    static const bool use_thp = !(getenv("PNCOV_HOST_ALLOC") && strcmp(getenv("PNCOV_HOST_ALLOC"), "mmap4k") == 0);
    if (use_mmap && nbytes >= (64ll << 20)) {
        const size_t len = ((size_t)nbytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
        void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            madvise(p, len, use_thp ? MADV_HUGEPAGE : MADV_NOHUGEPAGE);
            pn_host_touch(p, len);
            if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_host_map_mutex);
                g_host_maps[p] = len;
                *ptr = p;
                return 0;
            }
            cudaGetLastError();
            munmap(p, len);
        }
    }
    PN_CUDA(cudaHostAlloc(ptr, (size_t)nbytes, cudaHostAllocPortable));
    return 0;
}

PNCOV_API int pncov_host_free(void* ptr) {
    if (!ptr) return 0;
    size_t len = 0;
    {
        std::lock_guard<std::mutex> lk(g_host_map_mutex);
        auto it = g_host_maps.find(ptr);
        if (it != g_host_maps.end()) {
            len = it->second;
            g_host_maps.erase(it);
        }
    }
    if (len) {
        cudaError_t e = cudaHostUnregister(ptr);
        munmap(ptr, len);
        if (e != cudaSuccess) return fail_cuda(e, "cudaHostUnregister");
        return 0;
    }
    PN_CUDA(cudaFreeHost(ptr));
    return 0;
}

PNCOV_API int pncov_create(pncov_handle** out, int device, int kind, int dtype, int64_t d, const double* lengthscales,
                 int64_t n_lengthscales, double nu, double alpha, const void* x0, int64_t n0, const void* x1,
                 int64_t n1, int points_on_device) {
    if (!out) return fail(PNCOV_EINVAL, "out is NULL");
    *out = nullptr;
    if (n0 < 0 || (x1 && n1 < 0)) return fail(PNCOV_EINVAL, "negative point count");
    if ((n0 > 0 && !x0)) return fail(PNCOV_EINVAL, "x0 is NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PNCOV_ENODEVICE, "no CUDA device available (libpncov has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(PNCOV_EINVAL, "device index out of range");
    pncov_handle* h = new (std::nothrow) pncov_handle();
    if (!h) return fail(PNCOV_ENOMEM, "out of host memory");
    int rc = build_desc(h->k, kind, dtype, d, lengthscales, n_lengthscales, nu, alpha);
    if (rc) {
        delete h;
        return rc;
    }
    h->device = device;
    DeviceGuard guard(device);
    auto bail = [&](int code) {
        pncov_destroy(h);
        return code;
    };
    if (!guard.ok) return bail(fail(PNCOV_ENODEVICE, "cudaSetDevice failed"));
    int cc_major = 0, sms = 0;
    cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device);
    (void)sms;
    h->sm_count = pn_sm_count();
    if (cc_major < 10) return bail(fail(PNCOV_ENODEVICE, "libpncov is built for sm_100a (B200) only"));
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(fail_cuda(e, "cudaStreamCreate"));
    if ((e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(fail_cuda(e, "cudaStreamCreate"));
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return bail(fail_cuda(e, "cudaEventCreate"));
    if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return bail(fail_cuda(e, "cudaEventCreate"));
    if ((rc = upload_scale(h->k, &h->scale_dev, h->stream))) return bail(rc);
    if ((rc = upload_pm(h->k, &h->pm_dev, h->stream))) return bail(rc);
    if ((rc = upload_points(h, h->p0, x0, n0, points_on_device != 0))) return bail(rc);
    if (x1) {
        if ((rc = upload_points(h, h->p1, x1, n1, points_on_device != 0))) return bail(rc);
    } else {
        h->p1 = h->p0;
        h->symmetric = true;
    }
    if (h->k.extra_mode == 0 && h->k.d > 16) {
        // inner products as a GEMM: K-contiguous operands padded to a multiple of 16 coordinates
        h->dot_dpad = ((h->k.d + 15) / 16) * 16;
        const int nsets = h->symmetric ? 1 : 2;
        for (int sidx = 0; sidx < nsets; ++sidx) {
            const Points& P = sidx == 0 ? h->p0 : h->p1;
            if ((e = pn_dev_alloc(&h->dot_ops[sidx], (size_t)P.npad * h->dot_dpad * 8, h->stream)) != cudaSuccess)
                return bail(fail_cuda(e, "operand allocation"));
            if ((e = pn_launch_pad_rows((const double*)P.aos, P.n, h->k.d, P.npad, h->dot_dpad, (double*)h->dot_ops[sidx],
                                        h->stream)) != cudaSuccess)
                return bail(fail_cuda(e, "pad_rows"));
        }
        if (h->symmetric) h->dot_ops[1] = h->dot_ops[0];
    }
    if ((rc = gram_mma_prepare(h))) return bail(rc);
    if (const char* v = getenv("PNCOV_MATVEC_IDENT")) h->env_ident = strcmp(v, "off") == 0 ? 0 : (strcmp(v, "dfma") == 0 ? 1 : -1);
    if (const char* v = getenv("PNCOV_MATVEC_FACT")) h->env_fact = strcmp(v, "off") == 0 ? 0 : -1;
    if (const char* v = getenv("PNCOV_MATVEC_SYM")) h->env_sym = strcmp(v, "off") == 0 ? 0 : (strcmp(v, "on") == 0 ? 1 : -1);
    if (const char* v = getenv("PNCOV_GRAM_MIRROR")) h->env_mirror = strcmp(v, "off") == 0 ? 0 : (strcmp(v, "on") == 0 ? 1 : -1);
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return bail(fail_cuda(e, "upload"));
    *out = h;
    return 0;
}

PNCOV_API int pncov_destroy(pncov_handle* h) {
    if (!h) return 0;
    DeviceGuard guard(h->device);
    // device-pointer entry points may have queued work on caller-owned streams: drain the
    // device (what cudaFree used to do implicitly) before the buffers go back to the pool
    if (h->used_foreign_stream) cudaDeviceSynchronize();
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    cudaStream_t fs = h->stream;
    if (!h->symmetric) free_points(h->p1, fs);
    free_points(h->p0, fs);
    pn_dev_free(h->scale_dev, fs);
    pn_dev_free(h->pm_dev, fs);
    pn_dev_free(h->dot_ops[0], fs);
    if (!h->symmetric) pn_dev_free(h->dot_ops[1], fs);
    for (int sidx = 0; sidx < (h->symmetric ? 1 : 2); ++sidx) {
        pn_dev_free(h->gm_A[sidx], fs);
        pn_dev_free(h->gm_B[sidx], fs);
        pn_dev_free(h->gm_nrm[sidx], fs);
        pn_dev_free(h->gm_thr[sidx], fs);
    }
    pn_dev_free(h->vpack, fs);
    pn_dev_free(h->ypack, fs);
    pn_dev_free(h->partial, fs);
    pn_dev_free(h->bins, fs);
    pn_dev_free(h->vmax_dev, fs);
    pn_dev_free(h->stage, fs);
    pn_dev_free(h->a12[0], fs);
    pn_dev_free(h->b12[0], fs);
    if (!h->symmetric) {
        pn_dev_free(h->a12[1], fs);
        pn_dev_free(h->b12[1], fs);
    }
    if (fs) cudaStreamSynchronize(fs);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
    return 0;
}

PNCOV_API int pncov_shape(const pncov_handle* h, int64_t* n0, int64_t* n1) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (n0) *n0 = h->p0.n;
    if (n1) *n1 = h->p1.n;
    return 0;
}

PNCOV_API int pncov_gram_dtype(const pncov_handle* h, int* dtype) {
    if (!h || !dtype) return fail(PNCOV_EINVAL, "NULL argument");
    *dtype = h->k.gram_f32() ? PNCOV_F32 : PNCOV_F64;
    return 0;
}

PNCOV_API int pncov_gram_device(pncov_handle* h, void* out, int64_t ld, int transpose, void* stream_) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    const int64_t rows = transpose ? h->p1.n : h->p0.n;
    const int64_t cols = transpose ? h->p0.n : h->p1.n;
    if (rows == 0 || cols == 0) return 0;
    if (!out || ld < cols) return fail(PNCOV_EINVAL, "bad output buffer / leading dimension");
    DeviceGuard guard(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    h->used_foreign_stream = true;
    PN_CUDA(cudaEventRecord(h->ev0, stream));
    PN_CUDA(launch_gram_any(h, transpose != 0, 0, rows, out, ld, stream));
    PN_CUDA(cudaEventRecord(h->ev1, stream));
    h->timed = true;
    return 0;
}

PNCOV_API int pncov_gram(pncov_handle* h, void* out, int64_t ld, int transpose) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    const int64_t rows = transpose ? h->p1.n : h->p0.n;
    const int64_t cols = transpose ? h->p0.n : h->p1.n;
    if (rows == 0 || cols == 0) return 0;
    if (!out || ld < cols) return fail(PNCOV_EINVAL, "bad output buffer / leading dimension");
    DeviceGuard guard(h->device);
    const size_t ge = h->k.gram_elem();
    // row blocks of <= 1 GiB, double buffered: block b+1 is computed while block b is copied out
    // the staging rows are padded to an even length so that every row starts 16-byte aligned
    // (the fast kernel's vector stores); the copy below strips the padding again
    const int64_t ld_stage = cols + (cols & 1);
    // A whole symmetric matrix (x1 = NULL; every kernel gives K_ij and K_ji the same bits): only the columns from
    // the diagonal block on cross PCIe, the blocks below the diagonal are written by host threads as mirror
    // images while the next row blocks are in flight (pncov_host.cu).  Blocks of 128 MiB let the mirroring start
    // early.  Measured (tools/e2e_mirror_sweep.py, N = 32768, 16 host threads): 157 -> 119 ms per matrix; the
    // limit is then the host's memory bandwidth (8 threads mirror as fast as 16), so copying the last rows whole
    // to balance link and host (PNCOV_GRAM_MIRROR_FULL_DIV) does not pay.  PNCOV_GRAM_MIRROR=off copies everything.
    const bool mirror_ok = h->symmetric && rows == cols && (size_t)rows * ld_stage * ge > (1ull << 30) && h->env_mirror != 0;
    // PNCOV_GRAM_MIRROR=on forces the mirrored path, unset lets the measured policy decide
    const bool mirror = mirror_ok && (h->env_mirror == 1 || g_mirror_policy.choose());
    const auto t_begin = std::chrono::steady_clock::now();
    static const int mirror_mib = getenv("PNCOV_GRAM_MIRROR_BLOCK_MIB") ? atoi(getenv("PNCOV_GRAM_MIRROR_BLOCK_MIB")) : 128;
    static const int mirror_div = getenv("PNCOV_GRAM_MIRROR_FULL_DIV") ? atoi(getenv("PNCOV_GRAM_MIRROR_FULL_DIV")) : 0;
    int64_t block_rows = (int64_t)((mirror ? ((size_t)std::max(mirror_mib, 16) << 20) : (1ull << 30)) / ((size_t)ld_stage * ge));
    block_rows = std::max<int64_t>(PN_PAD, (block_rows / PN_PAD) * PN_PAD);
    block_rows = std::min<int64_t>(block_rows, pad_up(rows));
    const size_t block_bytes = (size_t)block_rows * ld_stage * ge;
    const int nbuf = rows > block_rows ? 2 : 1;
    const int64_t full_from = (mirror && mirror_div > 0) ? rows - ((rows / mirror_div) / 256) * 256 : rows;  // rows from here on are copied whole
    int rc;
    if ((rc = ensure(&h->stage, &h->stage_bytes, block_bytes * nbuf, h->stream))) return rc;
    cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    for (int i = 0; i < nbuf; ++i) {
        PN_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        PN_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
    }
    std::vector<cudaEvent_t> landed;  // mirror: one event per row block, recorded after its copy
    cudaError_t err = cudaSuccess;
    err = cudaEventRecord(h->ev0, h->stream);
    int b = 0;
    for (int64_t r0 = 0; r0 < rows && err == cudaSuccess; r0 += block_rows, ++b) {
        const int64_t nr = std::min(block_rows, rows - r0);
        const int buf = b % nbuf;
        char* dbuf = static_cast<char*>(h->stage) + (size_t)buf * block_bytes;
        if (b >= nbuf) err = cudaStreamWaitEvent(h->stream, copied[buf], 0);
        if (err == cudaSuccess) err = launch_gram_any(h, transpose != 0, r0, nr, dbuf, ld_stage, h->stream);
        if (err == cudaSuccess) err = cudaEventRecord(done[buf], h->stream);
        if (err == cudaSuccess) err = cudaStreamWaitEvent(h->copy_stream, done[buf], 0);
        // rows below full_from are copied whole, the others from their diagonal block on
        const int64_t split = std::min(std::max(full_from, r0), r0 + nr);  // rows [r0, split): upper part only
        g_pn_d2h_bytes += (int64_t)ge * ((split - r0) * (cols - (mirror ? r0 : 0)) + (r0 + nr - split) * cols);
        if (err == cudaSuccess && split > r0) {
            const int64_t c_first = mirror ? r0 : 0;
            err = cudaMemcpy2DAsync(static_cast<char*>(out) + ((size_t)r0 * ld + c_first) * ge, (size_t)ld * ge,
                                    dbuf + (size_t)c_first * ge, (size_t)ld_stage * ge, (size_t)(cols - c_first) * ge,
                                    (size_t)(split - r0), cudaMemcpyDeviceToHost, h->copy_stream);
        }
        if (err == cudaSuccess && split < r0 + nr)
            err = cudaMemcpy2DAsync(static_cast<char*>(out) + (size_t)split * ld * ge, (size_t)ld * ge,
                                    dbuf + (size_t)(split - r0) * ld_stage * ge, (size_t)ld_stage * ge, (size_t)cols * ge,
                                    (size_t)(r0 + nr - split), cudaMemcpyDeviceToHost, h->copy_stream);
        if (err == cudaSuccess) err = cudaEventRecord(copied[buf], h->copy_stream);
        if (mirror && err == cudaSuccess) {
            cudaEvent_t ev = nullptr;
            err = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (err == cudaSuccess) {
                landed.push_back(ev);
                err = cudaEventRecord(ev, h->copy_stream);
            }
        }
    }
    if (err == cudaSuccess) err = cudaEventRecord(h->ev1, h->stream);
    if (mirror) {
        // as each row block lands, its part right of the diagonal block is mirrored below the diagonal
        int64_t r0 = 0;
        for (size_t i = 0; i < landed.size(); ++i, r0 += block_rows) {
            if (err == cudaSuccess) err = cudaEventSynchronize(landed[i]);
            if (err == cudaSuccess) pn_host_mirror_async(out, ld, full_from, r0, std::min(r0 + block_rows, rows), (int)ge);
        }
        pn_host_mirror_wait();
        for (cudaEvent_t ev : landed) cudaEventDestroy(ev);
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(h->copy_stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
    if (mirror_ok && h->env_mirror != 1 && err == cudaSuccess)
        g_mirror_policy.record(mirror, (double)rows * cols * ge,
                               std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count());
    for (int i = 0; i < nbuf; ++i) {
        cudaEventDestroy(done[i]);
        cudaEventDestroy(copied[i]);
    }
    h->timed = true;
    if (err != cudaSuccess) return fail_cuda(err, "pncov_gram");
    return 0;
}

PNCOV_API int pncov_matmul_device(pncov_handle* h, const void* v, int64_t k, int64_t ldv, void* y, int64_t ldy, int transpose,
                        void* stream_) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (k < 0 || ldv < k || ldy < k) return fail(PNCOV_EINVAL, "bad column count / leading dimension");
    if (k > 0 && (!v || !y)) return fail(PNCOV_EINVAL, "NULL operand");
    DeviceGuard guard(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    h->used_foreign_stream = true;
    PN_CUDA(cudaEventRecord(h->ev0, stream));
    int rc = matmul_device_impl(h, v, k, ldv, y, ldy, transpose != 0, stream);
    if (rc) return rc;
    PN_CUDA(cudaEventRecord(h->ev1, stream));
    h->timed = true;
    return 0;
}

PNCOV_API int pncov_matmul_sym_part_device(pncov_handle* h, const void* v, void* y, int part, int nparts, void* stream_) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (nparts < 1 || part < 0 || part >= nparts) return fail(PNCOV_EINVAL, "bad part / nparts");
    if (!v || !y) return fail(PNCOV_EINVAL, "NULL operand");
    DeviceGuard guard(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    h->used_foreign_stream = true;
    PN_CUDA(cudaEventRecord(h->ev0, stream));
    int rc = matmul_device_impl(h, v, 1, 1, y, 1, false, stream, part, nparts);
    if (rc) return rc;
    PN_CUDA(cudaEventRecord(h->ev1, stream));
    h->timed = true;
    return 0;
}

PNCOV_API int pncov_matmul_sym_part_bins_device(pncov_handle* h, const void* v, void* bins, int part, int nparts, void* stream_) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (nparts < 1 || part < 0 || part >= nparts) return fail(PNCOV_EINVAL, "bad part / nparts");
    if (!v || !bins) return fail(PNCOV_EINVAL, "NULL operand");
    DeviceGuard guard(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    h->used_foreign_stream = true;
    PN_CUDA(cudaEventRecord(h->ev0, stream));
    int rc = matmul_device_impl(h, v, 1, 1, nullptr, 1, false, stream, part, nparts, bins);
    if (rc) return rc;
    PN_CUDA(cudaEventRecord(h->ev1, stream));
    h->timed = true;
    return 0;
}

PNCOV_API int pncov_matmulk_sym_part_bins_device(pncov_handle* h, const void* v, int64_t k, int64_t ldv, void* bins, int part,
                                                 int nparts, void* stream_) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (nparts < 1 || part < 0 || part >= nparts) return fail(PNCOV_EINVAL, "bad part / nparts");
    if (k < 1 || k > 8 || ldv < k) return fail(PNCOV_EINVAL, "1 .. 8 columns per call");
    if (!v || !bins) return fail(PNCOV_EINVAL, "NULL operand");
    DeviceGuard guard(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    h->used_foreign_stream = true;
    PN_CUDA(cudaEventRecord(h->ev0, stream));
    int rc = matmul_device_impl(h, v, k, ldv, nullptr, 1, false, stream, part, nparts, bins);
    if (rc) return rc;
    PN_CUDA(cudaEventRecord(h->ev1, stream));
    h->timed = true;
    return 0;
}

PNCOV_API int pncov_matmulk_sym_finish_device(pncov_handle* h, const void* bins, int64_t k, void* y, int64_t ldy, void* stream_) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (!bins || !y || k < 1 || k > 8 || ldy < k) return fail(PNCOV_EINVAL, "bad arguments");
    if (!h->last_sym || !h->vmax_dev) return fail(PNCOV_EINVAL, "no symmetric partial product was computed on this handle");
    DeviceGuard guard(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    h->used_foreign_stream = true;
    return sym_finish(h, bins, y, ldy, stream, k == 1 ? 1 : 8, k, 0);
}

PNCOV_API int pncov_matmul_sym_finish_device(pncov_handle* h, const void* bins, void* y, void* stream_) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (!bins || !y) return fail(PNCOV_EINVAL, "NULL operand");
    if (!h->last_sym || !h->vmax_dev) return fail(PNCOV_EINVAL, "no symmetric partial product was computed on this handle");
    DeviceGuard guard(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    h->used_foreign_stream = true;
    return sym_finish(h, bins, y, 1, stream);
}

PNCOV_API int pncov_matmul(pncov_handle* h, const void* v, int64_t k, int64_t ldv, void* y, int64_t ldy, int transpose) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (k < 0 || ldv < k || ldy < k) return fail(PNCOV_EINVAL, "bad column count / leading dimension");
    const int64_t rows = transpose ? h->p1.n : h->p0.n;
    const int64_t cols = transpose ? h->p0.n : h->p1.n;
    if (k == 0 || rows == 0) return 0;
    if (!v && cols > 0) return fail(PNCOV_EINVAL, "NULL operand");
    if (!y) return fail(PNCOV_EINVAL, "NULL result");
    DeviceGuard guard(h->device);
    const size_t es = h->k.elem();
    const size_t vbytes = (size_t)std::max<int64_t>(cols, 1) * k * es, ybytes = (size_t)rows * k * es;
    int rc;
    if ((rc = ensure(&h->stage, &h->stage_bytes, vbytes + ybytes + 256, h->stream))) return rc;
    char* dv = static_cast<char*>(h->stage);
    char* dy = dv + ((vbytes + 255) / 256) * 256;
    if (cols > 0)
        PN_CUDA(cudaMemcpy2DAsync(dv, (size_t)k * es, v, (size_t)ldv * es, (size_t)k * es, (size_t)cols, cudaMemcpyHostToDevice, h->stream));
    PN_CUDA(cudaEventRecord(h->ev0, h->stream));
    if ((rc = matmul_device_impl(h, dv, k, k, dy, k, transpose != 0, h->stream))) return rc;
    PN_CUDA(cudaEventRecord(h->ev1, h->stream));
    PN_CUDA(cudaMemcpy2DAsync(y, (size_t)ldy * es, dy, (size_t)k * es, (size_t)k * es, (size_t)rows, cudaMemcpyDeviceToHost, h->stream));
    PN_CUDA(cudaStreamSynchronize(h->stream));
    h->timed = true;
    return 0;
}

PNCOV_API int pncov_pairs(int device, int kind, int dtype, int64_t d, const double* lengthscales, int64_t n_lengthscales,
                double nu, double alpha, const void* x0, const void* x1, int64_t b, void* out) {
    KernelDesc k;
    int rc = build_desc(k, kind, dtype, d, lengthscales, n_lengthscales, nu, alpha);
    if (rc) return rc;
    if (b < 0) return fail(PNCOV_EINVAL, "negative batch size");
    if (b == 0) return 0;
    if (!x0 || !x1 || !out) return fail(PNCOV_EINVAL, "NULL argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PNCOV_ENODEVICE, "no CUDA device available (libpncov has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(PNCOV_EINVAL, "device index out of range");
    DeviceGuard guard(device);
    const size_t es = k.elem(), oe = k.gram_elem();
    const size_t pbytes = (size_t)b * d * es;
    char* buf = nullptr;
    PN_CUDA(pn_dev_alloc(reinterpret_cast<void**>(&buf), 2 * pbytes + (size_t)b * oe + 512, 0));
    void* scale_dev = nullptr;
    void* pm_dev = nullptr;
    auto cleanup = [&]() {
        cudaStreamSynchronize(0);
        pn_dev_free(buf, 0);
        pn_dev_free(scale_dev, 0);
        pn_dev_free(pm_dev, 0);
    };
    if ((rc = upload_scale(k, &scale_dev, 0)) || (rc = upload_pm(k, &pm_dev, 0))) {
        cleanup();
        return rc;
    }
    char* d0 = buf;
    char* d1 = buf + ((pbytes + 255) / 256) * 256;
    char* dout = d1 + ((pbytes + 255) / 256) * 256;
    // note: buf was sized with 512 bytes of slack for the two roundings above
    cudaError_t e = cudaMemcpy(d0, x0, pbytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d1, x1, pbytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && k.extra_mode >= 0) {
        PnExtraLaunch g{};
        g.mode = k.extra_mode;
        g.x0_soa = d0;
        g.x1_soa = d1;
        g.n0 = b;
        g.d = (int)d;
        g.pm = pm_dev;
        g.constant = k.constant;
        g.exponent = k.exponent;
        g.out = dout;
        e = pn_launch_pairs_extra(g, 0);
    } else if (e == cudaSuccess) {
        PnPairsLaunch p{};
        p.x0 = d0;
        p.x1 = d1;
        p.b = b;
        p.scale = scale_dev;
        p.d = (int)d;
        p.out = dout;
        p.in_f32 = dtype == PNCOV_F32;
        p.out_f32 = k.gram_f32();
        p.epi = k.epi;
        e = pn_launch_pairs(p, 0);
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, (size_t)b * oe, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return fail_cuda(e, "pncov_pairs");
    return 0;
}

// ---- device buffers and covariance arithmetic on them ---------------------------------------------
PNCOV_API int pncov_device_alloc(int device, void** ptr, int64_t nbytes) {
    if (!ptr || nbytes < 0) return fail(PNCOV_EINVAL, "bad arguments");
    *ptr = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PNCOV_ENODEVICE, "no CUDA device available (libpncov has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(PNCOV_EINVAL, "device index out of range");
    DeviceGuard guard(device);
    // the stream-ordered pool (blocks of repeated calls are reused; cudaMalloc / cudaFree of a multi-GB block
    // cost tens to hundreds of milliseconds each)
    cudaError_t e = pn_dev_alloc(ptr, (size_t)std::max<int64_t>(nbytes, 1), cudaStreamPerThread);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return fail(PNCOV_ENOMEM, "out of device memory");
    }
    if (e != cudaSuccess) return fail_cuda(e, "cudaMallocAsync");
    return 0;
}

PNCOV_API int pncov_device_free(int device, void* ptr) {
    if (!ptr) return 0;
    DeviceGuard guard(device);
    // the buffer may have been used on any stream: drain the device (what cudaFree did implicitly) before the block
    // goes back to the pool
    PN_CUDA(cudaDeviceSynchronize());
    pn_dev_free(ptr, cudaStreamPerThread);
    return 0;
}

PNCOV_API int pncov_copy_to_host(int device, void* dst, const void* src, int64_t nbytes, void* stream_) {
    if (nbytes < 0 || (nbytes > 0 && (!dst || !src))) return fail(PNCOV_EINVAL, "bad arguments");
    if (nbytes == 0) return 0;
    DeviceGuard guard(device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PN_CUDA(cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDeviceToHost, stream));
    PN_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

PNCOV_API int pncov_combine_device(int device, void* dst, const void* src, int64_t n, int op, double alpha, double beta,
                                   void* stream_) {
    if (n < 0 || (n > 0 && !dst)) return fail(PNCOV_EINVAL, "bad arguments");
    if (op != PNCOV_COMBINE_LINEAR && op != PNCOV_COMBINE_MULTIPLY) return fail(PNCOV_EINVAL, "unknown combine op");
    if (op == PNCOV_COMBINE_MULTIPLY && !src) return fail(PNCOV_EINVAL, "multiply needs a source");
    if (op == PNCOV_COMBINE_LINEAR && !src && beta != 0.0) return fail(PNCOV_EINVAL, "beta != 0 needs a source");
    if (n == 0) return 0;
    DeviceGuard guard(device);
    PN_CUDA(pn_launch_combine(static_cast<double*>(dst), static_cast<const double*>(src), n, op, alpha, beta,
                              static_cast<cudaStream_t>(stream_)));
    return 0;
}

// ---- Cholesky factor of a Gram matrix + jitter I, resident on the device ---------------------------
// n = 16384: a pass of the chained substitution (1-3 columns, two launches) takes 1.4-1.7 ms, the transposed DMMA
// substitution 9.4 ms for anything up to 128 columns; 4.3 against 0.9 ms at n = 8192, 2.3 against 0.45 ms at n = 4096
// (tools/bench_cholesky.py): four passes are still ahead of it
constexpr int64_t kWideSolveMinRhs = 13;

struct pncov_chol {
    int device = 0;
    int64_t n = 0, npad = 0;
    double* M = nullptr;      // (npad, npad) row-major; lower triangle = L after pncov_chol_factor
    double* panel = nullptr;  // 2 x (npad, 128): contiguous block columns of two consecutive steps
    double* invd = nullptr;   // (npad / 128) inverse diagonal blocks, 128 x 128 each
    double* rhs = nullptr;    // (npad, kWideSolveMinRhs - 1)
    int* info = nullptr;
    int* solve_sync = nullptr;  // 2 + 2 npad / 128: tickets and flags of the chained substitution
    cudaStream_t stream = nullptr, side = nullptr;  // side: look-ahead factorisation of the next diagonal block
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_panel = nullptr, ev_diag = nullptr;
    bool factored = false;
};

PNCOV_API int pncov_chol_create(pncov_chol** out, int device, int64_t n) {
    if (!out) return fail(PNCOV_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 1) return fail(PNCOV_EINVAL, "matrix order must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PNCOV_ENODEVICE, "no CUDA device available (libpncov has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(PNCOV_EINVAL, "device index out of range");
    pncov_chol* c = new (std::nothrow) pncov_chol();
    if (!c) return fail(PNCOV_ENOMEM, "out of host memory");
    c->device = device;
    c->n = n;
    c->npad = ((n + 127) / 128) * 128;
    DeviceGuard guard(device);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {  // look-ahead stream: highest priority, so its single CTA is dispatched first
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_panel, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_diag, cudaEventDisableTiming);
    // from the stream-ordered pool (kept up to its release threshold): a cudaMalloc / cudaFree pair of the 2.1 GB
    // factor buffer cost 0.03 + 0.6 s per Bayesian-quadrature update at 16384 nodes
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->M), (size_t)c->npad * c->npad * 8, c->stream);
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->panel), (size_t)2 * c->npad * 128 * 8, c->stream);
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->invd), (size_t)c->npad * 128 * 8, c->stream);
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->rhs), (size_t)c->npad * (kWideSolveMinRhs - 1) * 8, c->stream);
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->info), sizeof(int), c->stream);
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->solve_sync), sizeof(int) * (2 + 2 * (size_t)(c->npad / 128)), c->stream);
    if (e != cudaSuccess) {
        const bool oom = e == cudaErrorMemoryAllocation;
        int rc = oom ? fail(PNCOV_ENOMEM, "out of device memory") : fail_cuda(e, "pncov_chol_create");
        cudaGetLastError();
        pncov_chol_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

PNCOV_API int pncov_chol_destroy(pncov_chol* c) {
    if (!c) return 0;
    DeviceGuard guard(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->side) cudaStreamSynchronize(c->side);
    pn_dev_free(c->M, c->stream);
    pn_dev_free(c->panel, c->stream);
    pn_dev_free(c->invd, c->stream);
    pn_dev_free(c->rhs, c->stream);
    pn_dev_free(c->info, c->stream);
    pn_dev_free(c->solve_sync, c->stream);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_panel) cudaEventDestroy(c->ev_panel);
    if (c->ev_diag) cudaEventDestroy(c->ev_diag);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

PNCOV_API int pncov_chol_factor(pncov_chol* c, const double* a, int64_t ld, double jitter, int a_on_device, int* info) {
    if (!c || !a || !info) return fail(PNCOV_EINVAL, "NULL argument");
    if (ld < c->n) return fail(PNCOV_EINVAL, "leading dimension smaller than the matrix order");
    DeviceGuard guard(c->device);
    c->factored = false;
    // a device-resident input may still be in flight on the caller's stream: the handle works on
    // its own stream, so wait for the device once (microseconds next to the factorisation)
    if (a_on_device < 0 || a_on_device > 2) return fail(PNCOV_EINVAL, "a_on_device must be 0, 1 or 2");
    if (a_on_device == 1) PN_CUDA(cudaDeviceSynchronize());
    if (c->npad != c->n) PN_CUDA(cudaMemsetAsync(c->M, 0, (size_t)c->npad * c->npad * 8, c->stream));
    if (a_on_device == 2) {
        // host array of which the UPPER triangle counts (scipy.linalg.cho_factor's default, lower=False): only the
        // blocks on and above the diagonal cross PCIe -- half the bytes of the dominant step of a factorisation
        // from a host Gram matrix -- and the lower triangle the kernels work on is mirrored from them on the device
        const int64_t n = c->n, rb = 512;
        for (int64_t r0 = 0; r0 < n; r0 += rb) {
            const int64_t rows = std::min<int64_t>(rb, n - r0);
            PN_CUDA(cudaMemcpy2DAsync(c->M + r0 * c->npad + r0, (size_t)c->npad * 8, a + r0 * ld + r0, (size_t)ld * 8,
                                      (size_t)(n - r0) * 8, (size_t)rows, cudaMemcpyHostToDevice, c->stream));
        }
        PN_CUDA(pn_chol_mirror_upper(c->M, c->npad, n, c->stream));
    } else {
        PN_CUDA(cudaMemcpy2DAsync(c->M, (size_t)c->npad * 8, a, (size_t)ld * 8, (size_t)c->n * 8, (size_t)c->n,
                                  a_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    }
    PN_CUDA(cudaEventRecord(c->ev0, c->stream));
    PN_CUDA(pn_chol_factor(c->M, c->n, c->npad, jitter, c->panel, c->invd, c->info, c->solve_sync, c->stream, c->side, c->ev_panel,
                           c->ev_diag));
    PN_CUDA(cudaEventRecord(c->ev1, c->stream));
    PN_CUDA(cudaMemcpyAsync(info, c->info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    if (*info < 0) return fail(PNCOV_EINVAL, "internal: a hand-over wait of the factorisation gave up");
    c->factored = *info == 0;
    return 0;
}

PNCOV_API int pncov_chol_factor_gram(pncov_chol* c, pncov_handle* h, double jitter, int* info) {
    if (!c || !h || !info) return fail(PNCOV_EINVAL, "NULL argument");
    if (h->device != c->device) return fail(PNCOV_EINVAL, "operator and Cholesky handle live on different devices");
    if (h->p0.n != c->n || h->p1.n != c->n) return fail(PNCOV_EINVAL, "the operator is not square of the handle's order");
    if (h->k.gram_elem() != 8) return fail(PNCOV_EINVAL, "float64 Gram matrices only");
    DeviceGuard guard(c->device);
    c->factored = false;
    // the Gram matrix is written straight into the factorisation buffer: it never visits the host
    if (c->npad != c->n) PN_CUDA(cudaMemsetAsync(c->M, 0, (size_t)c->npad * c->npad * 8, c->stream));
    PN_CUDA(launch_gram_any(h, false, 0, c->n, c->M, c->npad, c->stream));
    PN_CUDA(cudaEventRecord(c->ev0, c->stream));
    PN_CUDA(pn_chol_factor(c->M, c->n, c->npad, jitter, c->panel, c->invd, c->info, c->solve_sync, c->stream, c->side, c->ev_panel,
                           c->ev_diag));
    PN_CUDA(cudaEventRecord(c->ev1, c->stream));
    PN_CUDA(cudaMemcpyAsync(info, c->info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    if (*info < 0) return fail(PNCOV_EINVAL, "internal: a hand-over wait of the factorisation gave up");
    c->factored = *info == 0;
    return 0;
}

PNCOV_API int pncov_chol_solve(pncov_chol* c, const double* b, int64_t nrhs, int64_t ldb, double* x, int64_t ldx) {
    if (!c || !b || !x) return fail(PNCOV_EINVAL, "NULL argument");
    if (!c->factored) return fail(PNCOV_EINVAL, "no successful factorisation in this handle");
    if (nrhs < 0 || ldb < nrhs || ldx < nrhs) return fail(PNCOV_EINVAL, "bad right-hand-side shape");
    DeviceGuard guard(c->device);
    if (nrhs >= kWideSolveMinRhs) {
        // many right-hand sides: transposed on the device and solved by the blocked DMMA substitution of
        // the prediction path (pn_chol_solve_wide), in column chunks of at most 512 MiB per buffer
        const int64_t n = c->n, npad = c->npad;
        int64_t chunk = std::max<int64_t>(128, ((int64_t)(512ll << 20) / (npad * 8)) / 128 * 128);
        chunk = std::min<int64_t>(chunk, (nrhs + 127) / 128 * 128);
        void *B = nullptr, *Z = nullptr, *Y = nullptr;
        cudaStream_t s = c->stream;
        auto run = [&]() -> cudaError_t {
            cudaError_t e;
            if ((e = pn_dev_alloc(&B, (size_t)n * chunk * 8, s)) != cudaSuccess) return e;
            if ((e = pn_dev_alloc(&Z, (size_t)chunk * npad * 8, s)) != cudaSuccess) return e;
            if ((e = pn_dev_alloc(&Y, (size_t)chunk * npad * 8, s)) != cudaSuccess) return e;
            for (int64_t c0 = 0; c0 < nrhs; c0 += chunk) {
                const int64_t mc = std::min<int64_t>(chunk, nrhs - c0), mcpad = (mc + 127) / 128 * 128;
                // contiguous operands go as one linear copy (pitched copies of short rows from pageable memory are slow)
                if (ldb == mc) e = cudaMemcpyAsync(B, b, (size_t)n * mc * 8, cudaMemcpyHostToDevice, s);
                else e = cudaMemcpy2DAsync(B, (size_t)mc * 8, b + c0, (size_t)ldb * 8, (size_t)mc * 8, (size_t)n,
                                           cudaMemcpyHostToDevice, s);
                if (e != cudaSuccess) return e;
                if ((e = cudaMemsetAsync(Z, 0, (size_t)mcpad * npad * 8, s)) != cudaSuccess) return e;
                if ((e = pn_chol_transpose(static_cast<double*>(B), mc, n, mc, static_cast<double*>(Z), npad, s)) != cudaSuccess) return e;
                if ((e = pn_chol_solve_wide(c->M, npad, c->invd, static_cast<double*>(Z), static_cast<double*>(Y), mcpad, 0, s,
                                            c->side, c->ev_panel, c->ev_diag)) != cudaSuccess)
                    return e;
                if ((e = pn_chol_transpose(static_cast<double*>(Z), npad, mc, n, static_cast<double*>(B), mc, s)) != cudaSuccess) return e;
                if (ldx == mc) e = cudaMemcpyAsync(x, B, (size_t)n * mc * 8, cudaMemcpyDeviceToHost, s);
                else e = cudaMemcpy2DAsync(x + c0, (size_t)ldx * 8, B, (size_t)mc * 8, (size_t)mc * 8, (size_t)n,
                                           cudaMemcpyDeviceToHost, s);
                if (e != cudaSuccess) return e;
            }
            return cudaStreamSynchronize(s);
        };
        const cudaError_t e = run();
        pn_dev_free(B, s);
        pn_dev_free(Z, s);
        pn_dev_free(Y, s);
        if (e != cudaSuccess) return fail_cuda(e, "pncov_chol_solve (wide)");
        return 0;
    }
    // up to kWideSolveMinRhs - 1 columns: the whole block goes to the device once (one linear copy when contiguous:
    // a pitched copy of 16-byte rows from pageable memory costs milliseconds); the chained substitution then runs
    // over it in place, three adjacent columns per pass
    if (nrhs == 0) return 0;
    if (c->npad != c->n) PN_CUDA(cudaMemsetAsync(c->rhs, 0, (size_t)c->npad * nrhs * 8, c->stream));
    if (ldb == nrhs)
        PN_CUDA(cudaMemcpyAsync(c->rhs, b, (size_t)c->n * nrhs * 8, cudaMemcpyHostToDevice, c->stream));
    else
        PN_CUDA(cudaMemcpy2DAsync(c->rhs, (size_t)nrhs * 8, b, (size_t)ldb * 8, (size_t)nrhs * 8, (size_t)c->n,
                                  cudaMemcpyHostToDevice, c->stream));
    for (int64_t c0 = 0; c0 < nrhs; c0 += 3)
        PN_CUDA(pn_chol_solve(c->M, c->npad, c->invd, c->rhs + c0, nrhs, (int)std::min<int64_t>(3, nrhs - c0), c->solve_sync,
                              c->stream));
    if (ldx == nrhs)
        PN_CUDA(cudaMemcpyAsync(x, c->rhs, (size_t)c->n * nrhs * 8, cudaMemcpyDeviceToHost, c->stream));
    else
        PN_CUDA(cudaMemcpy2DAsync(x, (size_t)ldx * 8, c->rhs, (size_t)nrhs * 8, (size_t)nrhs * 8, (size_t)c->n,
                                  cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

PNCOV_API int pncov_chol_solve_gram(pncov_chol* c, pncov_handle* h, int64_t nvec, const double* vecs, int64_t ldv,
                                    double* dots, int64_t ldd, double* quad, int mode, float* ms) {
    if (!c || !h || !quad) return fail(PNCOV_EINVAL, "NULL argument");
    if (!c->factored) return fail(PNCOV_EINVAL, "no successful factorisation in this handle");
    if (h->device != c->device) return fail(PNCOV_EINVAL, "operator and Cholesky handle live on different devices");
    if (h->p0.n != c->n) return fail(PNCOV_EINVAL, "the operator's first point set is not the factorised one (n0 != n)");
    if (h->k.gram_elem() != 8) return fail(PNCOV_EINVAL, "float64 Gram matrices only");
    if (nvec < 0 || nvec > 4) return fail(PNCOV_EINVAL, "at most 4 vectors");
    if (mode != PNCOV_SOLVE_GRAM_REFERENCE_ORDER && mode != PNCOV_SOLVE_GRAM_FORWARD_ONLY)
        return fail(PNCOV_EINVAL, "unknown mode");
    const bool fwd = mode == PNCOV_SOLVE_GRAM_FORWARD_ONLY;
    if (nvec > 0 && (!vecs || !dots || ldv < c->n)) return fail(PNCOV_EINVAL, "bad vector arguments");
    const int64_t m = h->p1.n, n = c->n, npad = c->npad;
    if (nvec > 0 && ldd < m) return fail(PNCOV_EINVAL, "ldd smaller than the number of prediction points");
    if (ms) *ms = 0.f;
    if (m == 0) return 0;
    DeviceGuard guard(c->device);
    // row chunks of the transposed right-hand sides: two (chunk, npad) buffers of at most 512 MiB each
    int64_t chunk = std::max<int64_t>(128, ((int64_t)(512ll << 20) / (npad * 8)) / 128 * 128);
    chunk = std::min<int64_t>(chunk, (m + 127) / 128 * 128);
    void *Z = nullptr, *Y = nullptr, *vd = nullptr, *qd = nullptr, *dd = nullptr;
    cudaStream_t s = c->stream;
    auto release = [&]() {
        pn_dev_free(Z, s);
        pn_dev_free(Y, s);
        pn_dev_free(vd, s);
        pn_dev_free(qd, s);
        pn_dev_free(dd, s);
    };
    auto run = [&]() -> cudaError_t {
        cudaError_t e;
        if ((e = pn_dev_alloc(&Z, (size_t)chunk * npad * 8, s)) != cudaSuccess) return e;
        if ((e = pn_dev_alloc(&Y, (size_t)chunk * npad * 8, s)) != cudaSuccess) return e;
        if ((e = pn_dev_alloc(&qd, (size_t)m * 8, s)) != cudaSuccess) return e;
        if (nvec > 0) {
            // forward-only: vd holds the rows (L^-1 v_t)^T, produced by a 128-row substitution pass of its own
            if ((e = pn_dev_alloc(&vd, (size_t)(fwd ? 128 : nvec) * npad * 8, s)) != cudaSuccess) return e;
            if ((e = pn_dev_alloc(&dd, (size_t)nvec * m * 8, s)) != cudaSuccess) return e;
            void* dst = fwd ? Z : vd;
            if (fwd && (e = cudaMemsetAsync(Z, 0, (size_t)128 * npad * 8, s)) != cudaSuccess) return e;
            if ((e = cudaMemcpy2DAsync(dst, (size_t)npad * 8, vecs, (size_t)ldv * 8, (size_t)n * 8, (size_t)nvec,
                                       cudaMemcpyHostToDevice, s)) != cudaSuccess)
                return e;
        }
        if ((e = cudaEventRecord(c->ev0, s)) != cudaSuccess) return e;
        if (fwd && nvec > 0 &&
            (e = pn_chol_solve_wide(c->M, npad, c->invd, static_cast<double*>(Z), static_cast<double*>(vd), 128, 1, s, c->side, c->ev_panel, c->ev_diag)) != cudaSuccess)
            return e;
        for (int64_t r0 = 0; r0 < m; r0 += chunk) {
            const int64_t mc = std::min<int64_t>(chunk, m - r0), mcpad = (mc + 127) / 128 * 128;
            // padded columns must be exact zeros (they meet the identity block of the padded factor)
            if (npad != n && (e = cudaMemsetAsync(Z, 0, (size_t)mcpad * npad * 8, s)) != cudaSuccess) return e;
            if ((e = launch_gram_any(h, true, r0, mc, Z, npad, s)) != cudaSuccess) return e;  // rows: prediction points
            if ((e = pn_chol_solve_wide(c->M, npad, c->invd, static_cast<double*>(Z), static_cast<double*>(Y), mcpad, fwd ? 1 : 0, s, c->side, c->ev_panel, c->ev_diag)) != cudaSuccess) return e;
            if (fwd) {  // rows of Y = (L^-1 g_j)^T: quad = |y_j|^2, dots = (L^-1 v_t) . y_j
                if ((e = pn_chol_rowdots(static_cast<double*>(Y), static_cast<double*>(Y), npad, n, mc, static_cast<double*>(vd),
                                         (int)nvec, static_cast<double*>(qd) + r0, static_cast<double*>(dd) + r0, m, s)) != cudaSuccess)
                    return e;
                continue;
            }
            if ((e = launch_gram_any(h, true, r0, mc, Y, npad, s)) != cudaSuccess) return e;  // the Gram block again
            if ((e = pn_chol_rowdots(static_cast<double*>(Z), static_cast<double*>(Y), npad, n, mc, static_cast<double*>(vd),
                                     (int)nvec, static_cast<double*>(qd) + r0, static_cast<double*>(dd) + r0, m, s)) != cudaSuccess)
                return e;
        }
        if ((e = cudaEventRecord(c->ev1, s)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(quad, qd, (size_t)m * 8, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        if (nvec > 0 && (e = cudaMemcpy2DAsync(dots, (size_t)ldd * 8, dd, (size_t)m * 8, (size_t)m * 8, (size_t)nvec,
                                               cudaMemcpyDeviceToHost, s)) != cudaSuccess)
            return e;
        return cudaStreamSynchronize(s);
    };
    const cudaError_t e = run();
    release();
    if (e != cudaSuccess) return fail_cuda(e, "pncov_chol_solve_gram");
    if (ms) PN_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return 0;
}

PNCOV_API int pncov_chol_get_factor(pncov_chol* c, double* l, int64_t ld) {
    if (!c || !l) return fail(PNCOV_EINVAL, "NULL argument");
    if (!c->factored) return fail(PNCOV_EINVAL, "no successful factorisation in this handle");
    if (ld < c->n) return fail(PNCOV_EINVAL, "leading dimension smaller than the matrix order");
    DeviceGuard guard(c->device);
    PN_CUDA(cudaMemcpy2DAsync(l, (size_t)ld * 8, c->M, (size_t)c->npad * 8, (size_t)c->n * 8, (size_t)c->n,
                              cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

PNCOV_API int pncov_chol_last_factor_ms(pncov_chol* c, float* ms) {
    if (!c || !ms) return fail(PNCOV_EINVAL, "NULL argument");
    DeviceGuard guard(c->device);
    PN_CUDA(cudaEventSynchronize(c->ev1));
    PN_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return 0;
}

// ---- update chains of the matrix-based solver (SURVEY.md 8f N2 ii) -------------------------------------
struct pncov_chain {
    int device = 0;
    int64_t n = 0, capacity = 0, ld = 0;
    double *U = nullptr, *V = nullptr;  // (capacity, ld) row-major: one vector per row
    double *x = nullptr, *y = nullptr, *partial = nullptr;
    cudaStream_t stream = nullptr;
};
constexpr int kChainSplits = 32;

PNCOV_API int pncov_chain_destroy(pncov_chain* c) {
    if (!c) return 0;
    DeviceGuard guard(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    pn_dev_free(c->U, c->stream);
    pn_dev_free(c->V, c->stream);
    pn_dev_free(c->x, c->stream);
    pn_dev_free(c->y, c->stream);
    pn_dev_free(c->partial, c->stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

PNCOV_API int pncov_chain_create(pncov_chain** out, int device, int64_t n, int64_t capacity) {
    if (!out) return fail(PNCOV_EINVAL, "out is NULL");
    *out = nullptr;
    // 2 * capacity coefficients live in the shared memory of chain_apply_kernel (at most 200 KB)
    if (n < 1 || capacity < 1 || capacity > 12800) return fail(PNCOV_EINVAL, "bad chain shape (capacity is at most 12800 updates)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PNCOV_ENODEVICE, "no CUDA device available (libpncov has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(PNCOV_EINVAL, "device index out of range");
    pncov_chain* c = new (std::nothrow) pncov_chain();
    if (!c) return fail(PNCOV_ENOMEM, "out of host memory");
    c->device = device;
    c->n = n;
    c->capacity = capacity;
    c->ld = (n + 1) / 2 * 2;
    DeviceGuard guard(device);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    const size_t mat = (size_t)capacity * c->ld * 8;
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->U), mat, c->stream);
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->V), mat, c->stream);
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->x), (size_t)c->ld * 8, c->stream);
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->y), (size_t)c->ld * 8, c->stream);
    if (e == cudaSuccess) e = pn_dev_alloc(reinterpret_cast<void**>(&c->partial), (size_t)2 * capacity * kChainSplits * 8, c->stream);
    if (e != cudaSuccess) {
        pncov_chain_destroy(c);
        return fail_cuda(e, "pncov_chain_create");
    }
    *out = c;
    return 0;
}

PNCOV_API int pncov_chain_set(pncov_chain* c, int64_t index, const double* u, const double* v) {
    if (!c || !u || !v) return fail(PNCOV_EINVAL, "NULL argument");
    if (index < 0 || index >= c->capacity) return fail(PNCOV_EINVAL, "chain index out of range");
    DeviceGuard guard(c->device);
    PN_CUDA(cudaMemcpyAsync(c->U + index * c->ld, u, (size_t)c->n * 8, cudaMemcpyHostToDevice, c->stream));
    PN_CUDA(cudaMemcpyAsync(c->V + index * c->ld, v, (size_t)c->n * 8, cudaMemcpyHostToDevice, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));  // the host vectors may be freed after the call
    return 0;
}

PNCOV_API int pncov_chain_apply(pncov_chain* c, int64_t k, int mode, const double* x, int64_t ncols, int64_t ldx, double* y,
                                int64_t ldy) {
    if (!c || !x || !y) return fail(PNCOV_EINVAL, "NULL argument");
    if (k < 0 || k > c->capacity) return fail(PNCOV_EINVAL, "chain length out of range");
    if (mode < 0 || mode > 2) return fail(PNCOV_EINVAL, "unknown mode");
    if (ncols < 0 || ldx < ncols || ldy < ncols) return fail(PNCOV_EINVAL, "bad operand shape");
    DeviceGuard guard(c->device);
    for (int64_t col = 0; col < ncols; ++col) {
        if (ldx == 1)
            PN_CUDA(cudaMemcpyAsync(c->x, x, (size_t)c->n * 8, cudaMemcpyHostToDevice, c->stream));
        else
            PN_CUDA(cudaMemcpy2DAsync(c->x, 8, x + col, (size_t)ldx * 8, 8, (size_t)c->n, cudaMemcpyHostToDevice, c->stream));
        PN_CUDA(pn_launch_chain_apply(c->U, c->V, c->ld, c->n, (int)k, mode, c->x, c->y, c->partial, kChainSplits, c->stream));
        if (ldy == 1)
            PN_CUDA(cudaMemcpyAsync(y, c->y, (size_t)c->n * 8, cudaMemcpyDeviceToHost, c->stream));
        else
            PN_CUDA(cudaMemcpy2DAsync(y + col, (size_t)ldy * 8, c->y, 8, 8, (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
    }
    PN_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

PNCOV_API int64_t pncov_launch_count(void) { return g_pn_launches; }

PNCOV_API int64_t pncov_d2h_bytes(void) { return g_pn_d2h_bytes; }

PNCOV_API int pncov_set_option(pncov_handle* h, int option, int value) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (value < -1 || value > 1) return fail(PNCOV_EINVAL, "option values are -1 (default), 0 (off) or 1 (on)");
    switch (option) {
        case PNCOV_OPT_IDENTITY: h->opt_identity = value; return 0;  // gates the use, not the admissibility test
        case PNCOV_OPT_SYMMETRIC: h->opt_symmetric = value; return 0;
        case PNCOV_OPT_FACTORISED: h->opt_factorised = value; return 0;
        default: return fail(PNCOV_EINVAL, "unknown option");
    }
}

PNCOV_API int pncov_set_program(pncov_handle* h, int n_leaves, const int* kinds, const double* lengthscales, const double* nus,
                                const double* alphas, int n_steps, const int* ops, const int* leaf_index, const double* a,
                                const double* b) {
    if (!h) return fail(PNCOV_EINVAL, "handle is NULL");
    if (n_leaves == 0 && n_steps == 0) {  // back to the handle's own kernel
        h->has_prog = false;
        return 0;
    }
    if (n_leaves < 1 || n_leaves > PN_FUSED_MAX_LEAVES || n_steps < 1 || n_steps > PN_FUSED_MAX_STEPS)
        return fail(PNCOV_ENOTIMPL, "programs hold 1..6 leaves and 1..12 steps");
    if (!kinds || !lengthscales || !nus || !alphas || !ops || !leaf_index || !a || !b) return fail(PNCOV_EINVAL, "NULL argument");
    if (h->k.dtype != PNCOV_F64 || h->k.ard || h->k.d > 16)
        return fail(PNCOV_ENOTIMPL, "programs need float64 points, isotropic leaves and d <= 16");
    PnFusedSlot leaves[PN_FUSED_MAX_LEAVES] = {};
    for (int l = 0; l < n_leaves; ++l) {
        PnFusedSlot& L = leaves[l];
        if (kinds[l] == PNCOV_WHITENOISE) {  // _white_noise.py:38-55; sigma^2 travels in `alphas`
            if (!(alphas[l] >= 0)) return fail(PNCOV_EINVAL, "sigma_sq must be non-negative");
            L.whitenoise = 1;
            L.sigma_sq = alphas[l];
            continue;
        }
        if (kinds[l] != PNCOV_EXPQUAD && kinds[l] != PNCOV_MATERN && kinds[l] != PNCOV_RATQUAD)
            return fail(PNCOV_ENOTIMPL, "program leaves are ExpQuad, Matern, RatQuad and WhiteNoise");
        KernelDesc kd;
        int rc = build_desc(kd, kinds[l], PNCOV_F64, h->k.d, &lengthscales[l], 1, nus[l], alphas[l]);
        if (rc) return rc;
        L.epi = kd.epi;
    }
    // regroup by leaf evaluation: a LOAD / MULLEAF opens a slot, the steps after it -- [SCALEVAL] [SET | LIN | MUL]
    // [SCALEACC], each at most once and in this order -- become its canonical parameters
    PnFusedProgram F{};
    for (int i = 0; i < n_steps; ++i) {
        if (ops[i] < PN_F_LOAD || ops[i] > PN_F_SCALEACC) return fail(PNCOV_EINVAL, "unknown program step");
        if (ops[i] <= PN_F_MULLEAF) {
            if (leaf_index[i] < 0 || leaf_index[i] >= n_leaves) return fail(PNCOV_EINVAL, "leaf index out of range");
            if (F.n_slots == PN_FUSED_MAX_LEAVES) return fail(PNCOV_ENOTIMPL, "programs evaluate at most 6 leaves");
            PnFusedSlot& S = F.slots[F.n_slots++];
            S = leaves[leaf_index[i]];
            S.mul = ops[i] == PN_F_MULLEAF ? 1 : 0;
            if (S.mul && F.n_slots == 1) return fail(PNCOV_EINVAL, "a program starts with PNCOV_F_LOAD");
            S.kind = 1;  // no update of acc until a step says otherwise: fma(0, val, 1 * acc)
            S.a = 1.0;
            S.b = 0.0;
            S.s_val = 1.0;
            S.s_acc = 1.0;
            S.n_post = 0;
            if (S.whitenoise) F.has_wn = 1;
            if (!S.whitenoise && S.epi.epi == PN_EPI_RATQUADG) F.need_log = 1;
        } else {
            if (F.n_slots == 0) return fail(PNCOV_EINVAL, "a program starts with PNCOV_F_LOAD");
            PnFusedSlot& S = F.slots[F.n_slots - 1];
            // stage of the canonical form this step belongs to; stages must not go backwards or repeat
            const int stage = ops[i] == PN_F_SCALEVAL ? 1 : (ops[i] == PN_F_SCALEACC ? 3 : 2);
            int last = 0;
            for (int j = 0; j < S.n_post; ++j) {
                const int o = S.post[j].op;
                last = o == PN_F_SCALEVAL ? 1 : (o == PN_F_SCALEACC ? 3 : 2);
            }
            if (stage <= last || S.n_post == 4)
                return fail(PNCOV_ENOTIMPL, "between two leaf evaluations: [SCALEVAL] [SET | LIN | MUL] [SCALEACC], each at most once");
            PnFusedStep& st = S.post[S.n_post++];
            st.op = ops[i];
            st.a = a[i];
            st.b = b[i];
            if (ops[i] == PN_F_SCALEVAL) S.s_val = a[i];
            else if (ops[i] == PN_F_SCALEACC) S.s_acc = a[i];
            else if (ops[i] == PN_F_SET) { S.kind = 1; S.a = 0.0; S.b = 1.0; }
            else if (ops[i] == PN_F_LIN) { S.kind = 1; S.a = a[i]; S.b = b[i]; }
            else S.kind = 2;
        }
    }
    // a WhiteNoise leaf that is simply added (LOAD, LIN(1, b)) after the first slot touches only entries of equal points
    for (int l = 1; l < F.n_slots; ++l) {
        PnFusedSlot& S = F.slots[l];
        const bool next_is_product = l + 1 < F.n_slots && F.slots[l + 1].mul;
        if (S.whitenoise && !S.mul && S.kind == 1 && S.a == 1.0 && S.s_val == 1.0 && S.s_acc == 1.0 && !next_is_product)
            S.whitenoise = 2;
    }
    h->prog = F;
    h->has_prog = true;
    // `[o *] k [+ WhiteNoise]` (either order): one isotropic leaf, folded into the compile-time specialised kernels
    h->prog_simple = false;
    {
        const PnFusedSlot* leaf = nullptr;
        const PnFusedSlot* wn = nullptr;
        double scale = 1.0;
        bool scaled = false, ok = F.n_slots >= 1 && F.n_slots <= 2;
        for (int l = 0; ok && l < F.n_slots; ++l) {
            const PnFusedSlot& S = F.slots[l];
            if (S.mul) ok = false;
            if (S.whitenoise) {
                if (wn) ok = false;
                wn = &S;
            } else {
                if (leaf) ok = false;
                leaf = &S;
            }
            // first slot: SET [, SCALEACC o]; second slot: LIN(1, o) (o = 1 for the WhiteNoise term itself)
            if (l == 0) {
                if (S.n_post < 1 || S.n_post > 2 || S.post[0].op != PN_F_SET) ok = false;
                if (ok && S.n_post == 2) {
                    if (S.post[1].op != PN_F_SCALEACC || S.whitenoise) ok = false;
                    scale = S.post[1].a;
                    scaled = true;
                }
            } else {
                if (S.n_post != 1 || S.post[0].op != PN_F_LIN || S.post[0].a != 1.0) ok = false;
                if (ok && S.whitenoise && S.post[0].b != 1.0) ok = false;
                if (ok && !S.whitenoise && S.post[0].b != 1.0) {
                    scale = S.post[0].b;  // acc = fma(o, k, sigma^2 [x0 == x1]): same bits as o * k (+ sigma^2 on equal points)
                    scaled = true;
                }
            }
        }
        if (ok && leaf && h->k.d <= 8) {
            h->prog_simple = true;
            h->prog_epi = leaf->epi;
            h->prog_epi.post_flags = (scaled ? 1 : 0) | (wn ? 2 : 0);
            h->prog_epi.post_scale = scale;
            h->prog_epi.post_sigma_sq = wn ? wn->sigma_sq : 0.0;
        }
    }
    return 0;
}

PNCOV_API int pncov_matmul_path(const pncov_handle* h, int* path) {
    if (!h || !path) return fail(PNCOV_EINVAL, "NULL argument");
    if (h->last_use_mma < 0) {
        *path = 0;
        return 0;
    }
    const bool ident = h->last_use_mma != 0;
    *path = ident ? (h->last_sym ? 3 : 1) : (h->last_sym ? 4 : 2);
    return 0;
}

PNCOV_API int pncov_last_kernel_ms(const pncov_handle* h, float* ms) {
    if (!h || !ms) return fail(PNCOV_EINVAL, "NULL argument");
    if (!h->timed) return fail(PNCOV_EINVAL, "no timed call yet");
    DeviceGuard guard(h->device);
    PN_CUDA(cudaEventSynchronize(h->ev1));
    PN_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return 0;
}


// ---- one operator over several GPUs of this process (no torch, no launcher) --------------------------------------
struct pncov_sharded {
    int n_dev = 0;
    int64_t n0 = 0, n1 = 0, d = 0;
    size_t es = 8;
    bool symmetric = false;
    std::vector<int> devices;
    std::vector<int64_t> lo, hi;          // row block of every device
    std::vector<pncov_handle*> rows;      // K(x0[lo:hi], x1) on device i
    std::vector<pncov_handle*> whole;     // symmetric only: K(x0, x0) on device i (round-robin row blocks of the triangle)
    std::vector<void*> v_dev, bins_dev;   // symmetric one-column products: right-hand side and integer planes per device
    void* y_dev = nullptr;
    int whole_state = 0;                  // 0 = untried, 1 = symmetric kernel serves, -1 = it does not
};

namespace {

// run fn(i) for every device on its own thread; first failure wins (its message becomes this thread's last error)
int for_each_device(int n, const std::function<int(int)>& fn) {
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i)
        th.emplace_back([&, i] {
            rc[i] = fn(i);
            if (rc[i]) msg[i] = t_err;
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (rc[i]) return fail(rc[i], msg[i]);
    return 0;
}

}  // namespace

PNCOV_API int pncov_sharded_destroy(pncov_sharded* s) {
    if (!s) return 0;
    for (int i = 0; i < s->n_dev; ++i) {
        if (i < (int)s->rows.size()) pncov_destroy(s->rows[i]);
        if (i < (int)s->whole.size()) pncov_destroy(s->whole[i]);
        if (i < (int)s->v_dev.size() && s->v_dev[i]) pncov_device_free(s->devices[i], s->v_dev[i]);
        if (i < (int)s->bins_dev.size() && s->bins_dev[i]) pncov_device_free(s->devices[i], s->bins_dev[i]);
    }
    if (s->y_dev) pncov_device_free(s->devices[0], s->y_dev);
    delete s;
    return 0;
}

PNCOV_API int pncov_sharded_create(pncov_sharded** out, const int* device_ids, int n_dev, int kind, int dtype, int64_t d,
                                   const double* lengthscales, int64_t n_lengthscales, double nu, double alpha, const void* x0,
                                   int64_t n0, const void* x1, int64_t n1) {
    if (!out) return fail(PNCOV_EINVAL, "out is NULL");
    *out = nullptr;
    if (!device_ids || n_dev < 1 || n_dev > 64) return fail(PNCOV_EINVAL, "1 .. 64 device ids");
    if (n0 < 0 || (x1 && n1 < 0) || (n0 > 0 && !x0)) return fail(PNCOV_EINVAL, "bad point sets");
    if (dtype != PNCOV_F64 && dtype != PNCOV_F32) return fail(PNCOV_EINVAL, "unknown dtype");
    pncov_sharded* s = new (std::nothrow) pncov_sharded();
    if (!s) return fail(PNCOV_ENOMEM, "out of host memory");
    s->n_dev = n_dev;
    s->n0 = n0;
    s->n1 = x1 ? n1 : n0;
    s->d = d;
    s->es = dtype == PNCOV_F32 ? 4 : 8;
    s->symmetric = x1 == nullptr;
    s->devices.assign(device_ids, device_ids + n_dev);
    s->rows.assign(n_dev, nullptr);
    s->whole.assign(n_dev, nullptr);
    s->v_dev.assign(n_dev, nullptr);
    s->bins_dev.assign(n_dev, nullptr);
    const int64_t per = (n0 + n_dev - 1) / n_dev;
    for (int i = 0; i < n_dev; ++i) {
        s->lo.push_back(std::min<int64_t>(n0, i * per));
        s->hi.push_back(std::min<int64_t>(n0, (i + 1) * per));
    }
    const void* cols = x1 ? x1 : x0;
    const int64_t ncols = s->n1;
    int rc = for_each_device(n_dev, [&](int i) {
        const char* block = static_cast<const char*>(x0) + (size_t)s->lo[i] * d * s->es;
        // one device: keep x1 = NULL so that the symmetric kernels serve the whole operator
        if (n_dev == 1 && s->symmetric)
            return pncov_create(&s->rows[i], s->devices[i], kind, dtype, d, lengthscales, n_lengthscales, nu, alpha, x0, n0, nullptr, 0, 0);
        return pncov_create(&s->rows[i], s->devices[i], kind, dtype, d, lengthscales, n_lengthscales, nu, alpha, block,
                            s->hi[i] - s->lo[i], cols, ncols, 0);
    });
    if (rc == 0 && s->symmetric && n_dev > 1 && dtype == PNCOV_F64 && kind <= PNCOV_RATQUAD && n0 >= 2048 && d <= 96)
        rc = for_each_device(n_dev, [&](int i) {
            return pncov_create(&s->whole[i], s->devices[i], kind, dtype, d, lengthscales, n_lengthscales, nu, alpha, x0, n0, nullptr, 0, 0);
        });
    if (rc) {
        const std::string keep = t_err;
        pncov_sharded_destroy(s);
        return fail(rc, keep);
    }
    *out = s;
    return 0;
}

PNCOV_API int pncov_sharded_gram(pncov_sharded* s, void* out, int64_t ld) {
    if (!s) return fail(PNCOV_EINVAL, "handle is NULL");
    if (s->n0 == 0 || s->n1 == 0) return 0;
    if (!out || ld < s->n1) return fail(PNCOV_EINVAL, "bad output buffer / leading dimension");
    return for_each_device(s->n_dev, [&](int i) {
        if (s->hi[i] == s->lo[i]) return 0;
        int gdt = PNCOV_F64;
        pncov_gram_dtype(s->rows[i], &gdt);
        const size_t ge = gdt == PNCOV_F32 ? 4 : 8;
        return pncov_gram(s->rows[i], static_cast<char*>(out) + (size_t)s->lo[i] * ld * ge, ld, 0);
    });
}

PNCOV_API int pncov_sharded_matmul(pncov_sharded* s, const void* v, int64_t k, int64_t ldv, void* y, int64_t ldy) {
    if (!s) return fail(PNCOV_EINVAL, "handle is NULL");
    if (k < 0 || ldv < k || ldy < k) return fail(PNCOV_EINVAL, "bad column count / leading dimension");
    if (k == 0 || s->n0 == 0) return 0;
    if (!v || !y) return fail(PNCOV_EINVAL, "NULL operand");
    const int64_t n = s->n0;
    // symmetric operator, one right-hand side: every device takes the row blocks I = i (mod n_dev) of the triangle into
    // exact integer planes, the host adds the planes (integer addition: the result is the bits one device would give)
    if (s->whole[0] && k == 1 && s->whole_state >= 0) {
        std::vector<std::vector<long long>> planes(s->n_dev, std::vector<long long>((size_t)3 * n));
        int rc = for_each_device(s->n_dev, [&](int i) {
            const int dev = s->devices[i];
            int r;
            if (!s->v_dev[i] && (r = pncov_device_alloc(dev, &s->v_dev[i], n * 8))) return r;
            if (!s->bins_dev[i] && (r = pncov_device_alloc(dev, &s->bins_dev[i], n * 24))) return r;
            DeviceGuard guard(dev);
            cudaStream_t st = s->whole[i]->stream;
            PN_CUDA(cudaMemcpy2DAsync(s->v_dev[i], 8, v, (size_t)ldv * 8, 8, (size_t)n, cudaMemcpyHostToDevice, st));
            if ((r = pncov_matmul_sym_part_bins_device(s->whole[i], s->v_dev[i], s->bins_dev[i], i, s->n_dev, st))) return r;
            PN_CUDA(cudaMemcpyAsync(planes[i].data(), s->bins_dev[i], (size_t)n * 24, cudaMemcpyDeviceToHost, st));
            PN_CUDA(cudaStreamSynchronize(st));
            return 0;
        });
        if (rc == PNCOV_ENOTIMPL && s->whole_state == 0) {
            s->whole_state = -1;  // the symmetric kernel does not apply to this data: row-sharded products below
        } else if (rc) {
            return rc;
        } else {
            s->whole_state = 1;
            std::vector<long long>& acc = planes[0];
            for (int i = 1; i < s->n_dev; ++i) {
                for (int64_t j = 0; j < 2 * n; ++j) acc[j] += planes[i][j];
                double* a = reinterpret_cast<double*>(acc.data() + 2 * n);
                const double* b = reinterpret_cast<const double*>(planes[i].data() + 2 * n);
                for (int64_t j = 0; j < n; ++j) a[j] += b[j];
            }
            const int dev = s->devices[0];
            if (!s->y_dev && (rc = pncov_device_alloc(dev, &s->y_dev, n * 8))) return rc;
            DeviceGuard guard(dev);
            cudaStream_t st = s->whole[0]->stream;
            PN_CUDA(cudaMemcpyAsync(s->bins_dev[0], acc.data(), (size_t)n * 24, cudaMemcpyHostToDevice, st));
            if ((rc = pncov_matmul_sym_finish_device(s->whole[0], s->bins_dev[0], s->y_dev, st))) return rc;
            PN_CUDA(cudaMemcpy2DAsync(y, (size_t)ldy * 8, s->y_dev, 8, 8, (size_t)n, cudaMemcpyDeviceToHost, st));
            PN_CUDA(cudaStreamSynchronize(st));
            return 0;
        }
    }
    return for_each_device(s->n_dev, [&](int i) {
        if (s->hi[i] == s->lo[i]) return 0;
        return pncov_matmul(s->rows[i], v, k, ldv, static_cast<char*>(y) + (size_t)s->lo[i] * ldy * s->es, ldy, 0);
    });
}

}  // extern "C"

cudaError_t pn_launch_pairs(const PnPairsLaunch& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((p.b + 255) / 256);
    if (!p.in_f32)
        pairs_kernel<double, double><<<blocks, 256, 0, stream>>>(p);
    else if (p.out_f32)
        pairs_kernel<float, float><<<blocks, 256, 0, stream>>>(p);
    else
        pairs_kernel<float, double><<<blocks, 256, 0, stream>>>(p);
    ++g_pn_launches;
    return cudaGetLastError();
}
