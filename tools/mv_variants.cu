// Design-space probe for the matrix-free ExpQuad product (d = 8, one right-hand side):
// times variants of the norm-identity kernel (threads per CTA, rows per thread, tile size,
// exp2 table style, split dot product) and the direct-difference baseline on synthetic data.
// Results steer probnum_b200/csrc/pncov_matvec.cu; see profiles/README.md.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iprobnum_b200/csrc \
//        -o tools/build/mv_variants tools/mv_variants.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#include "pncov_common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int D = 8;
constexpr int LD = 12;

template <int TAB>
__device__ __forceinline__ double epi(double S, const PnEpiParams& P, const double* tab, PnFlags& f) {
    return PnEpilogue<PN_EPI_EXPQUAD, TAB>::fast(S, P, tab, nullptr, f);
}

// exp2 with all constants forced into registers (tests whether constant-bank operands slow DFMA)
struct RegCoef { double cu, e1, e2, e3, magic; };
__device__ __forceinline__ double epi_reg(double S, const RegCoef& C, const double* tab) {
    double t = pn_fma(S, C.cu, C.magic);
    int n = pn_lo(t);
    double r = pn_fma(S, C.cu, C.magic - t);
    double q = pn_fma(C.e3, r, C.e2);
    q = pn_fma(q, r, C.e1);
    double T = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tab) + ((n << 3) & ((PN_EXP2_N - 1) << 3)));
    double res = pn_fma(T * r, q, T);
    return pn_make(pn_hi(res) + ((n << (20 - PNCOV_EXP2_BITS)) & 0xfff00000), pn_lo(res));
}

// MODE 0: direct differences on raw points (ld = 8); MODE 1: identity, serial dot;
// MODE 2: identity, dot split into two partial sums
template <int THREADS, int R, int TJ, int TAB, int MODE, int C>
__global__ void __launch_bounds__(THREADS)
mv_kernel(const double* __restrict__ rows, const double* __restrict__ cols, const double* __restrict__ v,
          double* __restrict__ y, int n0, int n1, const __grid_constant__ PnEpiParams P) {
    constexpr int XLD = MODE == 0 ? D : LD;
    constexpr int TABD = TAB ? PN_EXP2S_DOUBLES : PN_EXP2_N;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* stage0 = tabE + TABD;
    constexpr int STAGE = TJ * (XLD + 1);
    const int tid = threadIdx.x;
    const double* tab_epi = TAB ? tabE + (tid & 15) : tabE;
    constexpr uint32_t XB = TJ * XLD * 8, VB = TJ * 8;
    const int n_jt = n1 / TJ;
    if (tid == 0) { pn_mbar_init(&bars[0], 1); pn_mbar_init(&bars[1], 1); pn_mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        pn_mbar_expect_tx(&bars[0], XB + VB);
        pn_bulk_g2s(stage0, cols, XB, &bars[0]);
        pn_bulk_g2s(stage0 + TJ * XLD, v, VB, &bars[0]);
    }
    if (TAB) { for (int i = tid; i < PN_EXP2S_DOUBLES; i += THREADS) tabE[i] = pn_exp2s_table_g[i >> 4]; }
    else { for (int i = tid; i < PN_EXP2_N; i += THREADS) tabE[i] = pn_exp2_table_g[i]; }
    const int64_t row0 = (int64_t)blockIdx.x * (THREADS * R) + tid;
    double xr[R][D], nxr[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int m = 0; m < D; ++m) xr[r][m] = rows[(row0 + (int64_t)r * THREADS) * XLD + m];
        nxr[r] = MODE ? rows[(row0 + (int64_t)r * THREADS) * XLD + 8] : 0.0;
        acc[r] = 0.0;
    }
    __syncthreads();
    for (int jt = 0; jt < n_jt; ++jt) {
        const int buf = jt & 1;
        if (tid == 0 && jt + 1 < n_jt) {
            double* dst = stage0 + (buf ^ 1) * STAGE;
            pn_mbar_expect_tx(&bars[buf ^ 1], XB + VB);
            pn_bulk_g2s(dst, cols + (size_t)(jt + 1) * TJ * XLD, XB, &bars[buf ^ 1]);
            pn_bulk_g2s(dst + TJ * XLD, v + (size_t)(jt + 1) * TJ, VB, &bars[buf ^ 1]);
        }
        pn_mbar_wait(&bars[buf], (uint32_t)((jt >> 1) & 1));
        const double* xt = stage0 + buf * STAGE;
        const double* vt = xt + TJ * XLD;
#pragma unroll 1
        for (int jj = 0; jj < TJ; jj += C) {
            double xb[C][D], nyb[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int m = 0; m < D; m += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(xt + (jj + c) * XLD + m);
                    xb[c][m] = t.x; xb[c][m + 1] = t.y;
                }
                nyb[c] = MODE ? xt[(jj + c) * XLD + 8] : 0.0;
            }
            double S[R * C];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    double s;
                    if constexpr (MODE == 0) {
                        s = 0.0;
#pragma unroll
                        for (int m = 0; m < D; ++m) { double df = xr[r][m] - xb[c][m]; s = pn_fma(df, df, s); }
                    } else if constexpr (MODE == 1) {
                        s = nxr[r] + nyb[c];
#pragma unroll
                        for (int m = 0; m < D; ++m) s = pn_fma(xr[r][m], xb[c][m], s);
                    } else {
                        double s0 = nxr[r], s1 = nyb[c];
#pragma unroll
                        for (int m = 0; m < D; m += 2) { s0 = pn_fma(xr[r][m], xb[c][m], s0); s1 = pn_fma(xr[r][m + 1], xb[c][m + 1], s1); }
                        s = s0 + s1;
                    }
                    S[r * C + c] = s;
                }
            PnFlags f; f.reset();
            double K[R * C];
#pragma unroll
            for (int i = 0; i < R * C; ++i) K[i] = epi<TAB>(S[i], P, tab_epi, f);
            if (f.flagged(P)) {
#pragma unroll
                for (int i = 0; i < R * C; ++i) K[i] = pn_epilogue(S[i], P, pn_exp2_table_g, pn_log2_table_g);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double vk = vt[jj + c];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = pn_fma(K[r * C + c], vk, acc[r]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t row = row0 + (int64_t)r * THREADS;
        if (row < n0) y[row] = acc[r];
    }
}

static PnEpiParams make_params(bool identity) {
    PnEpiParams P{};
    pn_fill_math_constants(P);
    P.epi = PN_EPI_EXPQUAD;
    const double log2e = 1.4426950408889634;
    P.c_u = identity ? -log2e : -log2e * 0.5;
    P.c_s = 1.0; P.c_su = P.c_u;
    double smax = 1021.0 / fabs(P.c_u);
    unsigned long long b; memcpy(&b, &smax, 8);
    P.s_hi_word = (unsigned)(b >> 32); P.s_lo_word = 0;
    return P;
}

template <int THREADS, int R, int TJ, int TAB, int MODE, int C>
void run(const char* name, const double* raw, const double* a12, const double* b12, const double* v, double* y,
         int n, double* ref_out) {
    auto kern = mv_kernel<THREADS, R, TJ, TAB, MODE, C>;
    constexpr int XLD = MODE == 0 ? D : LD;
    const size_t smem = 128 + 8 * (TAB ? PN_EXP2S_DOUBLES : PN_EXP2_N) + 2 * 8 * TJ * (XLD + 1);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    PnEpiParams P = make_params(MODE != 0);
    const double* rows = MODE == 0 ? raw : a12;
    const double* cols = MODE == 0 ? raw : b12;
    dim3 grid(n / (THREADS * R));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<grid, THREADS, smem>>>(rows, cols, v, y, n, n, P);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<grid, THREADS, smem>>>(rows, cols, v, y, n, n, P);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    std::vector<double> h(64);
    CK(cudaMemcpy(h.data(), y, 64 * 8, cudaMemcpyDeviceToHost));
    double maxrel = 0;
    if (ref_out[0] == 0.0 && MODE == 0) for (int i = 0; i < 64; ++i) ref_out[i] = h[i];
    for (int i = 0; i < 64; ++i) maxrel = fmax(maxrel, fabs(h[i] - ref_out[i]) / fabs(ref_out[i]));
    printf("%-44s thr=%3d R=%d TJ=%3d TAB=%d C=%d occ=%d warps/SM=%2d  %8.3f ms  %7.1f Gpair/s  maxrel=%.1e\n", name, THREADS, R,
           TJ, TAB, C, occ, occ * THREADS / 32, best, (double)n * n / best / 1e6, maxrel);
}


// ---- DMMA variants --------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// THREADS per CTA, RG row groups of 8 per warp, TJ columns per stage, NS column-steps of 8 per
// inner iteration, SPLITK: independent accumulators for the two k-steps
template <int THREADS, int RG, int TJ, int NS, int SPLITK>
__global__ void __launch_bounds__(THREADS)
mma_kernel(const double* __restrict__ a12, const double* __restrict__ b12, const double* __restrict__ v,
           double* __restrict__ y, int n0, int n1, const __grid_constant__ PnEpiParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* stage0 = tabE + PN_EXP2_N;
    constexpr int STAGE = TJ * (LD + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    constexpr uint32_t XB = TJ * LD * 8, VB = TJ * 8;
    const int n_jt = n1 / TJ;
    if (tid == 0) { pn_mbar_init(&bars[0], 1); pn_mbar_init(&bars[1], 1); pn_mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        pn_mbar_expect_tx(&bars[0], XB + VB);
        pn_bulk_g2s(stage0, b12, XB, &bars[0]);
        pn_bulk_g2s(stage0 + TJ * LD, v, VB, &bars[0]);
    }
    for (int i = tid; i < PN_EXP2_N; i += THREADS) tabE[i] = pn_exp2_table_g[i];
    const int64_t row0 = (int64_t)blockIdx.x * (THREADS / 32 * RG * 8) + warp * (RG * 8);
    RegCoef rc = {P.c_u, P.e1, P.e2, P.e3, PN_EXP2_MAGIC};
    asm volatile("" : "+d"(rc.cu), "+d"(rc.e1), "+d"(rc.e2), "+d"(rc.e3), "+d"(rc.magic));
    double af[RG][2], nx[RG], acc[RG];
#pragma unroll
    for (int rg = 0; rg < RG; ++rg) {
        const double* pr = a12 + (row0 + rg * 8 + g) * LD;
        af[rg][0] = pr[q]; af[rg][1] = pr[4 + q]; nx[rg] = pr[8]; acc[rg] = 0.0;
    }
    __syncthreads();
    for (int jt = 0; jt < n_jt; ++jt) {
        const int buf = jt & 1;
        if (tid == 0 && jt + 1 < n_jt) {
            double* dst = stage0 + (buf ^ 1) * STAGE;
            pn_mbar_expect_tx(&bars[buf ^ 1], XB + VB);
            pn_bulk_g2s(dst, b12 + (size_t)(jt + 1) * TJ * LD, XB, &bars[buf ^ 1]);
            pn_bulk_g2s(dst + TJ * LD, v + (size_t)(jt + 1) * TJ, VB, &bars[buf ^ 1]);
        }
        pn_mbar_wait(&bars[buf], (uint32_t)((jt >> 1) & 1));
        const double* bt = stage0 + buf * STAGE;
        const double* vt = bt + TJ * LD;
#pragma unroll 1
        for (int jj = 0; jj < TJ; jj += 8 * NS) {
            double S[NS][2 * RG];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = jj + 8 * s;
                const double b0 = bt[(j + g) * LD + q], b1 = bt[(j + g) * LD + 4 + q];
                const double ny0 = bt[(j + 2 * q) * LD + 8], ny1 = bt[(j + 2 * q + 1) * LD + 8];
#pragma unroll
                for (int rg = 0; rg < RG; ++rg) {
                    double c0 = nx[rg] + ny0, c1 = nx[rg] + ny1;
                    if (SPLITK == 1) {
                        double d0 = 0.0, d1 = 0.0;
                        dmma884(c0, c1, af[rg][0], b0);
                        dmma884(d0, d1, af[rg][1], b1);
                        c0 += d0; c1 += d1;
                    } else {
                        dmma884(c0, c1, af[rg][0], b0);
                        dmma884(c0, c1, af[rg][1], b1);
                    }
                    S[s][2 * rg] = c0; S[s][2 * rg + 1] = c1;
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = jj + 8 * s;
                PnFlags f; f.reset();
                double K[2 * RG];
                if (SPLITK == 2) {
#pragma unroll
                    for (int i = 0; i < 2 * RG; ++i) K[i] = epi_reg(S[s][i], rc, tabE);
                } else {
#pragma unroll
                for (int i = 0; i < 2 * RG; ++i) K[i] = epi<0>(S[s][i], P, tabE, f);
                }
                if (SPLITK != 2 && f.flagged(P)) {
#pragma unroll
                    for (int i = 0; i < 2 * RG; ++i) K[i] = pn_epilogue(S[s][i], P, pn_exp2_table_g, pn_log2_table_g);
                }
                const double2 vv = *reinterpret_cast<const double2*>(vt + j + 2 * q);
#pragma unroll
                for (int rg = 0; rg < RG; ++rg) {
                    acc[rg] = pn_fma(K[2 * rg], vv.x, acc[rg]);
                    acc[rg] = pn_fma(K[2 * rg + 1], vv.y, acc[rg]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int rg = 0; rg < RG; ++rg) {
        double s = acc[rg];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const int64_t row = row0 + rg * 8 + g;
        if (q == 0 && row < n0) y[row] = s;
    }
}

template <int THREADS, int RG, int TJ, int NS, int SPLITK>
void run_mma(const char* name, const double* a12, const double* b12, const double* v, double* y, int n, double* ref) {
    auto kern = mma_kernel<THREADS, RG, TJ, NS, SPLITK>;
    const size_t smem = 128 + 8 * PN_EXP2_N + 2 * 8 * TJ * (LD + 1);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    PnEpiParams P = make_params(true);
    dim3 grid(n / (THREADS / 32 * RG * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<grid, THREADS, smem>>>(a12, b12, v, y, n, n, P);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<grid, THREADS, smem>>>(a12, b12, v, y, n, n, P);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    std::vector<double> h(64);
    CK(cudaMemcpy(h.data(), y, 64 * 8, cudaMemcpyDeviceToHost));
    double maxrel = 0;
    for (int i = 0; i < 64; ++i) maxrel = fmax(maxrel, fabs(h[i] - ref[i]) / fabs(ref[i]));
    printf("%-44s thr=%3d RG=%d TJ=%3d NS=%d SPLITK=%d occ=%d warps/SM=%2d  %8.3f ms  %7.1f Gpair/s  maxrel=%.1e\n", name,
           THREADS, RG, TJ, NS, SPLITK, occ, occ * THREADS / 32, best, (double)n * n / best / 1e6, maxrel);
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 131072;
    std::vector<double> raw((size_t)n * D), a12((size_t)n * LD, 0.0), b12((size_t)n * LD, 0.0), v(n);
    srand(1);
    auto rnd = []() { double u = 0; for (int i = 0; i < 12; ++i) u += rand() / (double)RAND_MAX; return u - 6.0; };
    const double s = sqrt(0.5);
    for (int i = 0; i < n; ++i) {
        double nrm = 0;
        for (int m = 0; m < D; ++m) {
            double x = rnd(); raw[(size_t)i * D + m] = x;
            a12[(size_t)i * LD + m] = x * s; b12[(size_t)i * LD + m] = -2.0 * x * s; nrm += x * s * x * s;
        }
        a12[(size_t)i * LD + 8] = nrm; b12[(size_t)i * LD + 8] = nrm;
        v[i] = rnd();
    }
    double *draw, *da, *db, *dv, *dy;
    CK(cudaMalloc(&draw, raw.size() * 8)); CK(cudaMalloc(&da, a12.size() * 8)); CK(cudaMalloc(&db, b12.size() * 8));
    CK(cudaMalloc(&dv, n * 8)); CK(cudaMalloc(&dy, n * 8));
    CK(cudaMemcpy(draw, raw.data(), raw.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(da, a12.data(), a12.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b12.data(), b12.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dv, v.data(), n * 8, cudaMemcpyHostToDevice));
    double ref[64] = {0};
    const bool dfma = argc > 2;
    run<128, 2, 128, 0, 0, 2>("direct (shipping kernel shape)", draw, da, db, dv, dy, n, ref);
    if (dfma) {
        run<128, 2, 64, 0, 1, 2>("identity serial dot TJ=64", draw, da, db, dv, dy, n, ref);
        run<128, 2, 64, 0, 2, 2>("identity split dot TJ=64", draw, da, db, dv, dy, n, ref);
        run<128, 4, 64, 0, 1, 2>("identity R=4", draw, da, db, dv, dy, n, ref);
    }
    run_mma<128, 4, 128, 2, 0>("dmma NS=2 (shipping)", da, db, dv, dy, n, ref);
    run_mma<128, 4, 128, 2, 2>("dmma NS=2, constants in registers, no flags", da, db, dv, dy, n, ref);
    run_mma<128, 4, 128, 4, 0>("dmma NS=4", da, db, dv, dy, n, ref);
    run_mma<128, 4, 128, 4, 2>("dmma NS=4, register constants", da, db, dv, dy, n, ref);
    run_mma<128, 2, 128, 4, 2>("dmma RG=2 NS=4, register constants", da, db, dv, dy, n, ref);
    run_mma<256, 2, 128, 4, 2>("dmma RG=2 NS=4 256thr, register constants", da, db, dv, dy, n, ref);
    return 0;
}
