// FP64 Cholesky factorisation and solves on the device.
//
// Replaces scipy.linalg.cho_factor / cho_solve in the Bayesian-quadrature belief update
// (/root/reference/src/probnum/quad/solvers/belief_updates/_belief_update.py:61-103:
// compute_gram_cho_factor adds `jitter` to the diagonal and factors, gram_cho_solve solves with
// the factor).  At 16384 nodes the host dpotrf (1.47 TFLOP) takes ~15 s and dominates
// bayesquad_from_data once the Gram matrix itself is built on the GPU.
//
// Right-looking blocked algorithm on a row-major n x n matrix (lower triangle), block size 128:
//   for every block column k
//     chol_diag_kernel   one CTA: unblocked Cholesky of the 128 x 128 diagonal block in shared
//                        memory (one barrier per column) and the explicit inverse of its factor
//     chol_gemm_kernel   mode 0 ("TRSM as GEMM"): L[i,k] = A[i,k] inv(L_kk)^T for all i > k, into a
//                        contiguous panel that is then copied over A[i,k]
//                        mode 1: A[i,j] -= L[i,k] L[j,k]^T for all i >= j > k  (n^3/3 of the flops)
// Both GEMM modes are C = A B^T with K-contiguous operands, i.e. exactly the "row.col" form of
// mma.sync.m8n8k4.f64 (DMMA, the only FP64 tensor path of sm_100a): a CTA of 4 warps owns a
// 128 x 64 tile, a warp 32 x 64 of it (64 accumulators per lane; three CTAs per SM cover each
// other's prologue and read-modify-write epilogue), operands stream through a
// two-stage cp.async pipeline in K-chunks of 16 with a row stride of 20 doubles, which makes the
// fragment loads bank-conflict free.
// Solves use the stored inverse diagonal blocks: every substitution step is a 128 x 128 product
// plus one pass over a block row / block column of L.
#include <cstdlib>

#include "pncov_internal.h"

namespace {

constexpr int CB = 128;          // block size
constexpr int CK = 16;           // K chunk of the GEMM pipeline
constexpr int CLD = CK + 4;      // shared-memory row stride in doubles (= 4 mod 16: conflict-free)
constexpr int CG_THREADS = 128;
constexpr int CTN = 64;          // columns of a GEMM tile (rows: CB)
constexpr int CD_LD = CB + 1;    // diagonal kernel: padded row stride

__device__ __forceinline__ void ch_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ void ch_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pn_smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void ch_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ch_cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// phase time stamps of chol_diag_kernel for tools/chol_diag_probe.cu (compiled out of the library)
#ifdef PN_DIAG_CLOCKS
__device__ long long g_diag_clk[16];
#define PN_DIAG_STAMP(k)                              \
    do {                                              \
        __syncthreads();                              \
        if (threadIdx.x == 0) g_diag_clk[k] = clock64(); \
    } while (0)
#else
#define PN_DIAG_STAMP(k)
#endif

struct ChGemm {
    const double* A;  // mode 0: first row block of the panel inside the matrix; mode 1: contiguous panel
    int64_t lda;
    const double* B;  // mode 0: inverse of the diagonal factor (128 x 128); mode 1: contiguous panel
    int64_t ldb;
    double* C;        // mode 0: unused (the panel is copied over A afterwards); mode 1: block (0, 0) of the trailing matrix
    int64_t ldc;
    double* panel;    // mode 0: contiguous copy of the new block column (ld = 128)
    double* Lcol;     // mode 1: block (0, -1) of the trailing matrix' frame, i.e. where the panel belongs in the matrix (ld = ldc)
    int mode;
    int nb;           // mode 1: block rows of the trailing matrix (its block (0, 0) belongs to the diagonal kernel)
    // hand-over between the update of step k and the panel of step k + 1, which run concurrently on two streams:
    // the update's tiles of its first block column come first in its grid and count themselves into rowcount[global
    // block row] when stored; a panel CTA waits until its row has `target` of them
    int* rowcount;    // [npad / 128], zero before the factorisation; nullptr: no hand-over
    int row0;         // global block row of local block row 0
    int target;       // mode 0: completed first-column half tiles this row must have seen (0: do not wait)
    int* err;         // set to 1 if a wait gave up (never in a correct schedule; keeps a broken one from hanging)
};

__global__ void __launch_bounds__(CG_THREADS, 3) chol_gemm_kernel(const __grid_constant__ ChGemm p) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* As = reinterpret_cast<double*>(smem);  // [2][CB * CLD]
    double* Bs = As + 2 * CB * CLD;                // [2][CTN * CLD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp;  // 4 warps: 32 rows x 64 columns each

    // a CTA owns 128 rows x 64 columns: block (bi, bj), column half h
    int64_t bi, bj;
    const int64_t t2 = (int64_t)blockIdx.x;
    const int h = (int)(t2 & 1);
    if (p.mode == 0) {
        bi = t2 >> 1;
        bj = 0;
        if (p.target > 0) {  // this row of the block column must have received the update of the previous step
            if (tid == 0) {
                const int* cnt = p.rowcount + p.row0 + bi;
                long long spins = 0;
                for (;;) {
                    int seen;
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                    if (seen >= p.target) break;
                    if (++spins > (1ll << 22)) {  // seconds (a legitimate wait is microseconds): report instead of hanging
                        atomicExch(p.err, -1);
                        break;
                    }
                }
            }
            __syncthreads();
        }
    } else {
        // the nb - 1 tiles of the first block column first (the next panel waits for them), then the lower triangle
        // of the remaining blocks row by row; block (0, 0) is not in the grid
        const int64_t t = t2 >> 1;
        const int64_t ntiles = (int64_t)p.nb * (p.nb + 1) / 2 - 1;
        if (t >= ntiles) {
            // the last 2 nb CTAs of the grid put the panel into the matrix (rows below the diagonal block of its
            // column; nobody else touches them): they fill the slots of the update's last wave
            const int64_t rb = t - ntiles;
            const double2* src = reinterpret_cast<const double2*>(p.A + rb * CB * CB + h * CTN);
            double* dst = p.Lcol + rb * CB * p.ldc + h * CTN;
#pragma unroll 4
            for (int idx = tid; idx < CB * (CTN / 2); idx += CG_THREADS) {
                const int row = idx / (CTN / 2), c2 = idx % (CTN / 2);
                *reinterpret_cast<double2*>(dst + (int64_t)row * p.ldc + 2 * c2) = src[row * (CB / 2) + c2];
            }
            return;
        }
        if (t < p.nb - 1) {
            bi = t + 1;
            bj = 0;
        } else {
            const int64_t u = t - (p.nb - 1);
            int64_t r = (int64_t)((sqrt(8.0 * (double)u + 1.0) - 1.0) * 0.5);
            while ((r + 1) * (r + 2) / 2 <= u) ++r;
            while (r * (r + 1) / 2 > u) --r;
            bi = r + 1;
            bj = u - r * (r + 1) / 2 + 1;
        }
    }
    const double* At = p.A + bi * CB * p.lda;
    const double* Bt = (p.mode == 0 ? p.B : p.B + bj * CB * p.ldb) + (int64_t)h * CTN * p.ldb;
    double* Ct = p.C + bi * CB * p.ldc + (p.mode == 0 ? 0 : bj * CB) + h * CTN;

    auto load_chunk = [&](int kc, int buf) {
        double* as = As + buf * CB * CLD;
        double* bs = Bs + buf * CTN * CLD;
        constexpr int PIECES = CK / 2;  // 16-byte pieces per row
#pragma unroll
        for (int i = 0; i < CB * PIECES / CG_THREADS; ++i) {
            const int idx = tid + i * CG_THREADS;
            const int row = idx / PIECES, c2 = (idx % PIECES) * 2;
            ch_cp_async16(as + row * CLD + c2, At + (int64_t)row * p.lda + kc * CK + c2);
        }
#pragma unroll
        for (int i = 0; i < CTN * PIECES / CG_THREADS; ++i) {
            const int idx = tid + i * CG_THREADS;
            const int row = idx / PIECES, c2 = (idx % PIECES) * 2;
            ch_cp_async16(bs + row * CLD + c2, Bt + (int64_t)row * p.ldb + kc * CK + c2);
        }
        ch_cp_commit();
    };

    double acc[4][8][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;

    load_chunk(0, 0);
    constexpr int NCHUNK = CB / CK;
    for (int kc = 0; kc < NCHUNK; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < NCHUNK) {
            load_chunk(kc + 1, buf ^ 1);
            ch_cp_wait<1>();
        } else {
            ch_cp_wait<0>();
        }
        __syncthreads();
        const double* as = As + buf * CB * CLD + (wm * 32 + g) * CLD + q;
        const double* bs = Bs + buf * CTN * CLD + g * CLD + q;
#pragma unroll
        for (int k4 = 0; k4 < CK / 4; ++k4) {
            double a[4], b[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = as[r * 8 * CLD + k4 * 4];
#pragma unroll
            for (int c = 0; c < 8; ++c) b[c] = bs[c * 8 * CLD + k4 * 4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) ch_dmma(acc[r][c][0], acc[r][c][1], a[r], b[c]);
        }
        __syncthreads();  // everyone is done with `buf` before chunk kc + 2 lands in it
    }

#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = wm * 32 + r * 8 + g;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int col = c * 8 + 2 * q;
            double2* dst = reinterpret_cast<double2*>(Ct + (int64_t)row * p.ldc + col);
            if (p.mode == 0) {
                // panel only: the two column halves of a row block are different CTAs and each reads
                // all 128 columns of A, so A's own storage is filled by a copy after the kernel
                *reinterpret_cast<double2*>(p.panel + (bi * CB + row) * CB + h * CTN + col) =
                    make_double2(acc[r][c][0], acc[r][c][1]);
            } else {
                double2 v = *dst;
                v.x -= acc[r][c][0];
                v.y -= acc[r][c][1];
                *dst = v;
            }
        }
    }
    if (p.mode == 1 && bj == 0 && p.rowcount != nullptr) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(p.rowcount + p.row0 + bi, 1);
        }
    }
}

// 1 / sqrt(a) for a pivot a > 0: hardware seed (22 bits) and two Newton steps, branch-free.  The library rsqrt() is on
// the critical chain of every elimination step and costs ~250 cycles there (special-case handling, a call); this one
// is seven dependent operations.  Relative error below 2 ulp; subnormal pivots are flushed to zero by the seed and end
// as a non-finite factor, which the next pivot test reports.
__device__ __forceinline__ double ch_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    double e = pn_fma(-(h * y), y, 0.5);
    y = pn_fma(y, e, y);
    e = pn_fma(-(h * y), y, 0.5);
    return pn_fma(y, e, y);
}

// 16 elimination steps j = 16 JJ .. 16 JJ + 15 of the in-register Cholesky of chol_diag_kernel, two columns per
// barrier.  Thread (ty, tx) owns a[ty + 16 ii][tx + 16 cc] in r[ii][cc] for the 16 x 16 groups on and below the
// diagonal (cc <= ii); column group JJ is a compile-time constant, so every register index is static and finished
// column groups and the groups above the diagonal cost nothing.  The owners publish columns j and j + 1 as they are
// (column j + 1 without the update from column j); every thread derives L[.][j], then column j + 1 updated and scaled,
// for its own rows and columns from the published values -- the same operations, in the same order, as two single
// steps, with one barrier and one shared-memory round trip (the step is latency-bound).  rd receives 1 / L[j][j].
template <int JJ>
__device__ __forceinline__ void chol_diag_steps(double (&r)[8][8], double* colbuf, double* rd, int ty, int tx, int tid,
                                                int* info, int pivot_base) {
#pragma unroll 1
    for (int jt = 0; jt < 16; jt += 2) {
        const int j = 16 * JJ + jt;
        double* cb0 = colbuf + (jt & 2) * CB;  // [2][CB] per pair, double-buffered over pairs: one barrier per pair
        double* cb1 = cb0 + CB;
        if (tx == jt) {
#pragma unroll
            for (int ii = JJ; ii < 8; ++ii) cb0[ty + 16 * ii] = r[ii][JJ];
        } else if (tx == jt + 1) {
#pragma unroll
            for (int ii = JJ; ii < 8; ++ii) cb1[ty + 16 * ii] = r[ii][JJ];
        }
        __syncthreads();
        double ajj = cb0[j];
        if (!(ajj > 0.0)) {  // not positive definite (or NaN): record the first offender, keep going
            if (tid == 0) atomicCAS(info, 0, pivot_base + j + 1);
            ajj = 1.0;
        }
        const double d0 = ch_rsqrt(ajj);
        const double l10 = cb0[j + 1] * d0;  // L[j+1][j]
        double a11 = pn_fma(-l10, l10, cb1[j + 1]);
        if (!(a11 > 0.0)) {
            if (tid == 0) atomicCAS(info, 0, pivot_base + j + 2);
            a11 = 1.0;
        }
        const double d1 = ch_rsqrt(a11);
        if (tid == 0) {
            rd[j] = d0;
            rd[j + 1] = d1;
        }
        double li0[8 - JJ], lc0[8 - JJ], li1[8 - JJ], lc1[8 - JJ];
#pragma unroll
        for (int ii = JJ; ii < 8; ++ii) {  // L[i][j], L[i][j+1] for my rows
            li0[ii - JJ] = cb0[ty + 16 * ii] * d0;
            li1[ii - JJ] = pn_fma(-li0[ii - JJ], l10, cb1[ty + 16 * ii]) * d1;
        }
#pragma unroll
        for (int cc = JJ; cc < 8; ++cc) {  // L[c][j], L[c][j+1] for my columns
            lc0[cc - JJ] = cb0[tx + 16 * cc] * d0;
            lc1[cc - JJ] = pn_fma(-lc0[cc - JJ], l10, cb1[tx + 16 * cc]) * d1;
        }
        // a[i][c] -= L[i][j] L[c][j], then -= L[i][j+1] L[c][j+1], for the columns right of j resp. j + 1; entries above
        // the diagonal inside a diagonal group hold dead values that never feed a live one, so they need no predicate
#pragma unroll
        for (int cc = JJ; cc < 8; ++cc) {
            const bool upd0 = (cc > JJ) || (tx > jt), upd1 = (cc > JJ) || (tx > jt + 1);
#pragma unroll
            for (int ii = cc; ii < 8; ++ii) {
                if (upd0) r[ii][cc] = pn_fma(-li0[ii - JJ], lc0[cc - JJ], r[ii][cc]);
                if (upd1) r[ii][cc] = pn_fma(-li1[ii - JJ], lc1[cc - JJ], r[ii][cc]);
            }
        }
        if (tx == jt) {  // column j is final: L[j][j] = a_jj / sqrt(a_jj), L[i][j] = a_ij / sqrt(a_jj)
#pragma unroll
            for (int ii = JJ; ii < 8; ++ii) r[ii][JJ] = li0[ii - JJ];
        } else if (tx == jt + 1) {
#pragma unroll
            for (int ii = JJ; ii < 8; ++ii) r[ii][JJ] = li1[ii - JJ];
        }
    }
}

// Off-diagonal blocks of the inverse at block distance DIST (32 x 32 blocks): X_ib = -X_ii T_ib with
// T_ib = sum_{m=b}^{i-1} L_im X_mb, for the 4 - DIST pairs (i, b = i - DIST) at once.  L[i][l] sits at a[l][i] (upper
// triangle), X[i][c] at a[i][c] (lower triangle); T: (4 - DIST) blocks with row stride 33.  A thread owns a TR x TC
// micro tile of one block (unit-stride operands along a row of `a` for both factors).
template <int DIST, int TR, int TC>
__device__ __forceinline__ void chol_inv_offdiag(double* __restrict__ a, double* __restrict__ T, int tid) {
    constexpr int PER = (32 / TR) * (32 / TC);  // threads per block
    constexpr int NBLK = 4 - DIST;
    static_assert(PER * NBLK <= 256, "micro tiles do not fit the CTA");
    const int blk = tid / PER, u = tid % PER;
    const int r0 = (u / (32 / TC)) * TR, q0 = (u % (32 / TC)) * TC;  // tile origin inside the block
    const int i = blk + DIST, b = blk;
    const bool on = blk < NBLK;
    if (on) {
        double acc[TR][TC];
#pragma unroll
        for (int x = 0; x < TR; ++x)
#pragma unroll
            for (int y = 0; y < TC; ++y) acc[x][y] = 0.0;
        const double* lrow = a + 32 * i + r0;  // + l * CD_LD: L[32 i + r0 ..][l]
        const double* xcol = a + 32 * b + q0;  // + l * CD_LD: X[l][32 b + q0 ..]
        // m = b: X_bb is lower triangular and its upper part is not stored (L^T lives there)
#pragma unroll 4
        for (int l = 32 * b; l < 32 * b + 32; ++l) {
            double lv[TR], xv[TC];
#pragma unroll
            for (int x = 0; x < TR; ++x) lv[x] = lrow[l * CD_LD + x];
#pragma unroll
            for (int y = 0; y < TC; ++y) xv[y] = (l >= 32 * b + q0 + y) ? xcol[l * CD_LD + y] : 0.0;
#pragma unroll
            for (int x = 0; x < TR; ++x)
#pragma unroll
                for (int y = 0; y < TC; ++y) acc[x][y] = pn_fma(lv[x], xv[y], acc[x][y]);
        }
#pragma unroll 4
        for (int l = 32 * b + 32; l < 32 * i; ++l) {
            double lv[TR], xv[TC];
#pragma unroll
            for (int x = 0; x < TR; ++x) lv[x] = lrow[l * CD_LD + x];
#pragma unroll
            for (int y = 0; y < TC; ++y) xv[y] = xcol[l * CD_LD + y];
#pragma unroll
            for (int x = 0; x < TR; ++x)
#pragma unroll
                for (int y = 0; y < TC; ++y) acc[x][y] = pn_fma(lv[x], xv[y], acc[x][y]);
        }
#pragma unroll
        for (int x = 0; x < TR; ++x)
#pragma unroll
            for (int y = 0; y < TC; ++y) T[(blk * 32 + r0 + x) * 33 + q0 + y] = acc[x][y];
    }
    __syncthreads();
    if (on) {  // X_ib = -X_ii T: X_ii[r][k] = a[32 i + r][32 i + k] for k <= r
        double acc[TR][TC];
#pragma unroll
        for (int x = 0; x < TR; ++x)
#pragma unroll
            for (int y = 0; y < TC; ++y) acc[x][y] = 0.0;
        const double* xi = a + (32 * i + r0) * CD_LD + 32 * i;
        const double* tc = T + blk * 32 * 33 + q0;
#pragma unroll 4
        for (int k = 0; k < r0 + TR; ++k) {
            double xv[TR], tv[TC];
#pragma unroll
            for (int x = 0; x < TR; ++x) xv[x] = (k <= r0 + x) ? xi[x * CD_LD + k] : 0.0;
#pragma unroll
            for (int y = 0; y < TC; ++y) tv[y] = tc[k * 33 + y];
#pragma unroll
            for (int x = 0; x < TR; ++x)
#pragma unroll
                for (int y = 0; y < TC; ++y) acc[x][y] = pn_fma(xv[x], tv[y], acc[x][y]);
        }
#pragma unroll
        for (int x = 0; x < TR; ++x)
#pragma unroll
            for (int y = 0; y < TC; ++y) a[(32 * i + r0 + x) * CD_LD + 32 * b + q0 + y] = -acc[x][y];
    }
    __syncthreads();
}

// One CTA: Cholesky of the 128 x 128 diagonal block (lower triangle of A, row-major) and the
// inverse of its factor.  The block lives in registers (8 x 8 per thread, cyclic 16 x 16
// distribution, groups on and below the diagonal); per elimination step the owners publish one column through
// shared memory, one barrier, and everybody updates its registers.  info receives the 1-based index of the first
// non-positive pivot.  Phase times (tools/chol_diag_probe.cu) are in DESIGN.md 4.5.
// P != nullptr (look-ahead): the block first receives its own trailing update A -= P P^T from the 128 x 128
// head of the current panel, so the kernel depends on the panel alone and can take an SM before the
// wide update of the same step fills the machine.
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ invL,
                                                         int* __restrict__ info, int pivot_base,
                                                         const double* __restrict__ P) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* a = reinterpret_cast<double*>(smem);  // [CB][CD_LD]: upper triangle L^T, lower triangle inv(L)
    double* rd = a + CB * CD_LD;                  // [CB] reciprocals of the diagonal of L
    double* colbuf = rd + CB;                     // [4][CB]
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    PN_DIAG_STAMP(0);
    double r[8][8];
    if (P != nullptr) {
        // the head of the panel goes to shared memory asynchronously (row stride CD_LD: 8-byte pieces) while the
        // block itself is loaded
        for (int idx = tid; idx < CB * CB; idx += 256)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(pn_smem_u32(a + (idx >> 7) * CD_LD + (idx & (CB - 1)))),
                         "l"(P + idx)
                         : "memory");
        ch_cp_commit();
    }
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
#pragma unroll
        for (int cc = 0; cc <= ii; ++cc) r[ii][cc] = A[(int64_t)(ty + 16 * ii) * lda + tx + 16 * cc];
    if (P != nullptr) {
        // own tile of the trailing update, A -= P P^T (groups on and below the diagonal)
        ch_cp_wait<0>();
        __syncthreads();
        PN_DIAG_STAMP(12);
        const double* pr = a + ty * CD_LD;
        const double* pc = a + tx * CD_LD;
#pragma unroll 2
        for (int l = 0; l < CB; ++l) {
            double vi[8], vc[8];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) vi[ii] = pr[16 * ii * CD_LD + l];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) vc[cc] = pc[16 * cc * CD_LD + l];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii)
#pragma unroll
                for (int cc = 0; cc <= ii; ++cc) r[ii][cc] = pn_fma(-vi[ii], vc[cc], r[ii][cc]);
        }
        // `a` is written again only after the elimination steps, whose barriers order it after these reads
    }
    PN_DIAG_STAMP(1);
    chol_diag_steps<0>(r, colbuf, rd, ty, tx, tid, info, pivot_base);
    chol_diag_steps<1>(r, colbuf, rd, ty, tx, tid, info, pivot_base);
    PN_DIAG_STAMP(2);
    chol_diag_steps<2>(r, colbuf, rd, ty, tx, tid, info, pivot_base);
    chol_diag_steps<3>(r, colbuf, rd, ty, tx, tid, info, pivot_base);
    PN_DIAG_STAMP(3);
    chol_diag_steps<4>(r, colbuf, rd, ty, tx, tid, info, pivot_base);
    chol_diag_steps<5>(r, colbuf, rd, ty, tx, tid, info, pivot_base);
    PN_DIAG_STAMP(4);
    chol_diag_steps<6>(r, colbuf, rd, ty, tx, tid, info, pivot_base);
    chol_diag_steps<7>(r, colbuf, rd, ty, tx, tid, info, pivot_base);
    PN_DIAG_STAMP(5);
    // L to global memory (lower triangle) and, transposed, to shared memory for the inverse
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
#pragma unroll
        for (int cc = 0; cc <= ii; ++cc) {
            const int i = ty + 16 * ii, c = tx + 16 * cc;
            if (c <= i) A[(int64_t)i * lda + c] = r[ii][cc];
            if (c < i) a[c * CD_LD + i] = r[ii][cc];
        }
    __syncthreads();
    PN_DIAG_STAMP(6);
    // inverse X of L, blocked 4 x 4 with 32 x 32 blocks.  L[i][l] sits at a[l][i] (upper triangle),
    // X[i][c] goes to a[i][c] (lower triangle, free now).
    // (A) diagonal blocks: a thread per column, the column in registers; once x_k is final it goes into the rows
    // below (independent FMAs; L[i][k] is the same address for the whole warp)
    if (tid < CB) {
        const int b0 = (tid >> 5) << 5, cl = tid & 31;
        double x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = (i == cl) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            x[k] *= rd[b0 + k];
#pragma unroll
            for (int i = k + 1; i < 32; ++i) x[i] = pn_fma(-a[(b0 + k) * CD_LD + b0 + i], x[k], x[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i >= cl) a[(b0 + i) * CD_LD + tid] = x[i];
    }
    __syncthreads();
    PN_DIAG_STAMP(7);
    // (B) off-diagonal blocks by block distance; the blocks of one distance are independent
    double* T = colbuf + 4 * CB;
    chol_inv_offdiag<1, 4, 4>(a, T, tid);
    PN_DIAG_STAMP(8);
    chol_inv_offdiag<2, 4, 2>(a, T, tid);
    PN_DIAG_STAMP(9);
    chol_inv_offdiag<3, 2, 2>(a, T, tid);
    PN_DIAG_STAMP(10);
    // X to global memory, row-major, zeros above the diagonal
    for (int idx = tid; idx < CB * CB; idx += 256) {
        const int i = idx / CB, c = idx % CB;
        invL[idx] = (c <= i) ? a[i * CD_LD + c] : 0.0;
    }
    PN_DIAG_STAMP(11);
}

// M[i][i] += jitter for i < n; padded diagonal (i >= n) = 1
__global__ void chol_prepare_diag_kernel(double* M, int64_t ld, int64_t n, int64_t npad, double jitter) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) M[i * ld + i] += jitter;
    else if (i < npad) M[i * ld + i] = 1.0;
}

// ---- solves -------------------------------------------------------------------------------------------
// One to three right-hand sides: ONE launch per direction.  The substitution is a chain of T = n / 128
// dependent block steps; as launches (one per step or per group of steps) the chain costs 50 us a link.  Here
// every block row (forward) / block column (backward) has a CTA of its own, CTAs hand the solved blocks to each
// other through flags in global memory, and the only thing left on the critical path of step i is
//     see flag i-1  ->  load x_{i-1} (512 B..3 KB from L2)  ->  one 128 x 128 product  ->  the product with
//     inv(L_ii)  ->  store x_i, release flag i.
// CTA i streams its blocks L[i, 0..i-1] (forward) or L[T-1..i+1, i] (backward) through a three-stage cp.async
// ring of half blocks (64 x 128 doubles) as fast as the flags allow: far behind the frontier that is a plain
// streaming pass over the row, which all rows do at the same time.  Partial sums stay in registers over the whole
// row; they are reduced once, before the product with the inverse diagonal block.
// Block indices are handed out by an atomic ticket in start order, so a CTA only ever waits for CTAs that are
// already running: no co-residency requirement, any T.
constexpr int SC_THREADS = 256;
constexpr int SC_UR = 64;               // rows of a pipeline unit (half a block)
constexpr int SC_STAGES = 3;
constexpr int SC_UNIT = SC_UR * CB;     // doubles per unit
static_assert(sizeof(double) * (SC_STAGES * SC_UNIT + 4 * 3 * CB) <= 227 * 1024, "chain kernel: ring + operands exceed the shared memory of an SM");

struct ChChain {
    const double* M;     // factor, row-major, lower triangle
    int64_t ld;
    const double* invd;  // T inverse diagonal blocks
    double* b;           // NR adjacent columns of a row-major (T * 128, ldb) array: right-hand sides in, solutions out
    int64_t ldb;
    int* ticket;         // zero before the launch
    int* flags;          // [T], zero before the launch; flags[j] = 1: block j of the solution is in b
    int T;
};

__device__ __forceinline__ int sc_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sc_st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// forward: sums over the 32 lanes of 16 values per lane in 16 shuffles (halving exchange); on return v[0] of
// lane L is the total of value ((L >> 1) & 15) bit-reversed as: bit 4 of L -> 8, bit 3 -> 4, bit 2 -> 2, bit 1 -> 1
__device__ __forceinline__ double sc_reduce16(double (&v)[16], int lane) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool up = lane & 16;
        const double send = up ? v[k] : v[k + 8], keep = up ? v[k + 8] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool up = lane & 8;
        const double send = up ? v[k] : v[k + 4], keep = up ? v[k + 4] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const bool up = lane & 4;
        const double send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const bool up = lane & 2;
        const double send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// forward unit H of a block: acc[8 H + q][r] += sum_c S[warp + 8 q][c] v_r[c] over this lane's 4 columns
template <int NR, int H>
__device__ __forceinline__ void sc_apply_rows(const double* __restrict__ S, const double* __restrict__ v,
                                              double (&acc)[NR][16], int warp, int lane) {
    double4 v4[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) v4[r] = *reinterpret_cast<const double4*>(v + r * CB + 4 * lane);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const double4 m = *reinterpret_cast<const double4*>(S + (warp + 8 * q) * CB + 4 * lane);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double a = acc[r][8 * H + q];
            a = pn_fma(m.x, v4[r].x, a);
            a = pn_fma(m.y, v4[r].y, a);
            a = pn_fma(m.z, v4[r].z, a);
            a = pn_fma(m.w, v4[r].w, a);
            acc[r][8 * H + q] = a;
        }
    }
}

// backward unit h of a block: acc[r] += sum_l S[l][col] v_r[64 h + l] over this thread's 32 rows of the unit
template <int NR>
__device__ __forceinline__ void sc_apply_cols(const double* __restrict__ S, const double* __restrict__ v, int h,
                                              double (&acc)[NR], int col, int half) {
    const double* sp = S + (half * 32) * CB + col;
    const double* vp = v + 64 * h + half * 32;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) {
        const double m = sp[l * CB];
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r] = pn_fma(m, vp[r * CB + l], acc[r]);
    }
}

template <int NR, bool BWD>
__global__ void __launch_bounds__(SC_THREADS, 1) chol_chain_solve_kernel(const __grid_constant__ ChChain p) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* stage = reinterpret_cast<double*>(smem);  // [SC_STAGES][SC_UR][CB]
    double* vs = stage + SC_STAGES * SC_UNIT;         // [2][NR][CB]: the operand of the current / next block
    double* bs = vs + 2 * NR * CB;                    // [CB][NR]: this CTA's block of the right-hand side
    double* red = bs + CB * NR;                       // [NR][CB]
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(p.ticket, 1);
    __syncthreads();
    const int t = s_ticket;  // = number of off-diagonal blocks of this CTA
    if (t >= p.T) return;
    const int i = BWD ? p.T - 1 - t : t;
    double* bi = p.b + (int64_t)i * CB * p.ldb;
    for (int idx = tid; idx < CB * NR; idx += SC_THREADS) bs[idx] = bi[(idx / NR) * p.ldb + idx % NR];
    __syncthreads();
    const int U = 2 * (t + 1);
    auto issue = [&](int u) {
        if (u < U) {
            const int s = u >> 1;
            const double* src;
            int64_t ld;
            if (s < t) {
                const int j = BWD ? p.T - 1 - s : s;
                src = BWD ? p.M + ((int64_t)j * CB) * p.ld + (int64_t)i * CB : p.M + ((int64_t)i * CB) * p.ld + (int64_t)j * CB;
                ld = p.ld;
            } else {
                src = p.invd + (int64_t)i * CB * CB;
                ld = CB;
            }
            src += (int64_t)(u & 1) * SC_UR * ld;
            double* dst = stage + (u % SC_STAGES) * SC_UNIT;
#pragma unroll
            for (int k = 0; k < SC_UNIT / 2 / SC_THREADS; ++k) {
                const int c = tid + SC_THREADS * k, row = c >> 6, q = c & 63;
                ch_cp_async16(dst + row * CB + 2 * q, src + (int64_t)row * ld + 2 * q);
            }
        }
        ch_cp_commit();
    };
    constexpr int NACC = BWD ? 1 : 16;
    double acc[NR][NACC];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[r][a] = 0.0;
    const int col = tid & (CB - 1), half = tid >> 7;  // backward: a thread per column, the rows on the two thread halves
    // totals of the partial sums: fin(r, row, total) is called by exactly one thread per (row, r)
    auto reduce = [&](auto fin) {
        if constexpr (!BWD) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double tot = sc_reduce16(acc[r], lane);
                const int a = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                if (!(lane & 1)) fin(r, 64 * (a >> 3) + warp + 8 * (a & 7), tot);
            }
        } else {
            if (half) {
#pragma unroll
                for (int r = 0; r < NR; ++r) red[r * CB + col] = acc[r][0];
            }
            __syncthreads();
            if (!half) {
#pragma unroll
                for (int r = 0; r < NR; ++r) fin(r, col, acc[r][0] + red[r * CB + col]);
            }
        }
    };
    issue(0);
    issue(1);
    for (int s = 0; s <= t; ++s) {
        double* v = vs + (s & 1) * NR * CB;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int u = 2 * s + h;
            ch_cp_wait<1>();
            if (h == 0) {
                if (s < t) {
                    if (warp == 0) {  // the other warps are still on the previous block or wait at the barrier below
                        const int j = BWD ? p.T - 1 - s : s;
                        while (sc_ld_acquire(p.flags + j) == 0) {
                        }
                        const double* xj = p.b + (int64_t)j * CB * p.ldb;
                        for (int idx = lane; idx < CB * NR; idx += 32)
                            v[(idx % NR) * CB + idx / NR] = __ldcg(xj + (idx / NR) * p.ldb + idx % NR);
                    }
                } else {
                    // every off-diagonal block is in: own block -= sums; it becomes the operand of inv(L_ii)
                    reduce([&](int r, int row, double tot) { bs[row * NR + r] -= tot; });
                    __syncthreads();
                    for (int idx = tid; idx < CB * NR; idx += SC_THREADS) v[(idx % NR) * CB + idx / NR] = bs[idx];
#pragma unroll
                    for (int r = 0; r < NR; ++r)
#pragma unroll
                        for (int a = 0; a < NACC; ++a) acc[r][a] = 0.0;
                }
            }
            __syncthreads();  // unit u and the operand are in shared memory; everybody is done with unit u - 1
            issue(u + 2);     // into the stage unit u - 1 occupied
            const double* S = stage + (u % SC_STAGES) * SC_UNIT;
            if constexpr (!BWD) {
                if (h == 0) sc_apply_rows<NR, 0>(S, v, acc, warp, lane);
                else sc_apply_rows<NR, 1>(S, v, acc, warp, lane);
            } else {
                double a1[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) a1[r] = acc[r][0];
                sc_apply_cols<NR>(S, v, h, a1, col, half);
#pragma unroll
                for (int r = 0; r < NR; ++r) acc[r][0] = a1[r];
            }
        }
    }
    reduce([&](int r, int row, double tot) { bs[row * NR + r] = tot; });
    __syncthreads();
    for (int idx = tid; idx < CB * NR; idx += SC_THREADS) bi[(idx / NR) * p.ldb + idx % NR] = bs[idx];
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        sc_st_release(p.flags + i, 1);
    }
}

// ---- solves with many right-hand sides ------------------------------------------
// The prediction path solves (L L^T) W = K(nodes, x) for hundreds to thousands of columns.  The
// right-hand sides are kept TRANSPOSED, Z = K(x, nodes) (m x npad row-major, one prediction point per
// row), so that every step of the blocked substitution is a product with K-contiguous left operand:
//   forward  (k = 0 .. T-1):  Y_k = Z_k inv(L_kk)^T          ;  Z_i -= Y_k L_ik^T  for all i > k
//   backward (k = T-1 .. 0):  X_k = Y_k inv(L_kk)            ;  Y_j -= X_k L_kj    for all j < k
// The forward products have the second operand K-contiguous ("row.col", as the factorisation), the
// backward ones N-contiguous; the kernel stages that operand as [k][n] with a row stride of 72
// doubles (= 8 mod 16: conflict-free fragment loads).  Solved blocks are written to the OTHER of two
// buffers (forward Z -> Y, backward Y -> Z): no CTA ever reads a location another CTA writes.
constexpr int WLDN = CTN + 8;

struct ChWide {
    const double* A;  // (rows, 128) K-contiguous, leading dimension lda
    int64_t lda;
    const double* B;  // BKN = false: (cols, 128) K-contiguous; BKN = true: (128, cols) N-contiguous
    int64_t ldb;
    double* C;        // (rows, cols), leading dimension ldc
    int64_t ldc;
    int subtract;     // 0: C = A op(B), 1: C -= A op(B)
};

template <bool BKN>
__global__ void __launch_bounds__(CG_THREADS, 3) chol_wide_gemm_kernel(const __grid_constant__ ChWide p) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* As = reinterpret_cast<double*>(smem);  // [2][CB * CLD]
    double* Bs = As + 2 * CB * CLD;                // [2][CTN * CLD] or [2][CK * WLDN]
    constexpr int BSTAGE = BKN ? CK * WLDN : CTN * CLD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const double* At = p.A + (int64_t)blockIdx.y * CB * p.lda;
    const double* Bt = BKN ? p.B + (int64_t)blockIdx.x * CTN : p.B + (int64_t)blockIdx.x * CTN * p.ldb;
    double* Ct = p.C + (int64_t)blockIdx.y * CB * p.ldc + (int64_t)blockIdx.x * CTN;

    auto load_chunk = [&](int kc, int buf) {
        double* as = As + buf * CB * CLD;
        double* bs = Bs + buf * BSTAGE;
        constexpr int PIECES = CK / 2;
#pragma unroll
        for (int i = 0; i < CB * PIECES / CG_THREADS; ++i) {
            const int idx = tid + i * CG_THREADS;
            const int row = idx / PIECES, c2 = (idx % PIECES) * 2;
            ch_cp_async16(as + row * CLD + c2, At + (int64_t)row * p.lda + kc * CK + c2);
        }
        if (BKN) {
            constexpr int NP = CTN / 2;  // 16-byte pieces per k row
#pragma unroll
            for (int i = 0; i < CK * NP / CG_THREADS; ++i) {
                const int idx = tid + i * CG_THREADS;
                const int kr = idx / NP, c2 = (idx % NP) * 2;
                ch_cp_async16(bs + kr * WLDN + c2, Bt + (int64_t)(kc * CK + kr) * p.ldb + c2);
            }
        } else {
#pragma unroll
            for (int i = 0; i < CTN * PIECES / CG_THREADS; ++i) {
                const int idx = tid + i * CG_THREADS;
                const int row = idx / PIECES, c2 = (idx % PIECES) * 2;
                ch_cp_async16(bs + row * CLD + c2, Bt + (int64_t)row * p.ldb + kc * CK + c2);
            }
        }
        ch_cp_commit();
    };

    double acc[4][8][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;

    load_chunk(0, 0);
    constexpr int NCHUNK = CB / CK;
    for (int kc = 0; kc < NCHUNK; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < NCHUNK) {
            load_chunk(kc + 1, buf ^ 1);
            ch_cp_wait<1>();
        } else {
            ch_cp_wait<0>();
        }
        __syncthreads();
        const double* as = As + buf * CB * CLD + (warp * 32 + g) * CLD + q;
        const double* bs = Bs + buf * BSTAGE + (BKN ? q * WLDN + g : g * CLD + q);
#pragma unroll
        for (int k4 = 0; k4 < CK / 4; ++k4) {
            double a[4], b[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = as[r * 8 * CLD + k4 * 4];
#pragma unroll
            for (int c = 0; c < 8; ++c) b[c] = BKN ? bs[k4 * 4 * WLDN + c * 8] : bs[c * 8 * CLD + k4 * 4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) ch_dmma(acc[r][c][0], acc[r][c][1], a[r], b[c]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = warp * 32 + r * 8 + g;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double2* dst = reinterpret_cast<double2*>(Ct + (int64_t)row * p.ldc + c * 8 + 2 * q);
            if (p.subtract) {
                double2 v = *dst;
                v.x -= acc[r][c][0];
                v.y -= acc[r][c][1];
                *dst = v;
            } else {
                *dst = make_double2(acc[r][c][0], acc[r][c][1]);
            }
        }
    }
}

// Row reductions of the prediction path: quad[j] = sum_i W[j,i] G[j,i] and dots[t,j] = sum_i vecs[t,i] W[j,i]
// (W: the solved right-hand sides, G: the Gram block again, both (rows, ld); vecs: (nvec, ld), nvec <= 4).
// One CTA per row, fixed summation order.
constexpr int WR_THREADS = 256;
__global__ void __launch_bounds__(WR_THREADS) chol_rowdots_kernel(const double* __restrict__ W, const double* __restrict__ G,
                                                                  int64_t ld, int64_t n, const double* __restrict__ vecs,
                                                                  int nvec, double* __restrict__ quad,
                                                                  double* __restrict__ dots, int64_t ldd) {
    __shared__ double red[5][WR_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* w = W + (int64_t)blockIdx.x * ld;
    const double* gr = G + (int64_t)blockIdx.x * ld;
    double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t i = tid; i < n; i += WR_THREADS) {
        const double wi = w[i];
        s[0] = pn_fma(wi, gr[i], s[0]);
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (t < nvec) s[t + 1] = pn_fma(wi, vecs[(int64_t)t * ld + i], s[t + 1]);
    }
#pragma unroll
    for (int t = 0; t < 5; ++t) {
        double v = s[t];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (lane == 0) red[t][warp] = v;
    }
    __syncthreads();
    if (tid < 5) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < WR_THREADS / 32; ++k) v += red[tid][k];
        if (tid == 0) quad[blockIdx.x] = v;
        else if (tid - 1 < nvec) dots[(int64_t)(tid - 1) * ldd + blockIdx.x] = v;
    }
}

}  // namespace

// M: (npad, npad) row-major with ld = npad (a multiple of 128); the Gram matrix occupies the leading
// n x n block (lower triangle is read).  On return the lower triangle holds L, invd the inverses
// of the diagonal blocks of L; *info_dev the 1-based index of the first bad pivot (0: success, -1: a hand-over wait
// gave up).  panel: TWO contiguous block columns (2 x npad x 128 doubles); scratch: npad / 128 ints.
//
// Schedule (two streams; D = diagonal block, P = panel, U = trailing update of a step):
//   main:  U_0            | U_1            | U_2 ...         U_k waits for P_k; consecutive updates stay in stream order
//   side:  D_1 P_1        | D_2 P_2        | ...             D_{k+1} waits for U_{k-1}, P_{k+1} for D_{k+1}
// D_{k+1} applies its own tile of update k itself (from the head of panel k), so it needs nothing of U_k.  P_{k+1}
// needs block column k + 1 AFTER update k: U_k runs the tiles of that column first and counts them into `scratch`
// per block row; the CTAs of P_{k+1} wait for their row's count (chol_gemm_kernel).  With that the panel step -- a
// launch that cannot fill the machine -- runs under the update instead of between two updates.  P_{k+1} writes
// the panel buffer U_k does not read.  A waiting panel CTA holds an SM slot: the overlap is used only while the
// panel's grid leaves at least one of the three slots of every SM to the update (always up to n = 19072).
cudaError_t pn_chol_factor(double* M, int64_t n, int64_t npad, double jitter, double* panel, double* invd,
                           int* info_dev, int* scratch, cudaStream_t stream, cudaStream_t side, cudaEvent_t ev_panel,
                           cudaEvent_t ev_update) {
    cudaError_t err;
    const int64_t T = npad / CB;
    if ((err = cudaMemsetAsync(info_dev, 0, sizeof(int), stream)) != cudaSuccess) return err;
    if ((err = cudaMemsetAsync(scratch, 0, sizeof(int) * (size_t)T, stream)) != cudaSuccess) return err;
    chol_prepare_diag_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(M, npad, n, npad, jitter);
    ++g_pn_launches;
    const size_t smem_diag = sizeof(double) * (CB * CD_LD + 5 * CB + 3 * 32 * 33);
    const size_t smem_gemm = sizeof(double) * 2 * (CB + CTN) * CLD;
    if ((err = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_diag)) != cudaSuccess) return err;
    auto gemm = chol_gemm_kernel;
    if ((err = cudaFuncSetAttribute(gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gemm)) != cudaSuccess) return err;
    const int max_waiting = 2 * pn_sm_count();  // panel CTAs that may wait while an update runs
    struct LocalEvent {  // released (after the work that refers to it) on every way out
        cudaEvent_t e = nullptr;
        ~LocalEvent() {
            if (e) cudaEventDestroy(e);
        }
    } local;
    if ((err = cudaEventCreateWithFlags(&local.e, cudaEventDisableTiming)) != cudaSuccess) return err;
    const cudaEvent_t ev_diag = local.e;
    static const int kOverlapMinBlocks = getenv("PNCOV_CHOL_OVERLAP_MIN") ? atoi(getenv("PNCOV_CHOL_OVERLAP_MIN")) : 32;
    double* pbuf[2] = {panel, panel + npad * CB};
    // panel of block column k into pbuf[k & 1]; the update of step k copies it into the matrix
    auto launch_panel = [&](int64_t k, int target, cudaStream_t st) -> cudaError_t {
        const int64_t nb = T - k - 1;
        if (nb <= 0) return cudaSuccess;
        ChGemm p{};
        p.A = M + k * CB * (npad + 1) + CB * npad;  // block (k + 1, k)
        p.lda = npad;
        p.B = invd + k * CB * CB;
        p.ldb = CB;
        p.C = const_cast<double*>(p.A);
        p.ldc = npad;
        p.panel = pbuf[k & 1];
        p.mode = 0;
        p.rowcount = scratch;
        p.row0 = (int)(k + 1);
        p.target = target;
        p.err = info_dev;
        gemm<<<(unsigned)(2 * nb), CG_THREADS, smem_gemm, st>>>(p);
        ++g_pn_launches;
        return cudaSuccess;
    };
    chol_diag_kernel<<<1, 256, smem_diag, stream>>>(M, npad, invd, info_dev, 0, nullptr);
    ++g_pn_launches;
    if ((err = launch_panel(0, 0, stream)) != cudaSuccess) return err;
    if ((err = cudaEventRecord(ev_panel, stream)) != cudaSuccess) return err;  // P_0
    if ((err = cudaEventRecord(ev_update, stream)) != cudaSuccess) return err;  // "U_{-1}"
    for (int64_t k = 0; k + 1 < T; ++k) {
        const int64_t nb = T - k - 1;  // block rows of the trailing matrix
        double* trail = M + (k + 1) * CB * (npad + 1);  // block (k + 1, k + 1)
        // side: wait for U_{k-1} (its event is re-recorded below, so this call comes first) and P_k
        if ((err = cudaStreamWaitEvent(side, ev_update, 0)) != cudaSuccess) return err;
        if ((err = cudaStreamWaitEvent(side, ev_panel, 0)) != cudaSuccess) return err;
        // side: D_{k+1} (own tile of update k from the head of panel k fused in); issued BEFORE the update so that its
        // one large CTA (135 KB of shared memory) finds an SM before the update's CTAs fill them all
        chol_diag_kernel<<<1, 256, smem_diag, side>>>(trail, npad, invd + (k + 1) * CB * CB, info_dev, (int)((k + 1) * CB),
                                                      pbuf[k & 1]);
        ++g_pn_launches;
        // main: U_k, every tile but block (k + 1, k + 1)
        if ((err = cudaStreamWaitEvent(stream, ev_panel, 0)) != cudaSuccess) return err;
        const int64_t tiles = nb * (nb + 1) / 2 - 1;
        // panel k into the matrix: in the last CTAs of a long update; ahead of a short one as a copy of its own, which
        // also gives D_{k+1}, released by the same event, the head start it needs to find an SM
        const bool fused_copy = nb >= kOverlapMinBlocks;
        if (!fused_copy &&
            (err = cudaMemcpy2DAsync(trail - CB, (size_t)npad * 8, pbuf[k & 1], (size_t)CB * 8, (size_t)CB * 8, (size_t)(nb * CB),
                                     cudaMemcpyDeviceToDevice, stream)) != cudaSuccess)
            return err;
        if (tiles > 0 || fused_copy) {
            ChGemm u{};
            u.A = pbuf[k & 1];
            u.lda = CB;
            u.B = pbuf[k & 1];
            u.ldb = CB;
            u.C = trail;
            u.ldc = npad;
            u.mode = 1;
            u.nb = (int)nb;
            u.rowcount = scratch;
            u.row0 = (int)(k + 1);
            u.Lcol = trail - CB;  // block (k + 1, k)
            gemm<<<(unsigned)(2 * (tiles + (fused_copy ? nb : 0))), CG_THREADS, smem_gemm, stream>>>(u);
            ++g_pn_launches;
        }
        if ((err = cudaEventRecord(ev_update, stream)) != cudaSuccess) return err;
        // side: P_{k+1}, after the update in issue order (its CTAs may wait for tiles of U_k)
        // worth it only while the update outlasts the diagonal block and the panel (0.07 nb^2 us against ~90 us):
        // under a short update the panel would share the SMs with it and both sit on the critical path
        const bool overlap = nb - 1 > 0 && 2 * (nb - 1) <= max_waiting && nb >= kOverlapMinBlocks;
        if (overlap) {
            if ((err = launch_panel(k + 1, (int)(2 * (k + 1)), side)) != cudaSuccess) return err;
            if ((err = cudaEventRecord(ev_panel, side)) != cudaSuccess) return err;  // D_{k+1} and P_{k+1}
        } else {
            // P_{k+1} on the main stream behind U_k, once D_{k+1} is done
            if ((err = cudaEventRecord(ev_diag, side)) != cudaSuccess) return err;
            if ((err = cudaStreamWaitEvent(stream, ev_diag, 0)) != cudaSuccess) return err;
            if ((err = launch_panel(k + 1, 0, stream)) != cudaSuccess) return err;
            if ((err = cudaEventRecord(ev_panel, stream)) != cudaSuccess) return err;
        }
    }
    if ((err = cudaStreamWaitEvent(stream, ev_panel, 0)) != cudaSuccess) return err;
    return cudaGetLastError();
}

// b: nrhs <= 3 adjacent columns of a row-major array with npad rows and row stride ldb, overwritten with the
// solution of L L^T x = b.  sync: 2 + 2 npad / 128 ints of device scratch (tickets and flags of the two launches).
namespace {
template <int NR>
cudaError_t chol_chain_launch(const double* M, int64_t npad, const double* invd, double* b, int64_t ldb, int* sync,
                              cudaStream_t stream) {
    const int T = (int)(npad / CB);
    const size_t smem = sizeof(double) * (SC_STAGES * SC_UNIT + 2 * NR * CB + 2 * CB * NR);
    cudaError_t err;
    if ((err = cudaFuncSetAttribute(chol_chain_solve_kernel<NR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(chol_chain_solve_kernel<NR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
    if ((err = cudaMemsetAsync(sync, 0, sizeof(int) * (2 + 2 * (size_t)T), stream)) != cudaSuccess) return err;
    ChChain p{M, npad, invd, b, ldb, sync, sync + 2, T};
    chol_chain_solve_kernel<NR, false><<<(unsigned)T, SC_THREADS, smem, stream>>>(p);  // L y = b
    ++g_pn_launches;
    p.ticket = sync + 1;
    p.flags = sync + 2 + T;
    chol_chain_solve_kernel<NR, true><<<(unsigned)T, SC_THREADS, smem, stream>>>(p);  // L^T x = y
    ++g_pn_launches;
    return cudaGetLastError();
}
}  // namespace

cudaError_t pn_chol_solve(const double* M, int64_t npad, const double* invd, double* b, int64_t ldb, int nrhs, int* sync,
                          cudaStream_t stream) {
    switch (nrhs) {
        case 1: return chol_chain_launch<1>(M, npad, invd, b, ldb, sync, stream);
        case 2: return chol_chain_launch<2>(M, npad, invd, b, ldb, sync, stream);
        case 3: return chol_chain_launch<3>(M, npad, invd, b, ldb, sync, stream);
        default: return cudaErrorInvalidValue;
    }
}

// Z, Y: (mpad, npad) row-major (mpad a multiple of 128).  In: Z = right-hand sides, one per ROW (the
// transposed layout of the kernel above); out: Z = the solutions of L L^T w = z for every row, Y scratch.
// forward_only: stop after L y = z; the solutions y are the rows of Y (Z is clobbered).
cudaError_t pn_chol_solve_wide(const double* M, int64_t npad, const double* invd, double* Z, double* Y, int64_t mpad,
                               int forward_only, cudaStream_t stream, cudaStream_t side, cudaEvent_t ev_main,
                               cudaEvent_t ev_side) {
    cudaError_t err;
    const size_t smem = sizeof(double) * (2 * CB * CLD + 2 * (CTN * CLD > CK * WLDN ? CTN * CLD : CK * WLDN));
    if ((err = cudaFuncSetAttribute(chol_wide_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(chol_wide_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
    const int64_t T = npad / CB;
    const unsigned mt = (unsigned)(mpad / CB);
    if (mt > 65535) return cudaErrorInvalidConfiguration;
    auto gemm = [&](bool bkn, const double* A, const double* B, int64_t ldb, double* C, int subtract, int64_t cols,
                    cudaStream_t st) {
        ChWide p{};
        p.A = A;
        p.lda = npad;
        p.B = B;
        p.ldb = ldb;
        p.C = C;
        p.ldc = npad;
        p.subtract = subtract;
        const dim3 grid((unsigned)(cols / CTN), mt);
        if (bkn) chol_wide_gemm_kernel<true><<<grid, CG_THREADS, smem, st>>>(p);
        else chol_wide_gemm_kernel<false><<<grid, CG_THREADS, smem, st>>>(p);
        ++g_pn_launches;
    };
    // Look-ahead.  Step k of a pass is: the product with the inverse diagonal block ("diag", 2 x mt CTAs,
    // latency-bound), the update of the NEXT block ("head", same size) and the update of everything
    // beyond it ("rest", the bulk).  diag and head run on the high-priority side stream under the rest of
    // the previous step, which stays alone on the critical path:
    //   side:  diag_0 | head_0  diag_1 | head_1  diag_2 | ...     head_k waits for rest_{k-1} (same block)
    //   main:         | rest_0         | rest_1         | ...     rest_k waits for diag_k
    // head_k and rest_k touch different blocks; consecutive rests share blocks and stay in stream order.
    if ((err = cudaEventRecord(ev_main, stream)) != cudaSuccess) return err;  // the right-hand sides are in place
    if ((err = cudaStreamWaitEvent(side, ev_main, 0)) != cudaSuccess) return err;
    gemm(false, Z, invd, CB, Y, 0, CB, side);  // diag_0: Y_0 = Z_0 inv(L_00)^T
    if ((err = cudaEventRecord(ev_side, side)) != cudaSuccess) return err;
    for (int64_t k = 0; k + 1 < T; ++k) {  // L Y = Z
        if ((err = cudaStreamWaitEvent(stream, ev_side, 0)) != cudaSuccess) return err;  // Y_k
        if (k > 0 && (err = cudaStreamWaitEvent(side, ev_main, 0)) != cudaSuccess) return err;  // rest_{k-1}
        const double* Yk = Y + k * CB;
        const double* Lk = M + (k + 1) * CB * npad + k * CB;  // L[i, k] for the block rows i > k
        gemm(false, Yk, Lk, npad, Z + (k + 1) * CB, 1, CB, side);                                      // head_k
        gemm(false, Z + (k + 1) * CB, invd + (k + 1) * CB * CB, CB, Y + (k + 1) * CB, 0, CB, side);  // diag_{k+1}
        if ((err = cudaEventRecord(ev_side, side)) != cudaSuccess) return err;
        const int64_t rest = npad - (k + 2) * CB;
        if (rest > 0) gemm(false, Yk, Lk + CB * npad, npad, Z + (k + 2) * CB, 1, rest, stream);      // rest_k
        if ((err = cudaEventRecord(ev_main, stream)) != cudaSuccess) return err;
    }
    if ((err = cudaStreamWaitEvent(stream, ev_side, 0)) != cudaSuccess) return err;
    if (forward_only) return cudaGetLastError();
    // L^T X = Y, blocks from the last to the first; here main has just waited for the side stream
    if ((err = cudaEventRecord(ev_main, stream)) != cudaSuccess) return err;
    if ((err = cudaStreamWaitEvent(side, ev_main, 0)) != cudaSuccess) return err;
    gemm(true, Y + (T - 1) * CB, invd + (T - 1) * CB * CB, CB, Z + (T - 1) * CB, 0, CB, side);  // X_{T-1}
    if ((err = cudaEventRecord(ev_side, side)) != cudaSuccess) return err;
    for (int64_t k = T - 1; k > 0; --k) {
        if ((err = cudaStreamWaitEvent(stream, ev_side, 0)) != cudaSuccess) return err;  // X_k
        if (k < T - 1 && (err = cudaStreamWaitEvent(side, ev_main, 0)) != cudaSuccess) return err;  // rest_{k+1}
        const double* Xk = Z + k * CB;
        const double* Lk = M + k * CB * npad;  // L[k, j] for the block columns j < k
        gemm(true, Xk, Lk + (k - 1) * CB, npad, Y + (k - 1) * CB, 1, CB, side);                            // head_k
        gemm(true, Y + (k - 1) * CB, invd + (k - 1) * CB * CB, CB, Z + (k - 1) * CB, 0, CB, side);       // diag_{k-1}
        if ((err = cudaEventRecord(ev_side, side)) != cudaSuccess) return err;
        if (k > 1) gemm(true, Xk, Lk, npad, Y, 1, (k - 1) * CB, stream);                                   // rest_k
        if ((err = cudaEventRecord(ev_main, stream)) != cudaSuccess) return err;
    }
    if ((err = cudaStreamWaitEvent(stream, ev_side, 0)) != cudaSuccess) return err;
    return cudaGetLastError();
}

// dst (cols, rows; ldd) = src (rows, cols; lds)^T, 32 x 32 tiles through shared memory
namespace {
__global__ void __launch_bounds__(256) chol_transpose_kernel(const double* __restrict__ src, int64_t lds, int64_t rows,
                                                             int64_t cols, double* __restrict__ dst, int64_t ldd) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + 8 * i][tx] = src[r * lds + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
        if (r < rows && c < cols) dst[c * ldd + r] = tile[tx][ty + 8 * i];
    }
}
}  // namespace

cudaError_t pn_chol_transpose(const double* src, int64_t lds, int64_t rows, int64_t cols, double* dst, int64_t ldd,
                              cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    const dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    if (grid.y > 65535) return cudaErrorInvalidConfiguration;
    chol_transpose_kernel<<<grid, 256, 0, stream>>>(src, lds, rows, cols, dst, ldd);
    ++g_pn_launches;
    return cudaGetLastError();
}

// M[i][j] = M[j][i] for all i > j of the leading n x n block (row-major, ld): the lower triangle becomes the mirror
// image of the upper one.  32 x 32 tiles through shared memory; tiles left of the diagonal exit.
namespace {
__global__ void __launch_bounds__(256) chol_mirror_upper_kernel(double* __restrict__ M, int64_t ld, int64_t n) {
    if (blockIdx.x < blockIdx.y) return;
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;  // source tile: rows r0.., columns c0.. >= r0
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < n && c < n) tile[ty + 8 * i][tx] = M[r * ld + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t c = c0 + ty + 8 * i, r = r0 + tx;  // destination M[c][r] = source M[r][c]
        if (r < n && c < n && c > r) M[c * ld + r] = tile[tx][ty + 8 * i];
    }
}
}  // namespace

cudaError_t pn_chol_mirror_upper(double* M, int64_t ld, int64_t n, cudaStream_t stream) {
    const unsigned t = (unsigned)((n + 31) / 32);
    if (t == 0) return cudaSuccess;
    if (t > 65535) return cudaErrorInvalidConfiguration;
    chol_mirror_upper_kernel<<<dim3(t, t), 256, 0, stream>>>(M, ld, n);
    ++g_pn_launches;
    return cudaGetLastError();
}

cudaError_t pn_chol_rowdots(const double* W, const double* G, int64_t ld, int64_t n, int64_t rows, const double* vecs,
                            int nvec, double* quad, double* dots, int64_t ldd, cudaStream_t stream) {
    if (rows <= 0) return cudaSuccess;
    chol_rowdots_kernel<<<(unsigned)rows, WR_THREADS, 0, stream>>>(W, G, ld, n, vecs, nvec, quad, dots, ldd);
    ++g_pn_launches;
    return cudaGetLastError();
}
