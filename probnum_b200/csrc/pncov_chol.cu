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

struct ChGemm {
    const double* A;  // mode 0: first row block of the panel inside the matrix; mode 1: contiguous panel
    int64_t lda;
    const double* B;  // mode 0: inverse of the diagonal factor (128 x 128); mode 1: contiguous panel
    int64_t ldb;
    double* C;        // mode 0: unused (the panel is copied over A afterwards); mode 1: first trailing tile
    int64_t ldc;
    double* panel;    // mode 0: contiguous copy of the new block column (ld = 128)
    int mode;
    int tile0;        // index of the first half tile of this launch (mode 1: half tiles 0, 1 = the next diagonal block)
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
    const int64_t t2 = (int64_t)blockIdx.x + p.tile0;
    const int h = (int)(t2 & 1);
    if (p.mode == 0) {
        bi = t2 >> 1;
        bj = 0;
    } else {  // linear index over the lower triangle (bj <= bi)
        const int64_t t = t2 >> 1;
        bi = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
        while (bi * (bi + 1) / 2 > t) --bi;
        bj = t - bi * (bi + 1) / 2;
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
}

// 16 elimination steps j = 16 JJ .. 16 JJ + 15 of the in-register Cholesky of chol_diag_kernel.
// Thread (ty, tx) owns a[ty + 16 ii][tx + 16 cc] in r[ii][cc]; column block JJ is a compile-time
// constant, so every register index is static and finished column blocks cost nothing.
template <int JJ>
__device__ __forceinline__ void chol_diag_steps(double (&r)[8][8], double* colbuf, int ty, int tx, int tid, int* info,
                                                int pivot_base) {
#pragma unroll 1
    for (int jt = 0; jt < 16; ++jt) {
        const int j = 16 * JJ + jt;
        double* cb = colbuf + (j & 1) * CB;  // double-buffered: one barrier per step
        if (tx == jt) {                       // the owners of column j publish it (unscaled)
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) cb[ty + 16 * ii] = r[ii][JJ];
        }
        __syncthreads();
        double ajj = cb[j];
        if (!(ajj > 0.0)) {  // not positive definite (or NaN): record the first offender, keep going
            if (tid == 0) atomicCAS(info, 0, pivot_base + j + 1);
            ajj = 1.0;
        }
        const double dj = rsqrt(ajj);
        double li[8], lc[8 - JJ];
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) li[ii] = cb[ty + 16 * ii] * dj;  // L[i][j] for my rows (i >= j)
#pragma unroll
        for (int cc = JJ; cc < 8; ++cc) lc[cc - JJ] = cb[tx + 16 * cc] * dj;  // L[c][j] for my columns
        // a[i][c] -= L[i][j] L[c][j] for the columns right of j; rows above the diagonal hold dead
        // values that never feed a live one, so they need no predicate
#pragma unroll
        for (int cc = JJ; cc < 8; ++cc) {
            const bool upd = (cc > JJ) || (tx > jt);
#pragma unroll
            for (int ii = 0; ii < 8; ++ii)
                if (upd) r[ii][cc] = pn_fma(-li[ii], lc[cc - JJ], r[ii][cc]);
        }
        if (tx == jt) {  // column j is final: L[j][j] = a_jj / sqrt(a_jj), L[i][j] = a_ij / sqrt(a_jj)
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) r[ii][JJ] = li[ii];
        }
    }
}

// One CTA: Cholesky of the 128 x 128 diagonal block (lower triangle of A, row-major) and the
// inverse of its factor.  The block lives in registers (8 x 8 per thread, cyclic 16 x 16
// distribution); per elimination step the owners publish one column through shared memory, one
// barrier, and everybody updates its registers.  info receives the 1-based index of the first
// non-positive pivot.
// P != nullptr (look-ahead): the block first receives its own trailing update A -= P P^T from the 128 x 128
// head of the current panel, so the kernel depends on the panel alone and can take an SM before the
// wide update of the same step fills the machine.
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ invL,
                                                         int* __restrict__ info, int pivot_base,
                                                         const double* __restrict__ P) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* a = reinterpret_cast<double*>(smem);  // [CB][CD_LD]: upper triangle L^T, lower triangle inv(L)
    double* dg = a + CB * CD_LD;                  // [CB] diagonal of L
    double* colbuf = dg + CB;                     // [2][CB]
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    double r[8][8];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int i = ty + 16 * ii, c = tx + 16 * cc;
            r[ii][cc] = (c <= i) ? A[(int64_t)i * lda + c] : 0.0;
        }
    if (P != nullptr) {
        for (int idx = tid; idx < CB * CB; idx += 256) a[(idx >> 7) * CD_LD + (idx & (CB - 1))] = P[idx];
        __syncthreads();
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
                for (int cc = 0; cc < 8; ++cc) r[ii][cc] = pn_fma(-vi[ii], vc[cc], r[ii][cc]);
        }
#pragma unroll
        for (int ii = 0; ii < 8; ++ii)
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
                if (tx + 16 * cc > ty + 16 * ii) r[ii][cc] = 0.0;
        // `a` is written again only after the elimination steps, whose barriers order it after these reads
    }
    chol_diag_steps<0>(r, colbuf, ty, tx, tid, info, pivot_base);
    chol_diag_steps<1>(r, colbuf, ty, tx, tid, info, pivot_base);
    chol_diag_steps<2>(r, colbuf, ty, tx, tid, info, pivot_base);
    chol_diag_steps<3>(r, colbuf, ty, tx, tid, info, pivot_base);
    chol_diag_steps<4>(r, colbuf, ty, tx, tid, info, pivot_base);
    chol_diag_steps<5>(r, colbuf, ty, tx, tid, info, pivot_base);
    chol_diag_steps<6>(r, colbuf, ty, tx, tid, info, pivot_base);
    chol_diag_steps<7>(r, colbuf, ty, tx, tid, info, pivot_base);
    // L to global memory (lower triangle) and, transposed, to shared memory for the inverse
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int i = ty + 16 * ii, c = tx + 16 * cc;
            if (c <= i) A[(int64_t)i * lda + c] = r[ii][cc];
            if (c < i) a[c * CD_LD + i] = r[ii][cc];
            else if (c == i) dg[i] = r[ii][cc];
        }
    __syncthreads();
    // inverse X of L, blocked 4 x 4 with 32 x 32 blocks.  L[i][l] sits at a[l][i] (upper triangle),
    // X[i][c] goes to a[i][c] (lower triangle, free now).
    // (A) diagonal blocks: forward substitution inside the block, one column per thread
    if (tid < CB) {
        const int c = tid, b0 = (tid >> 5) << 5;
        for (int i = b0; i < b0 + 32; ++i) {
            double s0 = (i == c) ? 1.0 : 0.0, s1 = 0.0;
            int l = b0;
            for (; l + 1 < i; l += 2) {
                s0 = pn_fma(-a[l * CD_LD + i], (l >= c) ? a[l * CD_LD + c] : 0.0, s0);
                s1 = pn_fma(-a[(l + 1) * CD_LD + i], (l + 1 >= c) ? a[(l + 1) * CD_LD + c] : 0.0, s1);
            }
            if (l < i) s0 = pn_fma(-a[l * CD_LD + i], (l >= c) ? a[l * CD_LD + c] : 0.0, s0);
            if (i >= c) a[i * CD_LD + c] = (s0 + s1) / dg[i];  // own column only
        }
    }
    __syncthreads();
    // (B) off-diagonal blocks by block distance: X_ib = -X_ii (sum_{m=b}^{i-1} L_im X_mb); the blocks
    // of one distance are independent.  The inner sums T (up to 3 blocks of 32 x 32 at a time, row
    // stride 33) have their own shared-memory area: both triangles of `a` are still in use.
    double* T = colbuf + 2 * CB;
    const int cc = tid & 31, rr0 = tid >> 5;  // a warp walks 32 adjacent columns of one row: L broadcasts
    for (int dist = 1; dist < 4; ++dist) {
        const int nblk = 4 - dist;
        for (int o = rr0; o < nblk * 32; o += 8) {  // T[blk][rr][cc]
            const int blk = o >> 5, rr = o & 31;
            const int b = blk, i = blk + dist;
            const int row = 32 * i + rr, col = 32 * b + cc;
            double s0 = 0.0, s1 = 0.0;
            for (int l = 32 * b; l < 32 * i; l += 2) {
                s0 = pn_fma(a[l * CD_LD + row], (l >= col) ? a[l * CD_LD + col] : 0.0, s0);
                s1 = pn_fma(a[(l + 1) * CD_LD + row], (l + 1 >= col) ? a[(l + 1) * CD_LD + col] : 0.0, s1);
            }
            T[(blk * 32 + rr) * 33 + cc] = s0 + s1;
        }
        __syncthreads();
        for (int o = rr0; o < nblk * 32; o += 8) {  // X_ib = -X_ii T
            const int blk = o >> 5, rr = o & 31;
            const int b = blk, i = blk + dist;
            const int row = 32 * i + rr, col = 32 * b + cc;
            double s = 0.0;
            for (int k = 0; k <= rr; ++k) s = pn_fma(a[row * CD_LD + 32 * i + k], T[(blk * 32 + k) * 33 + cc], s);
            a[row * CD_LD + col] = -s;
        }
        __syncthreads();
    }
    // X to global memory, row-major, zeros above the diagonal
    for (int idx = tid; idx < CB * CB; idx += 256) {
        const int i = idx / CB, c = idx % CB;
        invL[idx] = (c <= i) ? a[i * CD_LD + c] : 0.0;
    }
}

// M[i][i] += jitter for i < n; padded diagonal (i >= n) = 1
__global__ void chol_prepare_diag_kernel(double* M, int64_t ld, int64_t n, int64_t npad, double jitter) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) M[i * ld + i] += jitter;
    else if (i < npad) M[i * ld + i] = 1.0;
}

// ---- solves -------------------------------------------------------------------------------------------
// Substitution runs over SUPER blocks of CS = 4 diagonal blocks (512 rows): one CTA resolves the
// 512 x 512 triangle of a super block (diagonal blocks through their stored inverses, the six
// off-diagonal blocks as 128 x 128 products), then one wide kernel applies the super block to
// everything below (forward) or left of it (backward).  The dependency chain is 2 x n / 512
// launches instead of 2 x n / 128.
constexpr int CS = 4;
constexpr int SV_THREADS = 256;

// out (op)= op(M) v for a 128 x 128 block M, v in shared memory as [nrhs][CB]; (row-major, ld); trans = 0: M v, 1: M^T v.
// acc = 0: out = result, acc = 1: out -= result (out: global [CB][nrhs]).  All SV_THREADS threads call it.
__device__ __forceinline__ void sv_block_apply(const double* __restrict__ M, int64_t ld, int trans,
                                               const double* __restrict__ vs, double* __restrict__ out, int nrhs,
                                               int acc, double* __restrict__ red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!trans) {  // a warp per row, lanes read 4 adjacent columns: coalesced 1 KB per row
        constexpr int RPW = CB / (SV_THREADS / 32);  // rows per warp; all loads in flight before the math
        double4 mr[RPW];
#pragma unroll
        for (int i = 0; i < RPW; ++i)
            mr[i] = *reinterpret_cast<const double4*>(M + (int64_t)(warp + i * (SV_THREADS / 32)) * ld + 4 * lane);
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            const int row = warp + i * (SV_THREADS / 32);
            const double4 m = mr[i];
            for (int r = 0; r < nrhs; ++r) {
                const double4 v4 = *reinterpret_cast<const double4*>(vs + r * CB + 4 * lane);  // vs is [rhs][CB]
                double s = m.x * v4.x;
                s = pn_fma(m.y, v4.y, s);
                s = pn_fma(m.z, v4.z, s);
                s = pn_fma(m.w, v4.w, s);
#pragma unroll
                for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
                if (lane == 0) out[row * nrhs + r] = acc ? out[row * nrhs + r] - s : s;
            }
        }
    } else {  // a thread per column, the two halves of the rows on the two thread halves; one pass over
              // the block for all right-hand sides
        const int col = tid & (CB - 1), half = tid >> 7;
        double s[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) s[r] = 0.0;
#pragma unroll 8
        for (int l = half * (CB / 2); l < (half + 1) * (CB / 2); ++l) {
            const double m = M[(int64_t)l * ld + col];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < nrhs) s[r] = pn_fma(m, vs[r * CB + l], s[r]);
        }
        if (half) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < nrhs) red[r * CB + col] = s[r];
        }
        __syncthreads();
        if (!half) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < nrhs) {
                    const double t = s[r] + red[r * CB + col];
                    out[col * nrhs + r] = acc ? out[col * nrhs + r] - t : t;
                }
        }
        __syncthreads();
    }
}

// One CTA: the triangle of a super block.  fwd: for j = 0..nb-1: y_j = invL_j b_j; b_{j2} -= L[j2,j] y_j (j2 > j).
// bwd: for j = nb-1..0: x_j = invL_j^T y_j; y_{j2} -= L[j,j2]^T x_j (j2 < j).  Ld: the (sb,sb) diagonal
// super block of L (row-major, ld), invd: its nb inverse diagonal blocks, b: its rows of the rhs.
__global__ void __launch_bounds__(SV_THREADS) chol_super_diag_kernel(const double* __restrict__ Ld, int64_t ld,
                                                                      const double* __restrict__ invd,
                                                                      double* __restrict__ b, int nrhs, int nb, int bwd) {
    __shared__ double bs[CS * CB * 8];  // the super block's rows of the right-hand side live here throughout
    __shared__ __align__(32) double vs[CB * 8];  // [rhs][CB]
    __shared__ double red[CB * 8];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nb * CB * nrhs; idx += SV_THREADS) bs[idx] = b[idx];
    for (int step = 0; step < nb; ++step) {
        const int j = bwd ? nb - 1 - step : step;
        double* bj = bs + j * CB * nrhs;
        __syncthreads();
        for (int idx = tid; idx < CB * nrhs; idx += SV_THREADS) vs[(idx % nrhs) * CB + idx / nrhs] = bj[idx];
        __syncthreads();
        sv_block_apply(invd + (int64_t)j * CB * CB, CB, bwd, vs, bj, nrhs, 0, red);
        __syncthreads();
        for (int idx = tid; idx < CB * nrhs; idx += SV_THREADS) vs[(idx % nrhs) * CB + idx / nrhs] = bj[idx];  // the solved block
        __syncthreads();
        if (!bwd) {
            for (int j2 = j + 1; j2 < nb; ++j2)
                sv_block_apply(Ld + ((int64_t)j2 * CB) * ld + j * CB, ld, 0, vs, bs + j2 * CB * nrhs, nrhs, 1, red);
        } else {
            for (int j2 = 0; j2 < j; ++j2)
                sv_block_apply(Ld + ((int64_t)j * CB) * ld + j2 * CB, ld, 1, vs, bs + j2 * CB * nrhs, nrhs, 1, red);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < nb * CB * nrhs; idx += SV_THREADS) b[idx] = bs[idx];
}

// forward: b_rows -= L[rows, super block] y_sb for the rows below the super block.  One CTA per 8
// rows, a warp per row; width = columns of the super block (<= 512).
__global__ void __launch_bounds__(128) chol_fwd_update_kernel(const double* __restrict__ L, int64_t ld,
                                                               const double* __restrict__ ysb, double* __restrict__ b,
                                                               int64_t rows, int nrhs, int width) {
    __shared__ double ys[CS * CB * 8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < width * nrhs; idx += 128) ys[idx] = ysb[idx];
    __syncthreads();
    for (int rr = 0; rr < 2; ++rr) {
        const int64_t row = (int64_t)blockIdx.x * 8 + warp * 2 + rr;
        if (row >= rows) break;
        const double* lr = L + row * ld;
        double s[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) s[r] = 0.0;
        double4 lvs[CS];  // the whole row segment in flight before the math
#pragma unroll
        for (int c = 0; c < CS; ++c)
            if (c * CB < width) lvs[c] = *reinterpret_cast<const double4*>(lr + c * CB + 4 * lane);
#pragma unroll
        for (int c = 0; c < CS; ++c) {
            if (c * CB >= width) break;
            const int c0 = c * CB;
            const double4 lv = lvs[c];
            const double* yv = ys + (c0 + 4 * lane) * nrhs;
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < nrhs) {
                    s[r] = pn_fma(lv.x, yv[r], s[r]);
                    s[r] = pn_fma(lv.y, yv[nrhs + r], s[r]);
                    s[r] = pn_fma(lv.z, yv[2 * nrhs + r], s[r]);
                    s[r] = pn_fma(lv.w, yv[3 * nrhs + r], s[r]);
                }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < nrhs) {
                double t = s[r];
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
                if (lane == 0) b[row * nrhs + r] -= t;
            }
    }
}

// backward: y_cols -= L[super block rows, cols]^T x_sb for all columns left of the super block.  One
// CTA per 128 columns, a thread per column, coalesced reads of the `height` rows of the super block.
__global__ void __launch_bounds__(CB) chol_bwd_update_kernel(const double* __restrict__ Li, int64_t ld,
                                                              const double* __restrict__ xsb, double* __restrict__ y,
                                                              int nrhs, int height) {
    __shared__ double xs[CS * CB * 8];
    const int t = threadIdx.x;
    for (int idx = t; idx < height * nrhs; idx += CB) xs[idx] = xsb[idx];
    __syncthreads();
    const int64_t col = (int64_t)blockIdx.x * CB + t;
    double s[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) s[r] = 0.0;
#pragma unroll 8
    for (int l = 0; l < height; ++l) {
        const double v = Li[(int64_t)l * ld + col];
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < nrhs) s[r] = pn_fma(v, xs[l * nrhs + r], s[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
        if (r < nrhs) y[col * nrhs + r] -= s[r];
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
// of the diagonal blocks of L; *info_dev the 1-based index of the first bad pivot (0: success).
cudaError_t pn_chol_factor(double* M, int64_t n, int64_t npad, double jitter, double* panel, double* invd,
                           int* info_dev, cudaStream_t stream, cudaStream_t side, cudaEvent_t ev_panel,
                           cudaEvent_t ev_diag) {
    cudaError_t err;
    if ((err = cudaMemsetAsync(info_dev, 0, sizeof(int), stream)) != cudaSuccess) return err;
    chol_prepare_diag_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(M, npad, n, npad, jitter);
    ++g_pn_launches;
    const size_t smem_diag = sizeof(double) * (CB * CD_LD + 3 * CB + 3 * 32 * 33);
    const size_t smem_gemm = sizeof(double) * 2 * (CB + CTN) * CLD;
    if ((err = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_diag)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(chol_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gemm)) != cudaSuccess) return err;
    const int64_t T = npad / CB;
    // Look-ahead: the factorisation of diagonal block k + 1 (one CTA, latency-bound) runs on `side`
    // under the update of step k; it applies its own tile of that update itself, so it is runnable as
    // soon as the panel is and gets its SM before the update's CTAs occupy every slot.
    chol_diag_kernel<<<1, 256, smem_diag, stream>>>(M, npad, invd, info_dev, 0, nullptr);
    ++g_pn_launches;
    for (int64_t k = 0; k < T; ++k) {
        double* diag = M + k * CB * (npad + 1);
        const int64_t nb = T - k - 1;
        if (nb == 0) break;
        ChGemm p{};
        p.A = diag + CB * npad;  // block (k + 1, k)
        p.lda = npad;
        p.B = invd + k * CB * CB;
        p.ldb = CB;
        p.C = const_cast<double*>(p.A);
        p.ldc = npad;
        p.panel = panel;
        p.mode = 0;
        chol_gemm_kernel<<<(unsigned)(2 * nb), CG_THREADS, smem_gemm, stream>>>(p);
        ++g_pn_launches;
        if ((err = cudaEventRecord(ev_panel, stream)) != cudaSuccess) return err;
        // L[i,k] into the matrix (rows below the diagonal block of column k; nobody else touches them)
        if ((err = cudaMemcpy2DAsync(const_cast<double*>(p.A), (size_t)npad * 8, panel, (size_t)CB * 8, (size_t)CB * 8,
                                     (size_t)(nb * CB), cudaMemcpyDeviceToDevice, stream)) != cudaSuccess)
            return err;
        ChGemm u{};
        u.A = panel;
        u.lda = CB;
        u.B = panel;
        u.ldb = CB;
        u.C = diag + CB * (npad + 1);  // block (k + 1, k + 1)
        u.ldc = npad;
        u.panel = nullptr;
        u.mode = 1;
        // side stream (high priority, issued before the wide update): tile (k + 1, k + 1) of the update
        // fused into the next diagonal factorisation
        if ((err = cudaStreamWaitEvent(side, ev_panel, 0)) != cudaSuccess) return err;
        chol_diag_kernel<<<1, 256, smem_diag, side>>>(u.C, npad, invd + (k + 1) * CB * CB, info_dev, (int)((k + 1) * CB),
                                                      panel);
        ++g_pn_launches;
        if ((err = cudaEventRecord(ev_diag, side)) != cudaSuccess) return err;
        // main stream: the other tiles of the trailing update
        const int64_t tiles = nb * (nb + 1) / 2 - 1;
        if (tiles > 0) {
            u.tile0 = 2;
            chol_gemm_kernel<<<(unsigned)(2 * tiles), CG_THREADS, smem_gemm, stream>>>(u);
            ++g_pn_launches;
        }
        if ((err = cudaStreamWaitEvent(stream, ev_diag, 0)) != cudaSuccess) return err;
    }
    return cudaGetLastError();
}

// b: (npad, nrhs) row-major, nrhs <= 8, overwritten with the solution of L L^T x = b
cudaError_t pn_chol_solve(const double* M, int64_t npad, const double* invd, double* b, int nrhs, cudaStream_t stream) {
    if (nrhs < 1 || nrhs > 8) return cudaErrorInvalidValue;  // the kernels' shared arrays hold at most 8 right-hand sides
    const int64_t T = npad / CB;
    const int64_t SB = (T + CS - 1) / CS;
    for (int64_t sb = 0; sb < SB; ++sb) {  // L y = b
        const int64_t k0 = sb * CS;
        const int nb = (int)((T - k0 < CS) ? T - k0 : CS);
        double* bk = b + k0 * CB * nrhs;
        chol_super_diag_kernel<<<1, SV_THREADS, 0, stream>>>(M + k0 * CB * (npad + 1), npad, invd + k0 * CB * CB, bk, nrhs,
                                                             nb, 0);
        ++g_pn_launches;
        const int64_t rows = npad - (k0 + nb) * CB;
        if (rows > 0) {
            chol_fwd_update_kernel<<<(unsigned)((rows + 7) / 8), 128, 0, stream>>>(
                M + (k0 + nb) * CB * npad + k0 * CB, npad, bk, bk + (int64_t)nb * CB * nrhs, rows, nrhs, nb * CB);
            ++g_pn_launches;
        }
    }
    for (int64_t sb = SB - 1; sb >= 0; --sb) {  // L^T x = y
        const int64_t k0 = sb * CS;
        const int nb = (int)((T - k0 < CS) ? T - k0 : CS);
        double* bk = b + k0 * CB * nrhs;
        chol_super_diag_kernel<<<1, SV_THREADS, 0, stream>>>(M + k0 * CB * (npad + 1), npad, invd + k0 * CB * CB, bk, nrhs,
                                                             nb, 1);
        ++g_pn_launches;
        if (k0 > 0) {
            chol_bwd_update_kernel<<<(unsigned)k0, CB, 0, stream>>>(M + k0 * CB * npad, npad, bk, b, nrhs, nb * CB);
            ++g_pn_launches;
        }
    }
    return cudaGetLastError();
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

cudaError_t pn_chol_rowdots(const double* W, const double* G, int64_t ld, int64_t n, int64_t rows, const double* vecs,
                            int nvec, double* quad, double* dots, int64_t ldd, cudaStream_t stream) {
    if (rows <= 0) return cudaSuccess;
    chol_rowdots_kernel<<<(unsigned)rows, WR_THREADS, 0, stream>>>(W, G, ld, n, vecs, nvec, quad, dots, ldd);
    ++g_pn_launches;
    return cudaGetLastError();
}
