// Host-side interfaces between the C-ABI layer (pncov_api.cu) and the kernel launchers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pncov_common.cuh"

// point sets are padded (with zeros) to a multiple of this many points so that every
// tile is a full, 16-byte aligned TMA bulk copy
#define PN_PAD 512

struct PnGramLaunch {
    const void* x0_soa;  // (d, ld0) row tile source: coordinate-major copy of the row point set
    const void* x1_soa;  // (d, ld1) column point set
    int64_t ld0, ld1;
    int64_t n0, n1;      // logical Gram shape
    const void* scale;   // d per-dimension scale factors (element type = point dtype) or nullptr
    int d;
    void* out;           // (n0, n1) row-major
    int64_t ldo;
    int in_f32;          // points are float32
    int out_f32;         // Gram entries are float32 (only ExpQuad on float32 points)
    int symmetric;       // x1 is x0 and the whole square matrix is requested: mirror stores allowed
    int sm_count;        // multiProcessorCount of the handle's device (grid sizing)
    PnEpiParams epi;
};

struct PnMatvecLaunch {
    const void* x0_aos;  // (n0pad, d) row point set (rows of K)
    const void* x1_aos;  // (n1pad, d) column point set, zero padded
    const void* x1_soa;  // (d, ld1) coordinate-major copy of the same points (generic kernels)
    int64_t ld1;
    const void* x0_soa;  // (d, ld0) coordinate-major copy of the row points (row-staged kernel, d > 8)
    int64_t ld0;
    int64_t n0, n1, n1pad;
    const void* scale;   // as above
    int d;
    const void* v;       // (n1pad, kc) packed, zero padded rows, element type = point dtype
    void* y;             // (n0, kc) packed
    int kc;              // columns in this pass (1, 2, 4 or 8)
    void* partial;       // (jsplit, n0, kc) scratch when jsplit > 1
    int jsplit;
    // norm-identity / DMMA variant (ExpQuad, RatQuad; float64; d <= 8; host-checked error bound):
    // centred + scaled operands, 12 doubles per point: a12 = [xt(8), |xt|^2, 0, 0, 0] for the
    // rows, b12 = [-2 xt(8), |xt|^2, 0, 0, 0] for the columns.  epi then holds the epilogue
    // parameters in the scaled units.
    const void* a12;
    const void* b12;
    int use_mma;         // 0: direct differences, 1: norm identity with DFMA, 2: norm identity with DMMA
    // symmetric halving (x1 is x0, kc == 1): each off-diagonal tile is evaluated once and serves y_I and
    // y_J; row blocks part, part + nparts, ... only (multi-GPU).  The partial sums go to the integer
    // planes `bins` (see pn_bin_add in pncov_common.cuh; zero on entry), not to y
    int sym, sym_part, sym_nparts;
    long long* bins;      // [hi plane | lo plane] (n0 * kc int64 each) followed by the non-finite plane (n0 * kc doubles)
    int vstride;          // doubles per row of the packed right-hand sides v (kc, or kc + 2 for the symmetric multi-column kernel)
    const double* vmax;   // device scalar: max |v_j| over the original right-hand side
    // ExpQuad in factorised form (use_mma == 2): a12[.][9] / b12[.][9] hold exp(-|xt|^2); v arrives
    // multiplied by the column factors, y leaves to be multiplied by the row factors
    int fact;
    int in_f32;
    PnEpiParams epi;
};

struct PnPairsLaunch {
    const void* x0;  // (b, d) row-major
    const void* x1;
    int64_t b;
    const void* scale;
    int d;
    void* out;
    int in_f32, out_f32;
    PnEpiParams epi;
};

// non-isotropic families and covariance arithmetic (pncov_extra.cu)
#define PN_PM_LD 12  // per-dimension ProductMatern parameters: {scale, p, c_0 .. c_8, pad}
struct PnExtraLaunch {
    int mode;            // 0: <x0, x1> (Linear / Polynomial), 1: equality (WhiteNoise), 2: ProductMatern
    const void* x0_soa;  // Gram: (d, ld0) coordinate-major; pairs: (B, d) row-major
    const void* x1_soa;
    int64_t ld0, ld1;
    int64_t n0, n1;      // pairs: n0 = B
    int d;
    const void* pm;      // ProductMatern: (d, PN_PM_LD) device array
    double constant;     // Linear / Polynomial: c; WhiteNoise: sigma^2
    int exponent;        // Polynomial: q >= 0 (Linear: 1)
    void* out;
    int64_t ldo;
    int vec_ok;          // filled by the launcher
};
cudaError_t pn_launch_gram_extra(const PnExtraLaunch& g, cudaStream_t stream);
// Linear / Polynomial for wide inputs: zero-padded row-major operands and the DMMA GEMM on them
cudaError_t pn_launch_pad_rows(const double* aos, int64_t n, int d, int64_t npad, int dpad, double* outp, cudaStream_t stream);
cudaError_t pn_launch_gram_dot_mma(const double* A, const double* B, int dpad, int64_t n0, int64_t n1, double* out,
                                   int64_t ldo, double constant, int exponent, cudaStream_t stream);
cudaError_t pn_launch_pairs_extra(const PnExtraLaunch& g, cudaStream_t stream);
cudaError_t pn_launch_combine(double* dst, const double* src, int64_t n, int op, double alpha, double beta,
                              cudaStream_t stream);

// update chains of the matrix-based solver (pncov_extra.cu)
cudaError_t pn_launch_chain_apply(const double* U, const double* V, int64_t ld, int64_t n, int k, int mode, const double* x,
                                  double* y, double* partial, int splits, cudaStream_t stream);

// FP64 Cholesky (pncov_chol.cu)
cudaError_t pn_chol_factor(double* M, int64_t n, int64_t npad, double jitter, double* panel, double* invd,
                           int* info_dev, int* scratch, cudaStream_t stream, cudaStream_t side, cudaEvent_t ev_panel,
                           cudaEvent_t ev_update);
cudaError_t pn_chol_solve(const double* M, int64_t npad, const double* invd, double* b, int64_t ldb, int nrhs, int* sync,
                          cudaStream_t stream);
// many right-hand sides, one per ROW of Z (mpad x npad); Y: scratch of the same shape
cudaError_t pn_chol_solve_wide(const double* M, int64_t npad, const double* invd, double* Z, double* Y, int64_t mpad,
                               int forward_only, cudaStream_t stream, cudaStream_t side, cudaEvent_t ev_main,
                               cudaEvent_t ev_side);
cudaError_t pn_chol_mirror_upper(double* M, int64_t ld, int64_t n, cudaStream_t stream);
cudaError_t pn_chol_transpose(const double* src, int64_t lds, int64_t rows, int64_t cols, double* dst, int64_t ldd,
                              cudaStream_t stream);
cudaError_t pn_chol_rowdots(const double* W, const double* G, int64_t ld, int64_t n, int64_t rows, const double* vecs,
                            int nvec, double* quad, double* dots, int64_t ldd, cudaStream_t stream);

// ExpQuad / RatQuad Gram for wide inputs (d > 16): guarded norm identity on FP64 tensor cores
// (pncov_gram_mma.cu).  All row-side pointers are already offset to the first row of the launch.
struct PnGramMmaLaunch {
    const double* A;       // (rows padded to 128, dpad) centred + scaled row points xt
    const double* B;       // (columns padded to 64, dpad) -2 xt of the column points
    const double* nrmA;    // |xt|^2 per row point
    const double* nrmB;
    const int* thrA;       // guard threshold of every point (high word, see the kernel's header)
    const int* thrB;
    const double* x0_aos;  // original points (n, d) row-major: direct-difference redo
    const double* x1_aos;
    const double* scale;   // d per-dimension scale factors of the direct form or nullptr
    int d, dpad;
    int64_t n0, n1;        // rows / columns of this launch
    int64_t row0;          // index of the first row within its point set (diagonal detection)
    int same_points;       // rows and columns are the same point set: entries (i, i) are redone (exact 1)
    double* out;
    int64_t ldo;
    int vec_ok;            // filled by the launcher
    int mirror;            // filled by the launcher: > 0 (the number of column tiles) for a whole symmetric matrix,
                           // whose tiles below the diagonal blocks are written as mirrored copies
    int k4_last;           // filled by the launcher: DMMA k-steps in the last 16-coordinate chunk (1 .. 4)
    unsigned v_hi;         // fast epilogue valid while hi(v) < v_hi (v = S for ExpQuad, 1 + c S for RatQuad)
    PnEpiParams epi;       // epilogue in the units of the operands
    PnEpiParams epi_direct;  // epilogue of the direct form (original units)
};
cudaError_t pn_launch_gram_mma_centre(const double* soa0, int64_t ld0, const double* soa1, int64_t ld1, double ntot, int d,
                                      double* centre, cudaStream_t stream);
cudaError_t pn_launch_gram_mma_operands(const double* aos, int64_t n, int64_t rows_alloc, int d, int dpad, const double* centre,
                                        const double* scale, double iso_scale, double gam_over_tol, double dup_factor,
                                        int expquad, double* A, double* B, double* nrm, int* thr, cudaStream_t stream);
cudaError_t pn_launch_gram_mma(const PnGramMmaLaunch& p, cudaStream_t stream);

cudaError_t pn_launch_gram(const PnGramLaunch& g, cudaStream_t stream);
// ProductMatern of uniform degree, whole symmetric Gram: the 64-grid interior through the mirrored-store kernel; *done =
// its extent (0: not applicable), the edges are the caller's
cudaError_t pn_launch_gram_pm_sym(const PnGramLaunch& g, const double* pm_host, cudaStream_t stream, int64_t* done);
// covariance arithmetic (Scaled / Sum / Product of isotropic leaves) in one launch; g.epi is ignored
cudaError_t pn_launch_gram_fused(const PnGramLaunch& g, const PnFusedProgram& F, cudaStream_t stream);
// pn_plan_matvec fills m.jsplit (how many column-range splits the launch will use, which
// sizes the `partial` scratch); pn_launch_matvec then runs with that plan
cudaError_t pn_plan_matvec(PnMatvecLaunch& m, int sm_count);
cudaError_t pn_launch_matvec(const PnMatvecLaunch& m, int sm_count, cudaStream_t stream);
cudaError_t pn_launch_pairs(const PnPairsLaunch& p, cudaStream_t stream);
// symmetric-halving products: max |v| into *vmax (device), and the integer planes back to doubles
// (fac: operand rows holding the factor of the factorised ExpQuad form at [i * fac_ld + fac_idx], or nullptr)
cudaError_t pn_launch_vmax(const double* v, int64_t ldv, int64_t n, int k, double* vmax, cudaStream_t stream);
cudaError_t pn_launch_unbin(const long long* bins, int64_t n0, int kc, int64_t n1, const double* vmax, const double* fac,
                            int fac_ld, int fac_idx, double* y, cudaStream_t stream);

// pncov_api.cu
extern int64_t g_pn_launches;
// multiProcessorCount of the CURRENT device (queried once per device): grids are sized from it
int pn_sm_count();

// host worker threads (pncov_host.cu): mirror image of the upper blocks of a symmetric host matrix
int pn_host_threads();
void pn_host_mirror_async(void* out, int64_t ld, int64_t n, int64_t r0, int64_t r1, int elem);
void pn_host_mirror_wait();
void pn_host_touch(void* p, size_t bytes);
