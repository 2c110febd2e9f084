/* pncov -- C ABI of the B200 covariance-function library (libpncov.so).
 *
 * This is the drop-in boundary for the one data-parallel hot path of ProbNum,
 * probnum.randprocs.covfuncs: the dense Gram matrix K(X0, X1) and the matrix-free
 * product K(X0, X1) @ V for the isotropic kernels ExpQuad, Matern (half-integer nu)
 * and RatQuad.  Plain C types only; no torch / numpy types cross this boundary.
 * Every entry point names the reference interface it replaces (paths relative to
 * /root/reference/src/probnum).
 *
 * Conventions
 *   - all functions return 0 on success, a negative PNCOV_E* code for invalid
 *     arguments and a positive code for CUDA runtime failures;
 *     pncov_last_error() returns a thread-local message for the last failure.
 *   - point sets are C-contiguous row-major (n, d) arrays of `dtype`
 *     (scalar-input kernels use d = 1), exactly what
 *     CovarianceFunction._preprocess_linop_input produces
 *     (randprocs/covfuncs/_covariance_function.py:676-693).
 *   - the caller owns every host buffer; the library keeps no host pointer after a
 *     call returns.  Device buffers, streams and scratch belong to the handle.
 *   - a handle is bound to one CUDA device and is not re-entrant.
 *   - there is no CPU fallback anywhere: without a usable device every compute
 *     entry point fails.
 */
#ifndef PNCOV_H_
#define PNCOV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNCOV_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define PNCOV_API __attribute__((visibility("default")))
#else
#define PNCOV_API
#endif

/* kernel families (randprocs/covfuncs/_exponentiated_quadratic.py, _matern.py,
 * _rational_quadratic.py) */
#define PNCOV_EXPQUAD 0
#define PNCOV_MATERN 1 /* nu = p + 1/2, p = 0..PNCOV_MAX_MATERN_P */
#define PNCOV_RATQUAD 2

/* Non-isotropic families, float64 points, dense Gram and elementwise evaluation
 * only (pncov_matmul returns PNCOV_ENOTIMPL for them).  They reuse pncov_create's arguments:
 *   PNCOV_LINEAR          k = <x0, x1> + c            alpha = c                 (_linear.py:44-50)
 *   PNCOV_POLYNOMIAL      k = (<x0, x1> + c)^q        alpha = c, nu = q in {0..64}  (_polynomial.py:50-56)
 *   PNCOV_WHITENOISE      k = sigma^2 [x0 == x1]      alpha = sigma^2           (_white_noise.py:38-55)
 *   PNCOV_PRODUCT_MATERN  k = prod_m Matern_{nu_m}(|x0_m - x1_m| / l_m)
 *                         lengthscales = [l_0 .. l_{d-1}, nu_0 .. nu_{d-1}], n_lengthscales = 2 d
 *                         (_product_matern.py:60-122)
 * lengthscales may be NULL (n_lengthscales 0) for the first three. */
#define PNCOV_LINEAR 3
#define PNCOV_POLYNOMIAL 4
#define PNCOV_WHITENOISE 5
#define PNCOV_PRODUCT_MATERN 6

#define PNCOV_F64 0
#define PNCOV_F32 1

#define PNCOV_MAX_MATERN_P 8

/* error codes */
#define PNCOV_EINVAL (-1)   /* invalid argument (shape, kind, parameter) */
#define PNCOV_ENOTIMPL (-2) /* outside the accelerated scope (e.g. non-half-integer nu) */
#define PNCOV_ENOMEM (-3)
#define PNCOV_ENODEVICE (-4)

typedef struct pncov_handle pncov_handle;

/* library / device ---------------------------------------------------------------- */
PNCOV_API int pncov_version(void);
PNCOV_API const char* pncov_last_error(void);
PNCOV_API int pncov_device_count(int* count);

/* pinned host memory for result arrays (so D2H runs at PCIe rate) */
PNCOV_API int pncov_host_alloc(void** ptr, int64_t nbytes);
PNCOV_API int pncov_host_free(void* ptr);

/* operator life cycle ---------------------------------------------------------------
 * pncov_create binds a covariance function and its two point sets to a device:
 * it replaces the construction of CovarianceLinearOperator(x0, x1, ...)
 * (randprocs/covfuncs/_covariance_linear_operator.py:50-65) and the argument
 * handling of CovarianceFunction.matrix / .linop (_covariance_function.py:359-488).
 *
 *   kind, dtype         PNCOV_* constants above
 *   d                   input dimension (>= 1)
 *   lengthscales        n_lengthscales = 1 (isotropic) or d (ARD) positive values;
 *                       RatQuad accepts 1 only (_rational_quadratic.py:59-69)
 *   nu                  Matern smoothness, must be p + 1/2 (_matern.py:136-151)
 *   alpha               RatQuad scale mixture > 0 (_rational_quadratic.py:66-68)
 *   x0, n0              first point set, n0 rows
 *   x1, n1              second point set or NULL (then x1 = x0, n1 ignored)
 *   points_on_device    0: x0/x1 are host pointers; 1: device pointers on `device`
 */
PNCOV_API int pncov_create(pncov_handle** out, int device, int kind, int dtype, int64_t d,
                 const double* lengthscales, int64_t n_lengthscales, double nu, double alpha,
                 const void* x0, int64_t n0, const void* x1, int64_t n1,
                 int points_on_device);
PNCOV_API int pncov_destroy(pncov_handle* h);

PNCOV_API int pncov_shape(const pncov_handle* h, int64_t* n0, int64_t* n1);
/* dtype of Gram entries: follows the reference under NumPy 2 (f64 -> f64;
 * f32 -> f32 for ExpQuad, f64 for Matern and RatQuad) */
PNCOV_API int pncov_gram_dtype(const pncov_handle* h, int* dtype);

/* dense Gram matrix -----------------------------------------------------------------
 * replaces CovarianceFunction._evaluate_matrix (_covariance_function.py:553-582)
 * for scalar-output isotropic kernels, i.e. IsotropicMixin._squared_euclidean_distances
 * (:757-781) fused with ExpQuad/Matern/RatQuad._evaluate.
 *   out    (n0, n1) row-major, leading dimension ld (elements), dtype pncov_gram_dtype
 *   transpose != 0 writes K^T, shape (n1, n0)  (CovarianceLinearOperator._transpose,
 *   _covariance_linear_operator.py:83-90)
 * pncov_gram:        `out` is host memory (pageable or pinned); blocks until done.
 * pncov_gram_device: `out` is device memory on the handle's device; asynchronous on
 *                    `stream` (a cudaStream_t passed through unchanged; 0 is the legacy
 *                    default stream, which is what torch's default stream is).  The
 *                    handle's scratch is reused across calls, so issue all *_device calls
 *                    of one handle on the same stream.
 */
PNCOV_API int pncov_gram(pncov_handle* h, void* out, int64_t ld, int transpose);
PNCOV_API int pncov_gram_device(pncov_handle* h, void* out, int64_t ld, int transpose, void* stream);

/* matrix-free product Y = K(X0, X1) @ V  (or K^T @ V) ---------------------------------
 * replaces CovarianceLinearOperator._matmul (_covariance_linear_operator.py:78-81),
 * i.e. both `keops_lazy_tensor @ x` and `self.todense() @ x`; K is never stored.
 *   V   (n1, k) row-major with leading dimension ldv; Y (n0, k) with ldy
 *       (transpose: V is (n0, k), Y is (n1, k)); dtype = the handle's dtype
 */
PNCOV_API int pncov_matmul(pncov_handle* h, const void* v, int64_t k, int64_t ldv, void* y, int64_t ldy,
                 int transpose);
PNCOV_API int pncov_matmul_device(pncov_handle* h, const void* v, int64_t k, int64_t ldv, void* y,
                        int64_t ldy, int transpose, void* stream);

/* Symmetric operators (x1 == NULL at creation), one right-hand side, float64, n >= 2048:
 * pncov_matmul* evaluate every off-diagonal tile of K once and use it for both y_I and y_J (half the
 * exp / DMMA work).  The partial sums of the CTAs are accumulated order-independently: each is split
 * exactly into two integers on a per-component power-of-two grid and added with integer atomics, so the
 * product is bit-for-bit reproducible from run to run (PNCOV_MATVEC_SYM=off in the environment still
 * selects the kernel that evaluates all N^2 pairs).
 * pncov_matmul_sym_part_device is the multi-GPU form: caller `part` of `nparts` (one per GPU, each
 * with a handle over the SAME full point set) gets the contribution of row blocks part,
 * part + nparts, ... in the full-length vector y (n0 doubles, device memory, overwritten); the sum
 * of all parts (an all-reduce) is K @ v.  This is the reduce-scatter form of the row-sharded
 * product + all-gather named in SURVEY.md 8e.
 * pncov_matmul_sym_part_bins_device leaves the part's integer planes instead (3 n0 words of 8 bytes:
 * two int64 planes and one float64 plane holding non-finite contributions only): summing the first
 * 2 n0 words as int64 and the last n0 as float64 over the parts and passing the sums to
 * pncov_matmul_sym_finish_device on any of the handles gives K @ v with the SAME bits for every
 * nparts, including 1 (integer addition is associative).
 * All return PNCOV_ENOTIMPL when the symmetric kernel does not apply to the handle. */
PNCOV_API int pncov_matmul_sym_part_device(pncov_handle* h, const void* v, void* y, int part, int nparts,
                                           void* stream);
PNCOV_API int pncov_matmul_sym_part_bins_device(pncov_handle* h, const void* v, void* bins, int part, int nparts,
                                                void* stream);
PNCOV_API int pncov_matmul_sym_finish_device(pncov_handle* h, const void* bins, void* y, void* stream);
/* The same for 1 .. 8 right-hand sides at once (v: (n, k) with leading dimension ldv, device memory).  For k > 1 the
 * planes hold 3 * 8 * n0 words (eight columns per pass, unused ones zero) and the kernel is the tensor-core form
 * (d <= 8, admissible norm identity); k = 1 is the call above.  PNCOV_ENOTIMPL when that kernel does not apply. */
PNCOV_API int pncov_matmulk_sym_part_bins_device(pncov_handle* h, const void* v, int64_t k, int64_t ldv, void* bins,
                                                 int part, int nparts, void* stream);
PNCOV_API int pncov_matmulk_sym_finish_device(pncov_handle* h, const void* bins, int64_t k, void* y, int64_t ldy,
                                              void* stream);

/* Per-handle choice of the product kernels (value: 1 on, 0 off, -1 default = automatic, subject to the
 * PNCOV_MATVEC_IDENT / _SYM / _FACT environment switches).  Switching all three off selects the most
 * conservative kernel -- direct differences like the reference, fixed summation order -- which is the
 * analogue of the reference's `_use_keops = False`
 * (tests/test_randprocs/test_covfuncs/test_linop_matrix.py:38-43). */
#define PNCOV_OPT_IDENTITY 0   /* norm identity |x|^2 + |y|^2 - 2 x.y on DMMA (ExpQuad / RatQuad) */
#define PNCOV_OPT_SYMMETRIC 1  /* symmetric halving (each off-diagonal tile evaluated once; reproducible) */
#define PNCOV_OPT_FACTORISED 2 /* ExpQuad as exp(-|x|^2) exp(-|y|^2) exp(2 x.y) */
PNCOV_API int pncov_set_option(pncov_handle* h, int option, int value);

/* covariance arithmetic in ONE launch ----------------------------------------------------
 * replaces ScaledCovarianceFunction / SumCovarianceFunction / ProductCovarianceFunction._evaluate
 * (/root/reference/src/probnum/randprocs/covfuncs/_arithmetic_fallbacks.py:55-56, :101-104, :163-166)
 * for expressions whose leaves are isotropic ExpQuad / Matern / RatQuad kernels and WhiteNoise on the
 * same points: the leaves share the squared distance S_ij, so the whole expression is evaluated per
 * entry and written once (8 bytes per entry instead of one write per operand plus 24 bytes per
 * combination).  After this call every dense-Gram entry point of the handle (pncov_gram*,
 * pncov_chol_factor_gram, pncov_chol_solve_gram) evaluates the program; the matrix-free product
 * returns PNCOV_ENOTIMPL.  The handle must hold float64 points, d <= 16, and have been created with
 * an isotropic kernel (its own kernel is ignored from here on).
 *   leaves (n_leaves <= 6, and at most 6 leaf evaluations in the steps): kinds[l] with lengthscales[l], nus[l] (Matern), alphas[l] (RatQuad alpha;
 *     WhiteNoise sigma^2);
 *   steps (n_steps <= 12; between two leaf evaluations [SCALEVAL] [SET | LIN | MUL] [SCALEACC], each at most once and
 *   in this order, else PNCOV_ENOTIMPL) over two values per entry, `acc` and `val`:
 *     PNCOV_F_LOAD     val = leaf[leaf_index](S)      PNCOV_F_MULLEAF  val = val * leaf[leaf_index](S)
 *     PNCOV_F_SCALEVAL val = val * a                  PNCOV_F_SET      acc = val
 *     PNCOV_F_LIN      acc = fma(b, val, a * acc)     PNCOV_F_MUL      acc = acc * val
 *     PNCOV_F_SCALEACC acc = acc * a
 *   -- the arithmetic of pncov_combine_device, so the result is bit-identical to building the
 *   operands one by one.  n_leaves = n_steps = 0 removes the program. */
#define PNCOV_F_LOAD 0
#define PNCOV_F_MULLEAF 1
#define PNCOV_F_SCALEVAL 2
#define PNCOV_F_SET 3
#define PNCOV_F_LIN 4
#define PNCOV_F_MUL 5
#define PNCOV_F_SCALEACC 6
PNCOV_API int pncov_set_program(pncov_handle* h, int n_leaves, const int* kinds, const double* lengthscales,
                                const double* nus, const double* alphas, int n_steps, const int* ops,
                                const int* leaf_index, const double* a, const double* b);

/* elementwise evaluation out[b] = k(x0[b], x1[b]) -------------------------------------
 * replaces the broadcasting path of CovarianceFunction.__call__
 * (_covariance_function.py:280-357) after the caller has broadcast both arguments
 * to a common batch of B points.  Host pointers; out dtype as pncov_gram_dtype.
 */
PNCOV_API int pncov_pairs(int device, int kind, int dtype, int64_t d, const double* lengthscales,
                int64_t n_lengthscales, double nu, double alpha, const void* x0, const void* x1,
                int64_t b, void* out);

/* covariance arithmetic on device-resident Gram matrices --------------------------------
 * replaces ScaledCovarianceFunction / SumCovarianceFunction / ProductCovarianceFunction._evaluate
 * (randprocs/covfuncs/_arithmetic_fallbacks.py:55-56, :101-104, :163-166), which combine full
 * host matrices with NumPy: the operands are built with pncov_gram_device into buffers from
 * pncov_device_alloc, combined in place, and only the result crosses PCIe.
 *   pncov_combine_device   n float64 elements, 16-byte aligned buffers on `device`, asynchronous on `stream`
 *       PNCOV_COMBINE_LINEAR    dst = alpha * dst + beta * src   (src may be NULL when beta == 0)
 *       PNCOV_COMBINE_MULTIPLY  dst = dst * src
 *   pncov_copy_to_host     device -> host copy of the result, ordered after `stream`, blocking
 */
#define PNCOV_COMBINE_LINEAR 0
#define PNCOV_COMBINE_MULTIPLY 1
PNCOV_API int pncov_device_alloc(int device, void** ptr, int64_t nbytes);
PNCOV_API int pncov_device_free(int device, void* ptr);
PNCOV_API int pncov_copy_to_host(int device, void* dst_host, const void* src_device, int64_t nbytes, void* stream);
PNCOV_API int pncov_combine_device(int device, void* dst, const void* src, int64_t n, int op, double alpha,
                                   double beta, void* stream);

/* Cholesky factor and solves on the device ------------------------------
 * replaces BQBeliefUpdate.compute_gram_cho_factor / gram_cho_solve
 * (quad/solvers/belief_updates/_belief_update.py:61-103), i.e. scipy.linalg.cho_factor of
 * `gram + jitter * I` and scipy.linalg.cho_solve with that factor.  float64, hand-written blocked
 * right-looking factorisation with the trailing updates on FP64 DMMA (csrc/pncov_chol.cu).
 *   pncov_chol_create       handle for matrices of order n on `device` (owns the n x n device copy)
 *   pncov_chol_factor       a: (n, n) row-major symmetric matrix, leading dimension ld, host memory
 *                           (a_on_device = 0) or device memory on the handle's device (= 1; the call
 *                           waits for all work already queued on that device first); only the lower
 *                           triangle is read.  a_on_device = 2: host memory of which only the UPPER
 *                           triangle is read -- what scipy.linalg.cho_factor(a) (lower=False, the call of
 *                           _belief_update.py:79) uses; half the host-to-device bytes, mirrored on the device.
 *                           *info = 0 on success, else the 1-based index of the
 *                           first non-positive pivot (scipy raises LinAlgError there); blocking.
 *   pncov_chol_factor_gram  same, but the matrix is the Gram matrix of a (square, float64) operator
 *                           handle: it is evaluated straight into the factorisation buffer and never
 *                           visits the host ("keep gram device-resident")
 *   pncov_chol_solve        x = (L L^T)^-1 b for b, x: (n, nrhs) row-major HOST arrays (nrhs <= 12: chained
 *                           substitution, one launch per direction for every three columns; from 13:
 *                           transposed on the device, blocked DMMA substitution)
 *   pncov_chol_solve_gram    the prediction path of
 *                           BQStandardBeliefUpdate.predict_integrand
 *                           (quad/solvers/belief_updates/_standard_update.py:131-151) and
 *                           IntegralVarianceReduction.__call__
 *                           (quad/solvers/acquisition_functions/_integral_variance_reduction.py:60-75):
 *                           with G = K(nodes, x), the (n, m) Gram matrix of operator h (x0 = the n
 *                           factorised nodes, x1 = the m prediction points) and W = (L L^T)^-1 G,
 *                             quad[j]           = sum_i W[i,j] G[i,j]             (np.sum(weights * kXx, axis=0))
 *                             dots[t * ldd + j] = sum_i vecs[t * ldv + i] W[i,j]  (weights.T @ fun_evals,
 *                                                                                 np.dot(kernel_means, weights))
 *                           for nvec <= 4 host vectors of length n.  G and W (n x m each) never leave
 *                           the device; *ms (may be NULL) = device time of the whole pass.
 *                           mode PNCOV_SOLVE_GRAM_REFERENCE_ORDER: both substitutions, then the sums
 *                           above, as the reference computes them.  PNCOV_SOLVE_GRAM_FORWARD_ONLY:
 *                           the same quantities from Y = L^-1 G alone, quad[j] = |Y[:,j]|^2 and
 *                           dots[t,j] = (L^-1 vecs[t]) . Y[:,j] -- half the flops, sums of squares
 *                           instead of cancelling sums, error ~ sqrt(cond) eps instead of cond eps.
 *   pncov_chol_get_factor   lower-triangular factor to a host array (the strict upper triangle of
 *                           the output is unspecified, as with scipy.linalg.cho_factor)
 *   pncov_chol_last_factor_ms   device time of the last factorisation (CUDA events)
 */
typedef struct pncov_chol pncov_chol;
PNCOV_API int pncov_chol_create(pncov_chol** out, int device, int64_t n);
PNCOV_API int pncov_chol_destroy(pncov_chol* c);
PNCOV_API int pncov_chol_factor(pncov_chol* c, const double* a, int64_t ld, double jitter, int a_on_device, int* info);
PNCOV_API int pncov_chol_factor_gram(pncov_chol* c, pncov_handle* h, double jitter, int* info);
PNCOV_API int pncov_chol_solve(pncov_chol* c, const double* b, int64_t nrhs, int64_t ldb, double* x, int64_t ldx);
#define PNCOV_SOLVE_GRAM_REFERENCE_ORDER 0
#define PNCOV_SOLVE_GRAM_FORWARD_ONLY 1
PNCOV_API int pncov_chol_solve_gram(pncov_chol* c, pncov_handle* h, int64_t nvec, const double* vecs, int64_t ldv,
                                    double* dots, int64_t ldd, double* quad, int mode, float* ms);
PNCOV_API int pncov_chol_get_factor(pncov_chol* c, double* l, int64_t ld);
PNCOV_API int pncov_chol_last_factor_ms(pncov_chol* c, float* ms);

/* Update chains of the matrix-based linear solver (SURVEY.md 8f N2 ii) ---------------------------
 * replaces the lazy sums of LambdaLinearOperators that SymmetricMatrixBasedSolver builds with
 * _mean_update (u v^T + v u^T, linalg/solvers/matrixbased.py:751-762) and _covariance_update
 * (Ws u^T, :764-775): the vectors live as rows of two device-resident (capacity, n) matrices U and
 * V, and a product with the first k of them is two kernel launches instead of k Python lambdas.
 *   pncov_chain_create   room for `capacity` (<= 12800) vector pairs of length n on `device`
 *   pncov_chain_set      vector pair number `index` from host memory (blocking)
 *   pncov_chain_apply    y = C_k x for the ncols columns of x, (n, ncols) row-major host arrays, with
 *                        mode 0: C_k = sum_{i<k} U_i V_i^T + V_i U_i^T   (mean update chain)
 *                        mode 1: C_k = sum_{i<k} V_i U_i^T               (covariance update chain, V = Ws)
 *                        mode 2: C_k = sum_{i<k} U_i V_i^T               (its transpose)
 */
typedef struct pncov_chain pncov_chain;
PNCOV_API int pncov_chain_create(pncov_chain** out, int device, int64_t n, int64_t capacity);
PNCOV_API int pncov_chain_destroy(pncov_chain* c);
PNCOV_API int pncov_chain_set(pncov_chain* c, int64_t index, const double* u, const double* v);
PNCOV_API int pncov_chain_apply(pncov_chain* c, int64_t k, int mode, const double* x, int64_t ncols, int64_t ldx,
                                double* y, int64_t ldy);

/* diagnostics ------------------------------------------------------------------------
 * number of kernel launches issued by this library in the calling process, and the
 * device time in ms of the last pncov_gram / pncov_matmul call of this handle (CUDA
 * events on the handle's stream around the kernels, excluding host<->device copies). */
/* one operator over several GPUs of ONE process -------------------------------------------------
 * "device_ids[], n_dev" in the ABI, so that a caller that binds this library without torch
 * or a process launcher (the reference's FFI: ctypes) can shard.  The rows of x0 are dealt contiguously to the
 * devices (host pointers, same argument meaning as pncov_create; a device id may repeat).
 *   pncov_sharded_gram    every device builds its row block and copies it into its slice of the host array
 *                         (CovarianceFunction.matrix, _covariance_function.py:359-413); no exchange between devices
 *   pncov_sharded_matmul  K @ V (CovarianceLinearOperator._matmul, _covariance_linear_operator.py:78-81), host
 *                         operands.  Symmetric operators (x1 == NULL) with one right-hand side use the
 *                         symmetric-halving kernel: device i takes the row blocks I = i (mod n_dev) of the triangle
 *                         into exact integer planes, which are added on the host -- the result has the bits a
 *                         single device gives.  Otherwise every device multiplies its row block.
 * Devices exchange nothing with each other on this path (results are assembled in host memory), so the library
 * needs no communicator; device-resident multi-process operands use torch.distributed / NCCL above the ABI
 * (probnum_b200/sharded.py) with pncov_matmul_sym_part_bins_device. */
typedef struct pncov_sharded pncov_sharded;
PNCOV_API int pncov_sharded_create(pncov_sharded** out, const int* device_ids, int n_dev, int kind, int dtype, int64_t d,
                                   const double* lengthscales, int64_t n_lengthscales, double nu, double alpha,
                                   const void* x0, int64_t n0, const void* x1, int64_t n1);
PNCOV_API int pncov_sharded_gram(pncov_sharded* s, void* out, int64_t ld);
PNCOV_API int pncov_sharded_matmul(pncov_sharded* s, const void* v, int64_t k, int64_t ldv, void* y, int64_t ldy);
PNCOV_API int pncov_sharded_destroy(pncov_sharded* s);

PNCOV_API int64_t pncov_launch_count(void);

/* Bytes pncov_gram has copied device-to-host in this process.  A whole symmetric result (x1 == NULL) may cross
 * PCIe as its blocks on and above the diagonal only, host threads of the library writing the rest as mirror
 * images; the library measures both ways on the process's first large symmetric results and keeps the faster
 * (PNCOV_GRAM_MIRROR=on|off in the environment forces one). */
PNCOV_API int64_t pncov_d2h_bytes(void);
/* which kernel family serves pncov_matmul for this handle: 0 = not decided yet (no product
 * so far), 1 = norm-identity + DMMA (ExpQuad / RatQuad, d <= 8, error bound satisfied),
 * 2 = direct differences, 3 = norm-identity + DMMA with symmetric halving, 4 = direct differences
 * with symmetric halving */
PNCOV_API int pncov_matmul_path(const pncov_handle* h, int* path);
PNCOV_API int pncov_last_kernel_ms(const pncov_handle* h, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* PNCOV_H_ */
