// Dense Gram matrix K[i, j] = k(x0_i, x1_j) for the isotropic kernels.
//
// Replaces, fused into one pass over the output,
//   IsotropicMixin._squared_euclidean_distances  randprocs/covfuncs/_covariance_function.py:757-781
//   ExpQuad/Matern/RatQuad._evaluate             _exponentiated_quadratic.py:89-100, _matern.py:175-193,
//                                                _rational_quadratic.py:71-84
//   CovarianceFunction._evaluate_matrix          _covariance_function.py:553-582
// (all paths relative to /root/reference/src/probnum).
//
// The kernel is bound by the N0 x N1 output write (HBM) as long as the FP64 work per
// entry stays under ~22 instructions; see DESIGN.md for the budget.  Layout:
//
//   * a CTA of 256 threads owns a strip of 512 output columns; each thread owns two
//     adjacent columns, so one row of the strip leaves the CTA as 256 16-byte
//     streaming stores (512 contiguous bytes per warp instruction);
//   * the strip's column points (coordinate-major, d x 512) are staged ONCE into shared
//     memory by 1-D TMA bulk copies; row points arrive in tiles of 64 rows through a
//     two-stage TMA + mbarrier pipeline, and are read back as warp-wide broadcasts;
//   * distances are direct differences ((x0 - x1) * s)^2 like the reference, so there
//     is no cancellation problem for near-duplicate or un-centred points.
#include "pncov_gram_tile.cuh"

namespace {

// ---- strip kernel (d <= G_DMAX) ----------------------------------------------------------
template <typename T, typename TOut, bool ARD, typename PROG>
__global__ void __launch_bounds__(G_THREADS)
    gram_strip_kernel(const __grid_constant__ PnGramLaunch g, int tiles_per_cta, int vec_ok, const __grid_constant__ PROG F) {
    constexpr bool FUSED = !std::is_same<PROG, PnNoProgram>::value;
    extern __shared__ __align__(128) unsigned char smem[];
    using P2 = typename Pair<T>::type;

    const int tid = threadIdx.x;
    const int d = g.d;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [0] x1 strip, [1],[2] x0 stages
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    T* sc = reinterpret_cast<T*>(smem + G_HDR_BYTES);
    T* x1s = reinterpret_cast<T*>(smem + G_HDR_BYTES + 128);
    T* x0s = x1s + (size_t)d * G_TN;  // [2][d][G_TM]

    const int64_t n_tiles = (g.n0 + G_TM - 1) / G_TM;
    const int64_t t0 = (int64_t)blockIdx.y * tiles_per_cta;
    const int64_t t1 = (t0 + tiles_per_cta < n_tiles) ? t0 + tiles_per_cta : n_tiles;
    if (t0 >= t1) return;
    const int64_t col0 = (int64_t)blockIdx.x * G_TN;

    const T* x0g = static_cast<const T*>(g.x0_soa);
    const T* x1g = static_cast<const T*>(g.x1_soa);

    if (tid == 0) {
        pn_mbar_init(&bars[0], 1);
        pn_mbar_init(&bars[1], 1);
        pn_mbar_init(&bars[2], 1);
        pn_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        pn_mbar_expect_tx(&bars[0], (uint32_t)(d * G_TN * sizeof(T)));
        for (int m = 0; m < d; ++m)
            pn_bulk_g2s(x1s + (size_t)m * G_TN, x1g + (size_t)m * g.ld1 + col0, G_TN * sizeof(T), &bars[0]);
        pn_mbar_expect_tx(&bars[1], (uint32_t)(d * G_TM * sizeof(T)));
        for (int m = 0; m < d; ++m)
            pn_bulk_g2s(x0s + (size_t)m * G_TM, x0g + (size_t)m * g.ld0 + t0 * G_TM, G_TM * sizeof(T), &bars[1]);
    }
    pn_stage_tables(tabE, tabL, tid, G_THREADS);
    if (ARD && tid < d) sc[tid] = static_cast<const T*>(g.scale)[tid];
    __syncthreads();
    pn_mbar_wait(&bars[0], 0);

    const int64_t j0 = col0 + 2 * tid;
    const T* x1t = x1s + 2 * tid;
    // how this thread stores its two columns (constant for the whole kernel): 3 = one vector
    // store, 2 = two scalar stores, 1 = first column only (last, odd column), 0 = outside
    const int store_mode = (j0 + 1 < g.n1) ? (vec_ok ? 3 : 2) : (j0 < g.n1 ? 1 : 0);
    const size_t row_stride = (size_t)g.ldo;
    TOut* optr = static_cast<TOut*>(g.out) + (size_t)t0 * G_TM * row_stride + j0;

    int it = 0;
    for (int64_t t = t0; t < t1; ++t, ++it) {
        const int buf = it & 1;
        if (tid == 0 && t + 1 < t1) {
            T* dst = x0s + (size_t)(buf ^ 1) * d * G_TM;
            pn_mbar_expect_tx(&bars[1 + (buf ^ 1)], (uint32_t)(d * G_TM * sizeof(T)));
            for (int m = 0; m < d; ++m)
                pn_bulk_g2s(dst + (size_t)m * G_TM, x0g + (size_t)m * g.ld0 + (t + 1) * G_TM,
                            G_TM * sizeof(T), &bars[1 + (buf ^ 1)]);
        }
        pn_mbar_wait(&bars[1 + buf], (uint32_t)((it >> 1) & 1));
        const T* x0t = x0s + (size_t)buf * d * G_TM;
        const int64_t row_base = t * G_TM;
        const int rows_here = (g.n0 - row_base < G_TM) ? (int)(g.n0 - row_base) : G_TM;

        for (int rg = 0; rg < rows_here; rg += G_R) {
            T S[2 * G_R];
#pragma unroll
            for (int i = 0; i < 2 * G_R; ++i) S[i] = T(0);
#pragma unroll 4
            for (int m = 0; m < d; ++m) {
                const P2 xb = *reinterpret_cast<const P2*>(x1t + (size_t)m * G_TN);
                T xa[G_R];
                if constexpr (std::is_same<T, double>::value) {
                    const double2 a01 = *reinterpret_cast<const double2*>(x0t + m * G_TM + rg);
                    const double2 a23 = *reinterpret_cast<const double2*>(x0t + m * G_TM + rg + 2);
                    xa[0] = a01.x; xa[1] = a01.y; xa[2] = a23.x; xa[3] = a23.y;
                } else {
                    const float4 a = *reinterpret_cast<const float4*>(x0t + m * G_TM + rg);
                    xa[0] = a.x; xa[1] = a.y; xa[2] = a.z; xa[3] = a.w;
                }
                T s = T(1);
                if constexpr (ARD) s = sc[m];
#pragma unroll
                for (int r = 0; r < G_R; ++r) {
                    T da = xa[r] - xb.x;
                    T db = xa[r] - xb.y;
                    if constexpr (ARD) {
                        da *= s;
                        db *= s;
                    }
                    if constexpr (std::is_same<T, double>::value) {
                        S[2 * r] = pn_fma(da, da, S[2 * r]);
                        S[2 * r + 1] = pn_fma(db, db, S[2 * r + 1]);
                    } else {
                        S[2 * r] = fmaf(da, da, S[2 * r]);
                        S[2 * r + 1] = fmaf(db, db, S[2 * r + 1]);
                    }
                }
            }
            TOut K[2 * G_R];
            if constexpr (FUSED) {
                unsigned eq = 0;
                if (F.has_wn) {  // S == 0 is necessary for equal points; squares may underflow, so confirm
#pragma unroll
                    for (int i = 0; i < 2 * G_R; ++i) eq |= (S[i] == T(0) ? 1u : 0u) << i;
                    if (eq) {
                        for (int m = 0; m < d; ++m) {
                            const P2 xb = *reinterpret_cast<const P2*>(x1t + (size_t)m * G_TN);
#pragma unroll
                            for (int r = 0; r < G_R; ++r) {
                                const T xa = x0t[m * G_TM + rg + r];
                                if (!(xa == xb.x)) eq &= ~(1u << (2 * r));
                                if (!(xa == xb.y)) eq &= ~(1u << (2 * r + 1));
                            }
                        }
                    }
                }
                fused_tile(S, K, F, eq, tabE, tabL);
            } else {
                epi_dispatch(S, K, g.epi, tabE, tabL);
                if constexpr (std::is_same<TOut, double>::value && std::is_same<T, double>::value) {
                    if (g.epi.post_flags) {  // uniform
                        unsigned eq = 0;
                        if (g.epi.post_flags & 2) {
#pragma unroll
                            for (int i = 0; i < 2 * G_R; ++i) eq |= (S[i] == 0.0 ? 1u : 0u) << i;
                            if (eq) {
                                for (int m = 0; m < d; ++m) {
                                    const P2 xb = *reinterpret_cast<const P2*>(x1t + (size_t)m * G_TN);
#pragma unroll
                                    for (int r = 0; r < G_R; ++r) {
                                        const T xa = x0t[m * G_TM + rg + r];
                                        if (!(xa == xb.x)) eq &= ~(1u << (2 * r));
                                        if (!(xa == xb.y)) eq &= ~(1u << (2 * r + 1));
                                    }
                                }
                            }
                        }
                        post_tile(K, g.epi, eq);
                    }
                }
            }
            const int rows_left = rows_here - rg;  // >= 1
#pragma unroll
            for (int r = 0; r < G_R; ++r) {
                if (r < rows_left) {
                    if (store_mode == 3) {
                        pn_store_cs2(optr, K[2 * r], K[2 * r + 1]);
                    } else if (store_mode != 0) {
                        pn_store_cs1(optr, K[2 * r]);
                        if (store_mode == 2) pn_store_cs1(optr + 1, K[2 * r + 1]);
                    }
                }
                optr += row_stride;
            }
        }
        optr += (size_t)(G_TM - ((rows_here + G_R - 1) / G_R) * G_R) * row_stride;  // to the next tile
        __syncthreads();  // everyone is done with x0 stage `buf` before it is refilled
    }
}

// ---- fast kernel: float64, compile-time d <= 8, compile-time epilogue, interior tiles only ------
// Preconditions (checked by the launcher): n0 % G_TM == 0, n1 % G_TN == 0, rows of the
// output 16-byte aligned.  Nothing in the loop is predicated; the column points of the
// thread's two columns live in registers; the output pointer advances by one row stride.
template <int D, bool ARD, int EPI>
__global__ void __launch_bounds__(G_THREADS)
    gram_fast_kernel(const __grid_constant__ PnGramLaunch g, int tiles_per_cta) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [0],[1]: x0 stages, [2]: tables
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    double* x0s = reinterpret_cast<double*>(smem + G_HDR_BYTES);  // [2][D][G_TM]

    const int n_tiles = (int)(g.n0 / G_TM);
    const int t0 = blockIdx.y * tiles_per_cta;
    const int t1 = min(t0 + tiles_per_cta, n_tiles);
    if (t0 >= t1) return;
    const int64_t j0 = (int64_t)blockIdx.x * G_TN + 2 * tid;
    const double* x0g = static_cast<const double*>(g.x0_soa);
    const double* x1g = static_cast<const double*>(g.x1_soa);
    constexpr uint32_t TILE_BYTES = D * G_TM * 8;

    if (tid == 0) {
        pn_mbar_init(&bars[0], 1);
        pn_mbar_init(&bars[1], 1);
        pn_mbar_init(&bars[2], 1);
        pn_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        pn_mbar_expect_tx(&bars[0], TILE_BYTES);
#pragma unroll
        for (int m = 0; m < D; ++m)
            pn_bulk_g2s(x0s + m * G_TM, x0g + (size_t)m * g.ld0 + (size_t)t0 * G_TM, G_TM * 8, &bars[0]);
        pn_stage_tables_tma(tabE, tabL, &bars[2], EPI == PN_EPI_RATQUADG);
    }

    double xb0[D], xb1[D], sc[D];
#pragma unroll
    for (int m = 0; m < D; ++m) {
        const double2 t = *reinterpret_cast<const double2*>(x1g + (size_t)m * g.ld1 + j0);
        xb0[m] = t.x;
        xb1[m] = t.y;
        if constexpr (ARD) sc[m] = static_cast<const double*>(g.scale)[m];
    }
    double* optr = static_cast<double*>(g.out) + (size_t)t0 * G_TM * g.ldo + j0;
    const size_t row_stride = (size_t)g.ldo;
    pn_mbar_wait(&bars[2], 0);

    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
        const int buf = it & 1;
        if (tid == 0 && t + 1 < t1) {
            double* dst = x0s + (buf ^ 1) * (D * G_TM);
            pn_mbar_expect_tx(&bars[buf ^ 1], TILE_BYTES);
#pragma unroll
            for (int m = 0; m < D; ++m)
                pn_bulk_g2s(dst + m * G_TM, x0g + (size_t)m * g.ld0 + (size_t)(t + 1) * G_TM, G_TM * 8,
                            &bars[buf ^ 1]);
        }
        pn_mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        const double* x0t = x0s + buf * (D * G_TM);
#pragma unroll 1
        for (int rg = 0; rg < G_TM; rg += G_R) {
            double S[2 * G_R];
#pragma unroll
            for (int i = 0; i < 2 * G_R; ++i) S[i] = 0.0;
#pragma unroll
            for (int m = 0; m < D; ++m) {
                const double2 a01 = *reinterpret_cast<const double2*>(x0t + m * G_TM + rg);
                const double2 a23 = *reinterpret_cast<const double2*>(x0t + m * G_TM + rg + 2);
                const double xa[G_R] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
                for (int r = 0; r < G_R; ++r) {
                    double da = xa[r] - xb0[m];
                    double db = xa[r] - xb1[m];
                    if constexpr (ARD) {
                        da *= sc[m];
                        db *= sc[m];
                    }
                    S[2 * r] = pn_fma(da, da, S[2 * r]);
                    S[2 * r + 1] = pn_fma(db, db, S[2 * r + 1]);
                }
            }
            double K[2 * G_R];
            epi_tile<EPI>(S, K, g.epi, tabE, tabL);
            if (g.epi.post_flags & 1) post_tile(K, g.epi, 0u);  // uniform; a WhiteNoise term keeps launches off this kernel
#pragma unroll
            for (int r = 0; r < G_R; ++r) {
                pn_store_cs2(optr, K[2 * r], K[2 * r + 1]);
                optr += row_stride;
            }
        }
        __syncthreads();  // everyone is done with x0 stage `buf` before it is refilled
    }
}


// ---- chunk kernel (any d): coordinates streamed through shared memory in chunks ------------
// d > G_DMAX.  The strip's column coordinates and the CTA's row coordinates arrive G_CD
// coordinates at a time through a two-stage TMA bulk-copy + mbarrier pipeline, so the loads of
// chunk c + 1 run under the arithmetic of chunk c (the first, unpipelined version spent more time
// waiting for a chunk than computing on it: 2.3 x off the FP64 roofline at d = 32..100).
constexpr int G_CD = 8;  // coordinates per pipeline stage
template <typename T, typename TOut, bool ARD>
__global__ void __launch_bounds__(G_THREADS)
    gram_chunk_kernel(const __grid_constant__ PnGramLaunch g, int vec_ok) {
    extern __shared__ __align__(128) unsigned char smem[];
    using P2 = typename Pair<T>::type;
    const int tid = threadIdx.x;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [0],[1]: chunk stages
    double* tabE = reinterpret_cast<double*>(smem + 128);
    double* tabL = tabE + PN_EXP2_N;
    T* sc = reinterpret_cast<T*>(smem + G_HDR_BYTES);                 // [2][G_CD]
    T* x1c = reinterpret_cast<T*>(smem + G_HDR_BYTES + 128);          // [2][G_CD][G_TN]
    T* x0c = x1c + (size_t)2 * G_CD * G_TN;                            // [2][G_CD][G_CHUNK_ROWS]
    constexpr int X1_STAGE = G_CD * G_TN, X0_STAGE = G_CD * G_CHUNK_ROWS;

    const int64_t col0 = (int64_t)blockIdx.x * G_TN;
    const int64_t row0 = (int64_t)blockIdx.y * G_CHUNK_ROWS;
    const T* x0g = static_cast<const T*>(g.x0_soa);
    const T* x1g = static_cast<const T*>(g.x1_soa);
    const int n_chunks = (g.d + G_CD - 1) / G_CD;

    auto issue = [&](int c, int buf) {  // thread 0 only
        const int c0 = c * G_CD;
        const int dc = (g.d - c0 < G_CD) ? g.d - c0 : G_CD;
        pn_mbar_expect_tx(&bars[buf], (uint32_t)(dc * (G_TN + G_CHUNK_ROWS) * sizeof(T)));
        for (int m = 0; m < dc; ++m) {
            pn_bulk_g2s(x1c + (size_t)buf * X1_STAGE + (size_t)m * G_TN, x1g + (size_t)(c0 + m) * g.ld1 + col0,
                        G_TN * sizeof(T), &bars[buf]);
            pn_bulk_g2s(x0c + (size_t)buf * X0_STAGE + (size_t)m * G_CHUNK_ROWS, x0g + (size_t)(c0 + m) * g.ld0 + row0,
                        G_CHUNK_ROWS * sizeof(T), &bars[buf]);  // rows are padded to PN_PAD
        }
    };

    if (tid == 0) {
        pn_mbar_init(&bars[0], 1);
        pn_mbar_init(&bars[1], 1);
        pn_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) issue(0, 0);
    pn_stage_tables(tabE, tabL, tid, G_THREADS);

    T S[2 * G_CHUNK_ROWS];
#pragma unroll
    for (int i = 0; i < 2 * G_CHUNK_ROWS; ++i) S[i] = T(0);

    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        const int c0 = c * G_CD;
        const int dc = (g.d - c0 < G_CD) ? g.d - c0 : G_CD;
        if (tid == 0 && c + 1 < n_chunks) issue(c + 1, buf ^ 1);
        if (ARD && tid < dc) sc[buf * G_CD + tid] = static_cast<const T*>(g.scale)[c0 + tid];
        __syncthreads();  // scale factors of this chunk visible (and the tables, first time round)
        pn_mbar_wait(&bars[buf], (uint32_t)((c >> 1) & 1));
        const T* x1b = x1c + (size_t)buf * X1_STAGE;
        const T* x0b = x0c + (size_t)buf * X0_STAGE;
        for (int m = 0; m < dc; ++m) {
            const P2 xb = *reinterpret_cast<const P2*>(x1b + (size_t)m * G_TN + 2 * tid);
            T s = T(1);
            if constexpr (ARD) s = sc[buf * G_CD + m];
#pragma unroll
            for (int r = 0; r < G_CHUNK_ROWS; ++r) {
                const T xa = x0b[m * G_CHUNK_ROWS + r];
                T da = xa - xb.x;
                T db = xa - xb.y;
                if constexpr (ARD) {
                    da *= s;
                    db *= s;
                }
                if constexpr (std::is_same<T, double>::value) {
                    S[2 * r] = pn_fma(da, da, S[2 * r]);
                    S[2 * r + 1] = pn_fma(db, db, S[2 * r + 1]);
                } else {
                    S[2 * r] = fmaf(da, da, S[2 * r]);
                    S[2 * r + 1] = fmaf(db, db, S[2 * r + 1]);
                }
            }
        }
        __syncthreads();  // everyone is done with stage `buf` before chunk c + 2 lands in it
    }
    TOut K[2 * G_CHUNK_ROWS];
    epi_dispatch(S, K, g.epi, tabE, tabL);
    const int64_t j0 = col0 + 2 * tid;
    TOut* out = static_cast<TOut*>(g.out);
#pragma unroll
    for (int r = 0; r < G_CHUNK_ROWS; ++r) {
        const int64_t row = row0 + r;
        if (row < g.n0) store_pair(out + row * g.ldo, j0, g.n1, K[2 * r], K[2 * r + 1], vec_ok != 0);
    }
}

template <int D, bool ARD, int EPI>
cudaError_t launch_fast_inst(const PnGramLaunch& g, cudaStream_t stream) {
    const int64_t strips = g.n1 / G_TN;
    const int64_t n_tiles = g.n0 / G_TM;
    int64_t tpc = (n_tiles * strips) / ((int64_t)g.sm_count * 48);
    tpc = tpc < 1 ? 1 : (tpc > 32 ? 32 : tpc);
    const int64_t gy = (n_tiles + tpc - 1) / tpc;
    if (gy > 65535) return cudaErrorInvalidConfiguration;
    const size_t smem = G_HDR_BYTES + 2 * sizeof(double) * D * G_TM;
    dim3 grid((unsigned)strips, (unsigned)gy);
    gram_fast_kernel<D, ARD, EPI><<<grid, G_THREADS, smem, stream>>>(g, (int)tpc);
    ++g_pn_launches;
    return cudaGetLastError();
}

template <int D, bool ARD>
cudaError_t launch_fast_epi(const PnGramLaunch& g, cudaStream_t stream) {
    switch (g.epi.epi) {
        case PN_EPI_EXPQUAD: return launch_fast_inst<D, ARD, PN_EPI_EXPQUAD>(g, stream);
        case PN_EPI_MATERN0:
            return launch_fast_inst<D, ARD, PN_EPI_MATERN0>(g, stream);
        case PN_EPI_MATERN1:
            return launch_fast_inst<D, ARD, PN_EPI_MATERN1>(g, stream);
        case PN_EPI_MATERN2:
            return launch_fast_inst<D, ARD, PN_EPI_MATERN2>(g, stream);
        case PN_EPI_MATERNP:
            return launch_fast_inst<D, ARD, PN_EPI_MATERNP>(g, stream);
        case PN_EPI_RATQUAD1:
            return launch_fast_inst<D, false, PN_EPI_RATQUAD1>(g, stream);
        case PN_EPI_RATQUAD2:
            return launch_fast_inst<D, false, PN_EPI_RATQUAD2>(g, stream);
        default:
            return launch_fast_inst<D, false, PN_EPI_RATQUADG>(g, stream);
    }
}

template <bool ARD>
cudaError_t launch_fast(const PnGramLaunch& g, cudaStream_t stream) {
    switch (g.d) {
        case 1: return launch_fast_epi<1, ARD>(g, stream);
        case 2: return launch_fast_epi<2, ARD>(g, stream);
        case 3: return launch_fast_epi<3, ARD>(g, stream);
        case 4: return launch_fast_epi<4, ARD>(g, stream);
        case 5: return launch_fast_epi<5, ARD>(g, stream);
        case 6: return launch_fast_epi<6, ARD>(g, stream);
        case 7: return launch_fast_epi<7, ARD>(g, stream);
        case 8: return launch_fast_epi<8, ARD>(g, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <typename T, typename TOut, bool ARD, typename PROG = PnNoProgram>
cudaError_t launch_typed(const PnGramLaunch& g, cudaStream_t stream, const PROG& F = PROG()) {
    if (g.n0 <= 0 || g.n1 <= 0) return cudaSuccess;
    const int64_t strips = (g.n1 + G_TN - 1) / G_TN;
    const int vec_ok = ((g.ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(g.out) % (2 * sizeof(TOut))) == 0) ? 1 : 0;
    if (g.d <= G_DMAX) {
        const int64_t n_tiles = (g.n0 + G_TM - 1) / G_TM;
        // enough CTAs for ~16 waves over the SMs, at least one row tile each
        int64_t tpc = (n_tiles * strips) / ((int64_t)g.sm_count * 48);
        if (tpc < 1) tpc = 1;
        if (tpc > 32) tpc = 32;
        const int64_t gy = (n_tiles + tpc - 1) / tpc;
        if (gy > 65535) return cudaErrorInvalidConfiguration;
        const size_t smem = G_HDR_BYTES + 128 + sizeof(T) * ((size_t)g.d * G_TN + 2 * (size_t)g.d * G_TM);
        auto kern = gram_strip_kernel<T, TOut, ARD, PROG>;
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        dim3 grid((unsigned)strips, (unsigned)gy);
        kern<<<grid, G_THREADS, smem, stream>>>(g, (int)tpc, vec_ok, F);
    } else {
        if (!std::is_same<PROG, PnNoProgram>::value) return cudaErrorInvalidValue;  // programs: d <= G_DMAX
        const int64_t gy = (g.n0 + G_CHUNK_ROWS - 1) / G_CHUNK_ROWS;
        if (gy > 65535) {
            // split tall outputs into several launches
            cudaError_t err = cudaSuccess;
            const int64_t rows_per_launch = 65535LL * G_CHUNK_ROWS;
            for (int64_t r0 = 0; r0 < g.n0 && err == cudaSuccess; r0 += rows_per_launch) {
                PnGramLaunch part = g;
                part.n0 = (g.n0 - r0 < rows_per_launch) ? g.n0 - r0 : rows_per_launch;
                part.x0_soa = static_cast<const T*>(g.x0_soa) + r0;
                part.out = static_cast<TOut*>(g.out) + r0 * g.ldo;
                err = launch_typed<T, TOut, ARD, PROG>(part, stream, F);
            }
            return err;
        }
        const size_t smem = G_HDR_BYTES + 128 + sizeof(T) * 2 * ((size_t)G_CD * G_TN + (size_t)G_CD * G_CHUNK_ROWS);
        auto kern = gram_chunk_kernel<T, TOut, ARD>;
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        dim3 grid((unsigned)strips, (unsigned)gy);
        kern<<<grid, G_THREADS, smem, stream>>>(g, vec_ok);
    }
    ++g_pn_launches;
    return cudaGetLastError();
}

}  // namespace

static cudaError_t launch_generic(const PnGramLaunch& g, cudaStream_t stream) {
    const bool ard = g.scale != nullptr;
    if (!g.in_f32) {
        return ard ? launch_typed<double, double, true>(g, stream) : launch_typed<double, double, false>(g, stream);
    }
    if (g.out_f32) {
        return ard ? launch_typed<float, float, true>(g, stream) : launch_typed<float, float, false>(g, stream);
    }
    return ard ? launch_typed<float, double, true>(g, stream) : launch_typed<float, double, false>(g, stream);
}

cudaError_t pn_launch_gram(const PnGramLaunch& g, cudaStream_t stream) {
    if (g.n0 <= 0 || g.n1 <= 0) return cudaSuccess;
    const bool aligned = (g.ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(g.out) % 16) == 0;
    int64_t n0i = (g.n0 / G_TM) * G_TM, n1i = (g.n1 / G_TN) * G_TN;  // interior extent
    // mirror stores pay only where the FP64 pipe, not HBM, limits the general kernel
    const bool post_wn = (g.epi.post_flags & 2) != 0;  // `+ WhiteNoise`: needs the kernels that can compare points
    const bool fp64_heavy = post_wn || (g.epi.epi != PN_EPI_EXPQUAD && g.epi.epi != PN_EPI_RATQUAD1 && g.epi.epi != PN_EPI_RATQUAD2);
    const bool sym = g.symmetric && fp64_heavy && g.n0 == g.n1 && g.x0_soa == g.x1_soa && !g.in_f32 && g.d <= 8 &&
                     aligned && g.n0 >= 4 * Y_T && !(post_wn && g.scale != nullptr);
    if (sym) n0i = n1i = (g.n0 / Y_T) * Y_T;  // the symmetric kernel works on the 64 grid
    if (g.in_f32 || g.d > 8 || !aligned || n0i == 0 || n1i == 0 || (post_wn && !sym)) return launch_generic(g, stream);
    // interior: whole 64-row tiles x whole 512-column strips, unpredicated specialised kernel
    PnGramLaunch in = g;
    in.n0 = n0i;
    in.n1 = n1i;
    in.symmetric = sym ? 1 : 0;
    if (in.symmetric) in.n0 = n1i;  // square interior; the rest goes to the edge launches
    const int64_t n0_done = in.n0;
    cudaError_t err = in.symmetric ? pn_gram_sym_launch(in, stream)  // pncov_gram_sym.cu
                                   : (g.scale ? launch_fast<true>(in, stream) : launch_fast<false>(in, stream));
    if (err != cudaSuccess) return err;
    if (n1i < g.n1) {  // right edge: all rows, remaining columns
        PnGramLaunch e = g;
        e.symmetric = 0;
        e.x1_soa = static_cast<const double*>(g.x1_soa) + n1i;
        e.n1 = g.n1 - n1i;
        e.out = static_cast<double*>(g.out) + n1i;
        if ((err = launch_generic(e, stream)) != cudaSuccess) return err;
    }
    if (n0_done < g.n0) {  // bottom edge: remaining rows, interior columns
        PnGramLaunch e = g;
        e.x0_soa = static_cast<const double*>(g.x0_soa) + n0_done;
        e.n0 = g.n0 - n0_done;
        e.n1 = n1i;
        e.symmetric = 0;
        e.out = static_cast<double*>(g.out) + n0_done * g.ldo;
        if ((err = launch_generic(e, stream)) != cudaSuccess) return err;
    }
    return cudaSuccess;
}

// Covariance arithmetic in one launch: the program F over isotropic float64 leaves that share S (d <= 16).
// A whole symmetric matrix (d <= 8, aligned) takes the mirrored-store kernel on its 64-grid interior -- the
// leaves' FP64 work per stored entry halves, which is what keeps a three-leaf sum HBM-bound -- and the
// general strip kernel on the edges; everything else is the strip kernel.
cudaError_t pn_launch_gram_fused(const PnGramLaunch& g, const PnFusedProgram& F, cudaStream_t stream) {
    if (g.n0 <= 0 || g.n1 <= 0) return cudaSuccess;
    if (g.in_f32 || g.scale != nullptr || g.d > G_DMAX || F.n_slots < 1 || F.n_slots > PN_FUSED_MAX_LEAVES) return cudaErrorInvalidValue;
    const bool aligned = (g.ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(g.out) % 16) == 0;
    const bool sym = g.symmetric && g.n0 == g.n1 && g.x0_soa == g.x1_soa && g.d <= 8 && aligned && g.n0 >= 4 * Y_T;
    if (!sym) return launch_typed<double, double, false, PnFusedProgram>(g, stream, F);
    const int64_t ni = (g.n0 / Y_T) * Y_T;
    PnGramLaunch in = g;
    in.n0 = in.n1 = ni;
    in.symmetric = 1;
    cudaError_t err = pn_gram_sym_launch_fused(in, F, stream);  // pncov_gram_sym.cu
    if (err != cudaSuccess) return err;
    if (ni < g.n1) {  // right edge: all rows, remaining columns
        PnGramLaunch e = g;
        e.symmetric = 0;
        e.x1_soa = static_cast<const double*>(g.x1_soa) + ni;
        e.n1 = g.n1 - ni;
        e.out = static_cast<double*>(g.out) + ni;
        if ((err = launch_typed<double, double, false, PnFusedProgram>(e, stream, F)) != cudaSuccess) return err;
        // bottom edge: remaining rows, interior columns
        e = g;
        e.symmetric = 0;
        e.x0_soa = static_cast<const double*>(g.x0_soa) + ni;
        e.n0 = g.n0 - ni;
        e.n1 = ni;
        e.out = static_cast<double*>(g.out) + ni * g.ldo;
        if ((err = launch_typed<double, double, false, PnFusedProgram>(e, stream, F)) != cudaSuccess) return err;
    }
    return cudaSuccess;
}

// ProductMatern whose dimensions share the degree p <= 3: symmetric interior (64 grid) of a whole symmetric Gram
// through the mirrored-store kernel.  Returns the interior extent in *done (0: not applicable, nothing launched);
// the caller writes the edges with the general kernel (pn_launch_gram_extra).
cudaError_t pn_launch_gram_pm_sym(const PnGramLaunch& g, const double* pm_host, cudaStream_t stream, int64_t* done) {
    *done = 0;
    const bool aligned = (g.ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(g.out) % 16) == 0;
    if (!(g.symmetric && g.n0 == g.n1 && g.x0_soa == g.x1_soa && g.d <= 8 && aligned && g.n0 >= 4 * Y_T)) return cudaSuccess;
    const int p = (int)pm_host[1];
    if (p < 0 || p > 3) return cudaSuccess;
    PnPmProgram F{};
    for (int m = 0; m < g.d; ++m) {
        if ((int)pm_host[m * PN_PM_LD + 1] != p) return cudaSuccess;  // mixed degrees: general kernel
        for (int i = 0; i < PN_PM_LD; ++i) F.pm[m][i] = pm_host[m * PN_PM_LD + i];
    }
    const int64_t ni = (g.n0 / Y_T) * Y_T;
    PnGramLaunch in = g;
    in.n0 = in.n1 = ni;
    in.symmetric = 1;
    cudaError_t err = pn_gram_sym_launch_pm(in, F, p, stream);  // pncov_gram_sym.cu
    if (err == cudaSuccess) *done = ni;
    return err;
}
