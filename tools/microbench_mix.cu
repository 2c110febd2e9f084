// How much FP64-pipe time does a DMMA.8x8x4 cost when it is mixed with DFMA at the ratio of the
// norm-identity matvec (1 DMMA : 9 DFMA per warp)?  Answers whether DMMA is worth using there.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/microbench_mix tools/microbench_mix.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NM DMMAs (independent accumulators) followed by NF DFMAs (NF/8 rounds over 8 chains) per iteration;
// INTERLEAVE: 1 DMMA then NF/NM DFMAs, repeated
template <int NM, int NF, bool INTERLEAVE>
__global__ void mix_kernel(double* out, int iters, double a, double b) {
    double c[8][2], f[8];
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; f[i] = threadIdx.x + i; }
    const double av = a + threadIdx.x, bv = b - threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        if (!INTERLEAVE) {
#pragma unroll
            for (int j = 0; j < NM; ++j) dmma884(c[j % 8][0], c[j % 8][1], av, bv);
#pragma unroll
            for (int j = 0; j < NF; ++j) f[j % 8] = fma(f[j % 8], a, b);
        } else {
#pragma unroll
            for (int j = 0; j < NM; ++j) {
                dmma884(c[j % 8][0], c[j % 8][1], av, bv);
#pragma unroll
                for (int k = 0; k < NF / NM; ++k) f[(j * (NF / NM) + k) % 8] = fma(f[(j * (NF / NM) + k) % 8], a, b);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NM, int NF, bool IL>
void run(const char* name, double* out, int blocks_per_sm, int threads, int sms) {
    const int iters = 1 << 12;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    mix_kernel<NM, NF, IL><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        mix_kernel<NM, NF, IL><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double warps = (double)sms * blocks_per_sm * threads / 32;
    const double warps_per_smsp = warps / (sms * 4.0);
    // cycles per iteration per SMSP (all its warps): time * clock / iters
    const double cyc = best * 1e-3 * 1.965e9 / iters / warps_per_smsp;
    const double dfma_cyc = 2.0 * NF;
    printf("%-34s NM=%d NF=%2d warps/SMSP=%4.1f  %8.3f ms  %6.1f cyc/iter/warp  -> %5.1f cyc per DMMA if DFMA=2 (nominal 1.965 GHz)\n",
           name, NM, NF, warps_per_smsp, best, cyc, NM ? (cyc - dfma_cyc) / NM : 0.0);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    run<0, 72, false>("DFMA only", out, 8, 256, sms);
    run<8, 0, false>("DMMA only", out, 8, 256, sms);
    run<8, 8, false>("8 DMMA + 8 DFMA", out, 8, 256, sms);
    run<8, 72, false>("8 DMMA + 72 DFMA (burst)", out, 8, 256, sms);
    run<8, 72, true>("8 x (1 DMMA + 9 DFMA)", out, 8, 256, sms);
    run<8, 72, false>("8 DMMA + 72 DFMA, 4 warps/SMSP", out, 4, 128, sms);
    run<8, 72, true>("8 x (1 DMMA + 9 DFMA), 4 w/SMSP", out, 4, 128, sms);
    run<8, 72, false>("8 DMMA + 72 DFMA, 2 warps/SMSP", out, 2, 128, sms);
    run<8, 0, false>("DMMA only, 4 warps/SMSP", out, 4, 128, sms);
    run<8, 0, false>("DMMA only, 1 warp/SMSP", out, 1, 128, sms);
    run<0, 72, false>("DFMA only, 1 warp/SMSP", out, 1, 128, sms);
    return 0;
}
