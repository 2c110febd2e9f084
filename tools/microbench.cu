// Hardware ceilings that MEASURED_PEAKS.json does not record and the design needs:
//   1. FP64 DFMA throughput (vector pipe), sustained
//   2. FP64 DMMA (mma.sync m8n8k4 / m16n8k8) throughput, alone and mixed with DFMA
//   3. write-only HBM bandwidth with 16-byte streaming stores (the Gram kernel's ceiling)
//   4. pinned D2H bandwidth (the end-to-end ceiling of CovarianceFunction.matrix)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/build/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// mode 0: m8n8k4 only; mode 1: m16n8k8 only; mode 2: m8n8k4 + 8 DFMA per 4 MMA; mode 3: m16n8k8 + DFMA
__global__ void dmma_kernel(double* out, int iters, int mode, double a, double b) {
    double c[8][2];
    double c4[4][4];
    double f[8];
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; f[i] = threadIdx.x + i; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c4[i][j] = i + j;
    double av = a + threadIdx.x, bv = b - threadIdx.x;
    double a4[4] = {av, av + 1, av + 2, av + 3};
    double b2[2] = {bv, bv + 1};
    if (mode == 0 || mode == 2) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], av, bv);
            if (mode == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fma(f[j], a, b);
            }
        }
    } else {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma1688(c4[j], a4, b2);
            if (mode == 3) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fma(f[j], a, b);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c4[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void write_kernel(double2* out, size_t n2, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) __stcs(out + i, make_double2(v, v + 1));
}
__global__ void write_kernel_plain(double2* out, size_t n2, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) out[i] = make_double2(v, v + 1);
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n", p.name, sms);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int blocks = sms * 8, threads = 256;

    {   // DFMA, short burst and ~2 s sustained
        int iters = 1 << 14;
        float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double flops = 2.0 * 8 * iters * (double)blocks * threads;
        printf(" \"dfma_tflops_burst\": %.2f,\n", flops / ms / 1e9);
        int big = 1 << 20;
        ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, big, 1.0000001, 1e-9); }, 2);
        flops = 2.0 * 8 * big * (double)blocks * threads;
        printf(" \"dfma_tflops_sustained\": %.2f, \"dfma_sustained_ms\": %.1f,\n", flops / ms / 1e9, ms);
    }
    for (int mode = 0; mode < 4; ++mode) {
        int iters = 1 << 13;
        float ms = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, iters, mode, 1.0000001, 1e-9); }, 5);
        double warps = (double)blocks * threads / 32;
        double mma_flops = (mode == 0 || mode == 2) ? 8.0 * (2.0 * 8 * 8 * 4) : 4.0 * (2.0 * 16 * 8 * 8);
        double fma_flops = (mode >= 2) ? 8.0 * 2 * 32 : 0.0;
        printf(" \"dmma_mode%d_mma_tflops\": %.2f, \"dmma_mode%d_fma_tflops\": %.2f,\n", mode,
               mma_flops * iters * warps / ms / 1e9, mode, fma_flops * iters * warps / ms / 1e9);
    }
    {
        size_t bytes = (size_t)8 << 30;
        double2* buf; CK(cudaMalloc(&buf, bytes));
        float ms = time_ms([&] { write_kernel<<<sms * 16, 512>>>(buf, bytes / 16, 1.0); }, 5);
        printf(" \"write_cs_gbs\": %.1f,\n", bytes / ms / 1e6);
        ms = time_ms([&] { write_kernel_plain<<<sms * 16, 512>>>(buf, bytes / 16, 1.0); }, 5);
        printf(" \"write_plain_gbs\": %.1f,\n", bytes / ms / 1e6);
        ms = time_ms([&] { CK(cudaMemsetAsync(buf, 0, bytes)); }, 5);
        printf(" \"memset_gbs\": %.1f,\n", bytes / ms / 1e6);
        void* host; size_t hb = (size_t)2 << 30;
        CK(cudaHostAlloc(&host, hb, cudaHostAllocDefault));
        ms = time_ms([&] { CK(cudaMemcpyAsync(host, buf, hb, cudaMemcpyDeviceToHost)); }, 3);
        printf(" \"d2h_pinned_gbs\": %.1f,\n", hb / ms / 1e6);
        ms = time_ms([&] { CK(cudaMemcpyAsync(buf, host, hb, cudaMemcpyHostToDevice)); }, 3);
        printf(" \"h2d_pinned_gbs\": %.1f\n", hb / ms / 1e6);
        CK(cudaFreeHost(host)); CK(cudaFree(buf));
    }
    printf("}\n");
    return 0;
}
