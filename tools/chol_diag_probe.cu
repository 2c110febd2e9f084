// Phase times of chol_diag_kernel (csrc/pncov_chol.cu) on one SM: clock64 stamps compiled in with PN_DIAG_CLOCKS.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I probnum_b200/csrc -o tools/build/chol_diag_probe tools/chol_diag_probe.cu
//   tools/build/chol_diag_probe
#define PN_DIAG_CLOCKS 1
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../probnum_b200/csrc/pncov_chol.cu"

int64_t g_pn_launches = 0;
int pn_sm_count() { return 148; }

int main() {
    const int n = CB;
    std::vector<double> hA((size_t)n * n), hP((size_t)n * n);
    srand(1);
    for (auto& v : hP) v = (rand() / (double)RAND_MAX - 0.5) * 0.1;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = (i == j) ? 4.0 : 0.0;
            for (int l = 0; l < n; ++l) s += 2.0 * hP[(size_t)i * n + l] * hP[(size_t)j * n + l];
            hA[(size_t)i * n + j] = s;
        }
    double *A, *A0, *P, *inv;
    int* info;
    cudaMalloc(&A, sizeof(double) * n * n);
    cudaMalloc(&A0, sizeof(double) * n * n);
    cudaMalloc(&P, sizeof(double) * n * n);
    cudaMalloc(&inv, sizeof(double) * n * n);
    cudaMalloc(&info, sizeof(int));
    cudaMemset(info, 0, sizeof(int));
    cudaMemcpy(A0, hA.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
    cudaMemcpy(P, hP.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(double) * (CB * CD_LD + 5 * CB + 3 * 32 * 33);
    cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char* names[] = {"load + P update", "columns 0-31", "columns 32-63", "columns 64-95", "columns 96-127", "L out + transpose", "inverse (A)",
                           "inverse (B) dist 1", "inverse (B) dist 2", "inverse (B) dist 3", "inverse out"};
    constexpr int NPH = 11;
    double acc[NPH] = {0};
    long long last_load = 0;
    float ms_sum = 0;
    const int reps = 20;
    for (int it = 0; it < reps + 2; ++it) {
        cudaMemcpy(A, A0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e0);
        chol_diag_kernel<<<1, 256, smem>>>(A, n, inv, info, 0, P);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        long long clk[16];
        cudaMemcpyFromSymbol(clk, g_diag_clk, sizeof(clk));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) {
            ms_sum += ms;
            for (int k = 0; k < NPH; ++k) acc[k] += (double)(clk[k + 1] - clk[k]);
            last_load = clk[12] - clk[0];
        }
    }
    int hinfo = -1;
    cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost);
    printf("cudaGetLastError: %s, info %d\n", cudaGetErrorString(cudaGetLastError()), hinfo);
    printf("kernel (events, includes launch): %.2f us\n", 1e3 * ms_sum / reps);
    double tot = 0;
    for (int k = 0; k < NPH; ++k) tot += acc[k] / reps;
    for (int k = 0; k < NPH; ++k) printf("%-20s %9.0f cycles  %5.1f %%\n", names[k], acc[k] / reps, 100.0 * acc[k] / reps / tot);
    printf("%-20s %9.0f cycles\n", "total", tot);
    printf("(block and panel head in shared memory / registers after %lld cycles of the first phase, last launch)\n", last_load);
    return 0;
}
