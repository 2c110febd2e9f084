// Host-side feasibility probe for mirroring the lower triangle of a symmetric Gram matrix on the CPU
// (so that only the upper trapezoid crosses PCIe): multithreaded blocked transposition rate.
//   g++ -O3 -march=native -pthread tools/host_mirror_bench.cpp -o tools/build/host_mirror_bench
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cstring>

static void mirror_block(double* K, long n, long r0, long r1, long j0, long j1) {
    // dest rows j in [j0, j1), cols [r0, r1)  <-  src rows [r0, r1), col j
    const long B = 32;
    for (long jb = j0; jb < j1; jb += B)
        for (long ib = r0; ib < r1; ib += B) {
            const long je = jb + B < j1 ? jb + B : j1, ie = ib + B < r1 ? ib + B : r1;
            for (long j = jb; j < je; ++j)
                for (long i = ib; i < ie; ++i) K[j * n + i] = K[i * n + j];
        }
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 16384;
    const int T = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
    const long blk = 4096;
    double* K = (double*)aligned_alloc(4096, (size_t)n * n * 8);
    memset(K, 1, (size_t)n * n * 8);
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        double bytes = 0;
        for (long r0 = 0; r0 < n; r0 += blk) {
            const long r1 = r0 + blk < n ? r0 + blk : n;
            if (r1 >= n) break;
            std::vector<std::thread> th;
            const long rows = n - r1, per = (rows + T - 1) / T;
            for (int t = 0; t < T; ++t) {
                const long j0 = r1 + t * per, j1 = j0 + per < n ? j0 + per : n;
                if (j0 < j1) th.emplace_back(mirror_block, K, n, r0, r1, j0, j1);
            }
            for (auto& x : th) x.join();
            bytes += (double)(n - r1) * (r1 - r0) * 8;
        }
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("n=%ld threads=%d mirrored %.2f GB in %.3f s = %.1f GB/s (written)\n", n, T, bytes / 1e9, dt, bytes / 1e9 / dt);
    }
    return 0;
}
