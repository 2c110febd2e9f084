// Host side of the symmetric Gram return path: worker threads that write the lower triangle of a
// symmetric result as the mirror image of its upper triangle while the next row block is still
// crossing PCIe.
//
// CovarianceFunction.matrix(X) must hand the caller a full (N, N) host array
// (/root/reference/src/probnum/randprocs/covfuncs/_covariance_function.py:407-411).  The device builds
// the matrix at 6.5 TB/s, the PCIe link moves 57 GB/s: the end-to-end call is bound by the
// device-to-host copy.  K(X, X) is symmetric bit for bit (every kernel of this library evaluates
// K_ij and K_ji with the same operations), so only the blocks on and above the diagonal are
// copied and the rest is produced here, on the host's own memory bandwidth, overlapped with the
// copy: half the PCIe bytes per call.
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <emmintrin.h>
#include <stdlib.h>
#include <string.h>

#include "pncov_internal.h"

namespace {

class HostPool {
  public:
    explicit HostPool(int n) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    void submit(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(std::move(f));
            ++pending_;
        }
        cv_.notify_one();
    }
    void wait_idle() {
        std::unique_lock<std::mutex> lk(m_);
        idle_.wait(lk, [this] { return pending_ == 0; });
    }

  private:
    void run() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                f = std::move(q_.back());
                q_.pop_back();
            }
            f();
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) idle_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::vector<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_, idle_;
    int pending_ = 0;
    bool stop_ = false;
};

HostPool* g_pool = nullptr;
std::mutex g_pool_mutex;
bool g_atfork_registered = false;

// a forked child has none of the parent's threads: forget the pool (leaked on purpose) and start a new one on demand
void pool_atfork_child() {
    g_pool = nullptr;
    new (&g_pool_mutex) std::mutex();
}

HostPool* pool() {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    if (!g_pool) {
        int n = (int)std::thread::hardware_concurrency();
        if (const char* env = getenv("PNCOV_HOST_THREADS")) n = atoi(env);
        n = n < 1 ? 1 : (n > 32 ? 32 : n);
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) {  // respect the process's CPU binding (one rank per GPU)
            const int allowed = CPU_COUNT(&set);
            if (allowed >= 1 && allowed < n) n = allowed;
        }
        g_pool = new HostPool(n);
        if (!g_atfork_registered) {
            pthread_atfork(nullptr, nullptr, pool_atfork_child);
            g_atfork_registered = true;
        }
    }
    return g_pool;
}

// dst[c][r] = src[r][c] for r in [r0, r1), c in [c0, c1) of one row-major matrix with leading dimension ld
// (elements).  The rows of such a matrix are hundreds of KB apart, so every row of a tile sits in its own 4 KB
// page: with small tiles the copy is bound by TLB misses, not by memory bandwidth (measured: 1.4 GB/s per thread
// with 32 x 32 tiles).  Tiles of 256 x 256 elements touch a page for 2 KB at a time on both sides; a tile is copied
// into a contiguous buffer, transposed there (L2-resident, SSE2 unpack for float64) and leaves as streaming stores
// of whole 2 KB destination rows.
constexpr int HT = 256;
constexpr int HP = HT + 8;  // padded row: a power-of-two stride would map a whole column of the tile to four L1 sets
template <typename T>
struct MirrorScratch {
    T a[HT][HP];
    T b[HT][HP];
};

template <typename T>
void mirror_rect(T* base, int64_t ld, int64_t r0, int64_t r1, int64_t c0, int64_t c1) {
    MirrorScratch<T>* sc = nullptr;
    if (posix_memalign(reinterpret_cast<void**>(&sc), 64, sizeof(MirrorScratch<T>)) != 0 || !sc) return;
    for (int64_t rr = r0; rr < r1; rr += HT) {
        const int nr = (int)((r1 - rr < HT) ? r1 - rr : HT);
        for (int64_t cc = c0; cc < c1; cc += HT) {
            const int nc = (int)((c1 - cc < HT) ? c1 - cc : HT);
            for (int i = 0; i < nr; ++i) memcpy(&sc->a[i][0], base + (rr + i) * ld + cc, (size_t)nc * sizeof(T));
            if (sizeof(T) == 8 && nr == HT && nc == HT) {
                const double* a = reinterpret_cast<const double*>(&sc->a[0][0]);
                double* b = reinterpret_cast<double*>(&sc->b[0][0]);
                for (int i = 0; i < HT; i += 2) {
                    for (int j = 0; j < HT; j += 2) {
                        const __m128d x = _mm_load_pd(a + i * HP + j), y = _mm_load_pd(a + (i + 1) * HP + j);
                        _mm_store_pd(b + j * HP + i, _mm_unpacklo_pd(x, y));
                        _mm_store_pd(b + (j + 1) * HP + i, _mm_unpackhi_pd(x, y));
                    }
                }
            } else {
                for (int i = 0; i < nr; ++i)
                    for (int j = 0; j < nc; ++j) sc->b[j][i] = sc->a[i][j];
            }
            for (int j = 0; j < nc; ++j) {
                T* d = base + (cc + j) * ld + rr;
                if (sizeof(T) == 8 && nr == HT && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
                    const double* b = reinterpret_cast<const double*>(&sc->b[j][0]);
                    double* dd = reinterpret_cast<double*>(d);
                    for (int i = 0; i < HT; i += 2) _mm_stream_pd(dd + i, _mm_load_pd(b + i));
                } else {
                    memcpy(d, &sc->b[j][0], (size_t)nr * sizeof(T));
                }
            }
        }
    }
    _mm_sfence();
    free(sc);
}

}  // namespace

int pn_host_threads() { return pool()->size(); }

// write zeros to [p, p + bytes) from all worker threads: populates the pages of a fresh mapping in parallel
// (page faults, not bandwidth, bound a single thread to ~2 GB/s)
void pn_host_touch(void* p, size_t bytes) {
    HostPool* hp = pool();
    const size_t chunk = (size_t)32 << 20;
    for (size_t o = 0; o < bytes; o += chunk) {
        const size_t len = bytes - o < chunk ? bytes - o : chunk;
        char* q = static_cast<char*>(p) + o;
        hp->submit([=] { memset(q, 0, len); });
    }
    hp->wait_idle();
}

// Queue the mirror image of rows [r0, r1) x columns [r1, n) of the (n, n) host matrix `out`: destination rows
// [r1, n), columns [r0, r1).  Split over the pool by destination rows.  elem = 4 or 8 bytes.
void pn_host_mirror_async(void* out, int64_t ld, int64_t n, int64_t r0, int64_t r1, int elem) {
    if (r1 >= n || r0 >= r1) return;
    HostPool* p = pool();
    const int64_t span = n - r1;
    int64_t chunk = (span + 4 * p->size() - 1) / (4 * p->size());
    chunk = ((chunk + HT - 1) / HT) * HT;
    for (int64_t c0 = r1; c0 < n; c0 += chunk) {
        const int64_t c1 = (c0 + chunk < n) ? c0 + chunk : n;
        if (elem == 8)
            p->submit([=] { mirror_rect<double>(static_cast<double*>(out), ld, r0, r1, c0, c1); });
        else
            p->submit([=] { mirror_rect<float>(static_cast<float*>(out), ld, r0, r1, c0, c1); });
    }
}

void pn_host_mirror_wait() { pool()->wait_idle(); }
