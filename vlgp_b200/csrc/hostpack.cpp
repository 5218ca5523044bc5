// Host-side packing loops of the upload path (plain C++, no CUDA): spike counts arrive as float64 arrays (the
// reference promotes y to float64, vlgp/core.py:60) and are stored in HBM as uint8 when every entry is an integer in
// [0, 255].  The conversion + exactness check reads 8 bytes and writes 1 per entry and runs in the
// host threads of the pinned upload pipeline (capi.cu: vlgp_trials_set_y_parts); it is the largest host cost of a
// vem() call with host buffers (204 MB of float64 per call at 256 trials x 1000 bins x 100 neurons), so it gets an AVX2
// body chosen at run time: 2.8 -> 5.8 GB/s of float64 per thread.
#include <stdint.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/vlgp_b200.h"
#include "hostpack.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define VLGP_HAVE_AVX2_DISPATCH 1
#endif

namespace {

bool f64_to_u8_scalar(const double *src, unsigned char *dst, int64_t cnt) {
    bool ok = true;
    for (int64_t k = 0; k < cnt; ++k) {
        const double v = src[k];
        const bool in = (v >= 0.0) & (v <= 255.0);                     // also false for NaN
        const unsigned char c = in ? (unsigned char)(int)v : (unsigned char)0;
        dst[k] = c;
        ok &= in & ((double)c == v);                                   // exact only for integer counts
    }
    return ok;
}

#ifdef VLGP_HAVE_AVX2_DISPATCH
__attribute__((target("avx2"))) bool f64_to_u8_avx2(const double *src, unsigned char *dst, int64_t cnt) {
    int64_t k = 0;
    __m256d bad = _mm256_setzero_pd();
    const __m256d lo = _mm256_set1_pd(0.0), hi = _mm256_set1_pd(255.0);
    const __m256d ones = _mm256_castsi256_pd(_mm256_set1_epi64x(-1));
    for (; k + 16 <= cnt; k += 16) {
        __m128i q[4];
        for (int j = 0; j < 4; ++j) {
            const __m256d v = _mm256_loadu_pd(src + k + 4 * j);
            const __m128i i32 = _mm256_cvttpd_epi32(v);               // truncates; out of range / NaN -> INT_MIN
            const __m256d back = _mm256_cvtepi32_pd(i32);
            // an exact count: v equals its truncation and lies in [0, 255] (ordered compares: NaN fails all three)
            const __m256d good = _mm256_and_pd(
                _mm256_cmp_pd(v, back, _CMP_EQ_OQ),
                _mm256_and_pd(_mm256_cmp_pd(v, lo, _CMP_GE_OQ), _mm256_cmp_pd(v, hi, _CMP_LE_OQ)));
            bad = _mm256_or_pd(bad, _mm256_andnot_pd(good, ones));
            q[j] = i32;
        }
        // saturating packs: what they store for an inexact entry is irrelevant, the caller discards the whole buffer
        const __m128i w0 = _mm_packus_epi32(q[0], q[1]), w1 = _mm_packus_epi32(q[2], q[3]);
        _mm_storeu_si128((__m128i *)(dst + k), _mm_packus_epi16(w0, w1));
    }
    bool ok = _mm256_movemask_pd(bad) == 0;
    if (k < cnt) ok &= f64_to_u8_scalar(src + k, dst + k, cnt - k);
    return ok;
}
#endif

}   // namespace

// dst[k] = (uint8) src[k] for k < cnt; returns 1 when every src[k] is an integer in [0, 255] (then dst is exact), else 0
// (dst then holds unspecified bytes).  Thread-safe; called concurrently on disjoint ranges.
int vlgp_host_f64_to_u8(const double *src, unsigned char *dst, int64_t cnt) {
#ifdef VLGP_HAVE_AVX2_DISPATCH
    static const int have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return f64_to_u8_avx2(src, dst, cnt) ? 1 : 0;
#endif
    return f64_to_u8_scalar(src, dst, cnt) ? 1 : 0;
}

// Which body vlgp_host_f64_to_u8 runs on this machine: 1 = AVX2, 0 = scalar (tests / diagnostics).
int vlgp_host_pack_isa(void) {
#ifdef VLGP_HAVE_AVX2_DISPATCH
    return __builtin_cpu_supports("avx2") ? 1 : 0;
#else
    return 0;
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent host thread pool of the packing pipeline.  The gather / scatter / convert steps of one upload or download
// are split over a few host threads per 8 MB staging chunk; creating those threads anew for every chunk cost ~100
// thread creations per vem() call with host buffers.  The workers are created once (grown on demand up to 15), sleep on
// a condition variable between calls and are never joined (the pool object is leaked on purpose: no destructor runs at
// process exit).  run() executes fn(0..n-1), each index exactly once, on the workers and on the calling thread, and
// returns when all have finished.  Calls are serialised; a forked child (no threads survive a fork) rebuilds the pool.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

class HostPool {
  public:
    static HostPool &instance() {
        static HostPool *p = new HostPool();
        return *p;
    }

    void run(int n, vlgp_host_task fn, void *arg) {
        if (n <= 1) {
            if (n == 1) fn(0, arg);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mutex_);
        if (pid_ != getpid()) reset_after_fork();
        const int want = n - 1 < kMaxWorkers ? n - 1 : kMaxWorkers;
        while ((int)workers_.size() < want) workers_.emplace_back(&HostPool::worker, this);
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = fn;
            arg_ = arg;
            n_tasks_ = n;
            next_ = 0;
            pending_ = n;
            ++generation_;
        }
        cv_work_.notify_all();
        std::unique_lock<std::mutex> lk(m_);
        drain(lk);                                             // the caller works too
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
        n_tasks_ = 0;
    }

    int n_workers() {
        std::lock_guard<std::mutex> serial(run_mutex_);
        return (int)workers_.size();
    }

  private:
    static constexpr int kMaxWorkers = 15;
    HostPool() : pid_(getpid()) {}

    // takes task indices until none is left; called and returns with the lock held
    void drain(std::unique_lock<std::mutex> &lk) {
        while (next_ < n_tasks_) {
            const int t = next_++;
            vlgp_host_task fn = fn_;
            void *arg = arg_;
            lk.unlock();
            fn(t, arg);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }

    void worker() {
        std::unique_lock<std::mutex> lk(m_);
        uint64_t seen = 0;
        for (;;) {
            cv_work_.wait(lk, [&] { return generation_ != seen; });
            seen = generation_;
            drain(lk);
        }
    }

    void reset_after_fork() {
        // the parent's worker threads do not exist in this process: forget them (their std::thread objects must not be
        // destroyed while joinable, so the vector is leaked) and start with fresh synchronisation state
        new (&workers_) std::vector<std::thread>();
        new (&m_) std::mutex();
        new (&cv_work_) std::condition_variable();
        new (&cv_done_) std::condition_variable();
        fn_ = nullptr;
        n_tasks_ = next_ = pending_ = 0;
        pid_ = getpid();
    }

    std::mutex run_mutex_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> workers_;
    vlgp_host_task fn_ = nullptr;
    void *arg_ = nullptr;
    int n_tasks_ = 0, next_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    pid_t pid_;
};

}   // namespace

void vlgp_host_parallel(int n, vlgp_host_task fn, void *arg) { HostPool::instance().run(n, fn, arg); }

// Self-test of the pool (tests/test_abi.py): `rounds` calls with n_tasks tasks each; every task index of every call
// must run exactly once and run() must not return before all of them have.  Returns 0 when that held, else the number
// of violations; *workers (may be NULL) receives the number of pool threads afterwards.
int vlgp_host_pool_selftest(int n_tasks, int rounds, int *workers) {
    if (n_tasks < 0 || rounds < 0) return -1;
    struct Ctx {
        std::vector<std::atomic<int>> hits;
        std::atomic<int64_t> sum{0};
        explicit Ctx(int n) : hits(n) {}
    };
    int bad = 0;
    for (int r = 0; r < rounds; ++r) {
        Ctx c(n_tasks);
        for (auto &h : c.hits) h.store(0);
        vlgp_host_parallel(
            n_tasks,
            [](int t, void *a) {
                Ctx *cx = static_cast<Ctx *>(a);
                cx->hits[t].fetch_add(1);
                volatile double x = 1.0;                      // a little work so that the tasks overlap in time
                for (int i = 0; i < 200 * (1 + t % 3); ++i) x = x * 1.0000001 + 1e-9;
                cx->sum.fetch_add(t + 1);
            },
            &c);
        for (int t = 0; t < n_tasks; ++t)
            if (c.hits[t].load() != 1) ++bad;
        if (c.sum.load() != (int64_t)n_tasks * (n_tasks + 1) / 2) ++bad;
    }
    if (workers) *workers = HostPool::instance().n_workers();
    return bad;
}
