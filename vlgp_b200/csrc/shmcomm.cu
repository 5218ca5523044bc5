// Host-side allreduce between the processes of ONE node through a POSIX shared-memory segment.
//
// Several per-iteration quantities of the engine are needed on the HOST, not on the device: the (ll, dll) pair of every
// H-step objective evaluation feeds scipy's L-BFGS-B, the norms feed vem's convergence test.  Reducing them with NCCL
// costs a collective launch plus a device round trip per call for 16-80 bytes, dozens of times per EM iteration; the
// processes share a box (one process per GPU of one node), so the partial results -- already copied to the host -- are
// summed here instead: every rank publishes its vector in its own slot, waits until all ranks have published the same
// call number, and adds the slots in rank order (the same order on every rank: bit-identical results everywhere).
// Two banks indexed by call parity make the slots reusable without a second barrier: a rank can only publish call k + 2
// after every rank has published k + 1, i.e. after every rank has finished reading call k.
// Pure host code (no CUDA calls): the entry points work without a GPU and are exercised by the CPU test-suite.
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

#include "../../include/vlgp_b200.h"

namespace {

constexpr int kMaxRanks = 64;
constexpr int kMaxLen = 256;
constexpr uint64_t kMagic = 0x766c67705f73686dULL;     // "vlgp_shm"

struct alignas(64) Slot {
    std::atomic<uint64_t> seq;
    double data[kMaxLen];
};

struct Segment {
    std::atomic<uint64_t> magic;
    int32_t n_ranks;
    int32_t pad_[13];
    Slot slot[2][kMaxRanks];        // [bank][rank]
};

struct Handle {
    Segment *seg = nullptr;
    int rank = 0, n_ranks = 1;
    uint64_t call = 0;
    char name[96] = {0};
    double timeout_s = 120.0;
};

static_assert(std::atomic<uint64_t>::is_always_lock_free, "shared-memory sequence numbers need lock-free atomics");

bool wait_for(const std::atomic<uint64_t> &a, uint64_t want, double timeout_s) {
    for (int spin = 0; spin < 4096; ++spin) {
        if (a.load(std::memory_order_acquire) >= want) return true;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t it = 0;
    while (a.load(std::memory_order_acquire) < want) {
        if ((++it & 0xff) == 0) {
            std::this_thread::yield();
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) return false;
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    return true;
}

}   // namespace

extern "C" {

int vlgp_shm_open(const char *name, int rank, int n_ranks, void **handle) {
    if (!name || !handle || name[0] != '/' || strlen(name) >= sizeof(Handle::name)) return VLGP_ERR_ARG;
    if (n_ranks < 1 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks) return VLGP_ERR_ARG;
    *handle = nullptr;
    int fd = -1;
    if (rank == 0) {
        shm_unlink(name);      // a stale segment of a crashed job with the same name
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, sizeof(Segment)) != 0) {
            if (fd >= 0) close(fd);
            return VLGP_ERR_NCCL;
        }
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {         // rank 0 may not have created / sized it yet
            fd = shm_open(name, O_RDWR, 0600);
            struct stat st;
            if (fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size >= sizeof(Segment)) break;
            if (fd >= 0) close(fd);
            fd = -1;
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0) return VLGP_ERR_NCCL;
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
    }
    void *p = mmap(nullptr, sizeof(Segment), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return VLGP_ERR_NCCL;
    Segment *seg = (Segment *)p;
    if (rank == 0) {           // ftruncate zero-fills: all sequence numbers start at 0
        seg->n_ranks = n_ranks;
        seg->magic.store(kMagic, std::memory_order_release);
    } else if (!wait_for(seg->magic, kMagic, 60.0) || seg->magic.load() != kMagic || seg->n_ranks != n_ranks) {
        munmap(p, sizeof(Segment));
        return VLGP_ERR_NCCL;
    }
    Handle *h = new Handle();
    h->seg = seg;
    h->rank = rank;
    h->n_ranks = n_ranks;
    snprintf(h->name, sizeof(h->name), "%s", name);
    *handle = h;
    return VLGP_OK;
}

int vlgp_shm_allreduce(void *handle, double *buf, int n, int op) {
    Handle *h = (Handle *)handle;
    if (!h || !buf || n < 0 || n > kMaxLen || (op != 0 && op != 1)) return VLGP_ERR_ARG;
    if (h->n_ranks == 1 || n == 0) return VLGP_OK;
    const uint64_t k = ++h->call;
    Slot *bank = h->seg->slot[k & 1];
    Slot &mine = bank[h->rank];
    memcpy(mine.data, buf, (size_t)n * sizeof(double));
    mine.seq.store(k, std::memory_order_release);
    for (int r = 0; r < h->n_ranks; ++r)
        if (!wait_for(bank[r].seq, k, h->timeout_s)) return VLGP_ERR_NCCL;      // a peer died or fell out of step
    for (int i = 0; i < n; ++i) {
        double acc = bank[0].data[i];
        for (int r = 1; r < h->n_ranks; ++r) {
            const double x = bank[r].data[i];
            acc = op == 1 ? (x > acc ? x : acc) : acc + x;
        }
        buf[i] = acc;
    }
    return VLGP_OK;
}

int vlgp_shm_close(void *handle, int unlink_name) {
    Handle *h = (Handle *)handle;
    if (!h) return VLGP_OK;
    if (unlink_name) shm_unlink(h->name);
    munmap((void *)h->seg, sizeof(Segment));
    delete h;
    return VLGP_OK;
}

}   // extern "C"
