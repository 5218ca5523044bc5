// Small sum-allreduce between the GPUs of one node through peer memory (NVLink / NVSwitch), callable from INSIDE a
// kernel: the producing kernel's epilogue pushes its partial result straight into every peer's mailbox with remote
// stores, raises a flag, waits for the peers' flags and adds the contributions in rank order.  This is the exchange step
// of the path's three latency-bound reductions -- the M-step's per-Newton-iteration sufficient statistics (25 dependent
// reductions of ~20 KB per M-step, vlgp/core.py:174-202 when trials are sharded), the H-step's per-evaluation partial
// sums, the second moments of mu -- fused with the kernels that produce / consume them instead of an NCCL launch in
// between (mstep.cu: reduce + exchange + solve in one kernel; hstep.cu: final reduction + exchange).
//
// Memory: every rank owns one mailbox per channel (channel 0: main stream, channel 1: the overlapped M-step's stream),
// cudaMalloc'ed and exported with cudaIpcGetMemHandle; the peers map it with cudaIpcOpenMemHandle (p2p.cu).  Layout of
// a channel: [bank 0|1][source rank][PAY doubles payload | NCH flags]; a call is identified by the channel's sequence
// number (host-tracked, identical on every rank because the ranks issue the same calls in the same order).  The bank is
// the parity of the sequence number: a rank can push call k + 2 only after every peer has raised its flag for k + 1,
// which it does after it has finished reading call k -- two banks make the slots reusable without a second barrier.
// Results are bit-identical on every rank (fixed summation order).  A peer that never shows up ends the wait after
// VLGP_P2P_TIMEOUT_CYCLES and raises the error flag the host checks at its next synchronisation.
#pragma once
#include <stdint.h>

#define VLGP_P2P_MAX_RANKS 8
#define VLGP_P2P_PAY 65536            // payload doubles per (bank, source)
#define VLGP_P2P_NCH 256              // chunks (flags) per (bank, source)
#define VLGP_P2P_CHUNK 256            // doubles per chunk of the generic kernel
#define VLGP_P2P_SLOT (VLGP_P2P_PAY + VLGP_P2P_NCH)
#define VLGP_P2P_CHANNEL_DOUBLES ((size_t)2 * VLGP_P2P_MAX_RANKS * VLGP_P2P_SLOT)
#define VLGP_P2P_TIMEOUT_CYCLES 8000000000LL       // ~4 s at 2 GHz

struct P2PDev {
    int n_ranks, rank;
    unsigned long long seq;               // this call's sequence number (>= 1)
    int *err;                             // device error flag (timeouts)
    double *mail[VLGP_P2P_MAX_RANKS];     // the channel's base in every rank's mailbox, as mapped in THIS process
};

#ifdef __CUDACC__
__device__ __forceinline__ size_t p2p_slot(unsigned long long seq, int src) {
    return ((size_t)(seq & 1ULL) * VLGP_P2P_MAX_RANKS + src) * VLGP_P2P_SLOT;
}

// Sum-allreduce of buf[0..n) (global or shared memory, in place) by ONE CTA; `chunk` < NCH identifies the flag, `off`
// (+ n <= PAY) the payload range this CTA owns in the slots.  Every CTA of every rank that takes part in call `seq`
// must use the same (chunk, off, n).  Ends with a block barrier.
__device__ __forceinline__ void p2p_allreduce_cta(const P2PDev &p, int chunk, size_t off, double *buf, int n) {
    const int tid = threadIdx.x, nt = blockDim.x, R = p.n_ranks;
    if (R <= 1) return;
    const size_t mine = p2p_slot(p.seq, p.rank);
    for (int dst = 0; dst < R; ++dst) {
        double *slot = p.mail[dst] + mine + off;
        for (int i = tid; i < n; i += nt) slot[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < R) {
        unsigned long long *f = (unsigned long long *)(p.mail[tid] + mine + VLGP_P2P_PAY) + chunk;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(p.seq) : "memory");
        const unsigned long long *g = (const unsigned long long *)(p.mail[p.rank] + p2p_slot(p.seq, tid) + VLGP_P2P_PAY) + chunk;
        const long long t0 = clock64();
        unsigned long long seen;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(g) : "memory");
            if (seen >= p.seq) break;
            if (clock64() - t0 > VLGP_P2P_TIMEOUT_CYCLES) {
                atomicExch(p.err, 1);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        double s = 0.0;
        for (int r = 0; r < R; ++r) s += __ldcg(p.mail[p.rank] + p2p_slot(p.seq, r) + off + i);
        buf[i] = s;
    }
    __syncthreads();
}
#endif

struct vlgp_ctx;
// Host side (p2p.cu).  vlgp_p2p_next: the descriptor of the next call on the context's current channel (advances its
// sequence number); n_ranks == 1 in the result means "no exchange" (single rank, or peer memory not enabled).
bool vlgp_p2p_enabled(const vlgp_ctx *ctx);
P2PDev vlgp_p2p_next(vlgp_ctx *ctx);
int vlgp_p2p_allreduce(vlgp_ctx *ctx, double *d_buf, size_t n);      // generic kernel on ctx->stream, n <= PAY
int vlgp_p2p_check(vlgp_ctx *ctx);                                   // after a synchronisation: error flag -> status
void vlgp_p2p_destroy(vlgp_ctx *ctx);
