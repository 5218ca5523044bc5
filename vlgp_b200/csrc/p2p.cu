// Host side of the peer-memory allreduce (p2p.cuh): mailbox allocation, IPC handle exchange, channel bookkeeping.
#include "common.cuh"
#include "p2p.cuh"

struct P2PState {
    int n_ranks = 1, rank = 0;
    double *base = nullptr;                               // this rank's mailbox (2 channels)
    double *peer[VLGP_P2P_MAX_RANKS] = {nullptr};         // every rank's mailbox as mapped here (peer[rank] == base)
    unsigned long long seq[2] = {0, 0};
    int *d_err = nullptr;
    int *h_err = nullptr;                                 // pinned
};

bool vlgp_p2p_enabled(const vlgp_ctx *ctx) { return ctx->p2p != nullptr; }

P2PDev vlgp_p2p_next(vlgp_ctx *ctx) {
    P2PDev d{};
    d.n_ranks = 1;
    P2PState *st = (P2PState *)ctx->p2p;
    if (!st) return d;
    const int ch = ctx->p2p_chan;
    d.n_ranks = st->n_ranks;
    d.rank = st->rank;
    d.seq = ++st->seq[ch];
    d.err = st->d_err;
    for (int r = 0; r < st->n_ranks; ++r) d.mail[r] = st->peer[r] + (size_t)ch * VLGP_P2P_CHANNEL_DOUBLES;
    ctx->counters[3]++;
    return d;
}

namespace {
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(P2PDev p, double *buf, size_t n) {
    const size_t off = (size_t)blockIdx.x * VLGP_P2P_CHUNK;
    const int cnt = (int)(n - off < (size_t)VLGP_P2P_CHUNK ? n - off : (size_t)VLGP_P2P_CHUNK);
    p2p_allreduce_cta(p, blockIdx.x, off, buf + off, cnt);
}
}   // namespace

int vlgp_p2p_allreduce(vlgp_ctx *ctx, double *d_buf, size_t n) {
    if (!ctx->p2p || n == 0) return VLGP_OK;
    for (size_t done = 0; done < n; done += VLGP_P2P_PAY) {
        const size_t cnt = n - done < (size_t)VLGP_P2P_PAY ? n - done : (size_t)VLGP_P2P_PAY;
        P2PDev d = vlgp_p2p_next(ctx);
        p2p_allreduce_kernel<<<(unsigned)((cnt + VLGP_P2P_CHUNK - 1) / VLGP_P2P_CHUNK), 256, 0, ctx->stream>>>(d, d_buf + done, cnt);
        CKL();
    }
    return VLGP_OK;
}

int vlgp_p2p_check(vlgp_ctx *ctx) {
    P2PState *st = (P2PState *)ctx->p2p;
    if (!st) return VLGP_OK;
    CK(cudaMemcpyAsync(st->h_err, st->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (*st->h_err)
        return vlgp_fail(ctx, VLGP_ERR_NCCL, "peer-memory allreduce timed out (a rank is gone or out of step)");
    return VLGP_OK;
}

void vlgp_p2p_destroy(vlgp_ctx *ctx) {
    P2PState *st = (P2PState *)ctx->p2p;
    if (!st) return;
    for (int r = 0; r < st->n_ranks; ++r)
        if (r != st->rank && st->peer[r]) cudaIpcCloseMemHandle(st->peer[r]);
    if (st->base) cudaFree(st->base);
    if (st->d_err) cudaFree(st->d_err);
    if (st->h_err) cudaFreeHost(st->h_err);
    delete st;
    ctx->p2p = nullptr;
}

extern "C" {

// Collective over the ranks of the communicator (all of them on this node: needs the shared-memory host channel for the
// handle exchange).  *enabled = 1 when every rank mapped every peer's mailbox; otherwise nothing changes (NCCL path).
static int p2p_abort(vlgp_ctx *ctx, P2PState *st, const char *why) {
    ctx->p2p = st;
    vlgp_p2p_destroy(ctx);
    return vlgp_fail(ctx, VLGP_ERR_NCCL, "comm_enable_p2p: %s", why);
}

int vlgp_comm_enable_p2p(vlgp_ctx *ctx, int *enabled) {
    if (!ctx) return VLGP_ERR_ARG;
    if (enabled) *enabled = 0;
    if (ctx->n_ranks <= 1) return VLGP_OK;
    REQUIRE(ctx->shm != nullptr, "comm_enable_p2p: attach the shared-memory host channel first (single node only)");
    REQUIRE(ctx->p2p == nullptr, "comm_enable_p2p: already enabled");
    if (ctx->n_ranks > VLGP_P2P_MAX_RANKS) return VLGP_OK;
    CK(cudaSetDevice(ctx->device));
    P2PState *st = new P2PState();
    st->n_ranks = ctx->n_ranks;
    st->rank = ctx->rank_id;
    const size_t bytes = 2 * VLGP_P2P_CHANNEL_DOUBLES * sizeof(double);
    double ok = 1.0;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&st->base, bytes) != cudaSuccess || cudaMemset(st->base, 0, bytes) != cudaSuccess ||
        cudaMalloc(&st->d_err, sizeof(int)) != cudaSuccess || cudaMemset(st->d_err, 0, sizeof(int)) != cudaSuccess ||
        cudaHostAlloc(&st->h_err, sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, st->base) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        cudaGetLastError();
        ok = 0.0;
    }
    // gather the handles: one shared-memory allreduce per source rank (the 64 handle bytes as small integers)
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::vector<cudaIpcMemHandle_t> all(ctx->n_ranks);
    for (int r = 0; r < ctx->n_ranks; ++r) {
        double tmp[65];
        for (int i = 0; i < 65; ++i) tmp[i] = 0.0;
        if (r == ctx->rank_id) {
            const unsigned char *b = (const unsigned char *)&mine;
            for (int i = 0; i < 64; ++i) tmp[i] = (double)b[i];
            tmp[64] = 1.0 - ok;
        }
        if (vlgp_shm_allreduce(ctx->shm, tmp, 65, 0) != VLGP_OK) return p2p_abort(ctx, st, "handle exchange timed out");
        unsigned char *b = (unsigned char *)&all[r];
        for (int i = 0; i < 64; ++i) b[i] = (unsigned char)tmp[i];
        if (tmp[64] != 0.0) ok = 0.0;
    }
    if (ok != 0.0) {
        for (int r = 0; r < ctx->n_ranks; ++r) {
            if (r == ctx->rank_id) {
                st->peer[r] = st->base;
                continue;
            }
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0.0;
                break;
            }
            st->peer[r] = (double *)ptr;
        }
    }
    // unanimous: everyone mapped everything, or nobody uses it
    double bad = ok != 0.0 ? 0.0 : 1.0;
    if (vlgp_shm_allreduce(ctx->shm, &bad, 1, 0) != VLGP_OK) return p2p_abort(ctx, st, "handle exchange timed out");
    ctx->p2p = st;
    if (bad != 0.0) {
        vlgp_p2p_destroy(ctx);
        return VLGP_OK;
    }
    if (enabled) *enabled = 1;
    return VLGP_OK;
}

}   // extern "C"
