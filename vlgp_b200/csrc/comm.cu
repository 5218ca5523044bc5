// Multi-GPU plumbing: one process per GPU; the only data-path collective is a sum-allreduce of small fp64 buffers
// (M-step sufficient statistics, H-step partial sums, vem norms) over NCCL, issued on the context's stream so that it
// orders with the producing and consuming kernels without a host round trip.
// libnccl is dlopen'ed (RTLD_LOCAL) from the path the host passes, so this library has no link-time dependency on NCCL
// and never collides with another libnccl already loaded in the process.
#include <dlfcn.h>

#include "common.cuh"
#include "p2p.cuh"

struct NcclId {
    char internal[128];      // NCCL_UNIQUE_ID_BYTES
};

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclId *id) = nullptr;
    int (*CommInitRank)(void **comm, int nranks, NcclId id, int rank) = nullptr;
    int (*AllReduce)(const void *send, void *recv, size_t count, int dtype, int op, void *comm,
                     cudaStream_t s) = nullptr;
    int (*CommDestroy)(void *comm) = nullptr;
    int (*CommSplit)(void *comm, int color, int key, void **newcomm, void *config) = nullptr;   // NCCL >= 2.18
    const char *(*GetErrorString)(int) = nullptr;
};

static const int kNcclFloat64 = 8;   // ncclFloat64
static const int kNcclSum = 0, kNcclMax = 2;

static int load_nccl(vlgp_ctx *ctx, const char *path) {
    if (ctx->nccl) return VLGP_OK;
    const char *p = (path && path[0]) ? path : "libnccl.so.2";
    void *h = dlopen(p, RTLD_NOW | RTLD_LOCAL);
    if (!h) return vlgp_fail(ctx, VLGP_ERR_NCCL, "dlopen(%s): %s", p, dlerror());
    NcclApi *api = new NcclApi();
    api->handle = h;
    api->GetUniqueId = (int (*)(NcclId *))dlsym(h, "ncclGetUniqueId");
    api->CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    api->AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclAllReduce");
    api->CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    api->GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    api->CommSplit = (int (*)(void *, int, int, void **, void *))dlsym(h, "ncclCommSplit");     // optional
    if (!api->GetUniqueId || !api->CommInitRank || !api->AllReduce || !api->CommDestroy || !api->GetErrorString) {
        delete api;
        dlclose(h);
        return vlgp_fail(ctx, VLGP_ERR_NCCL, "%s does not export the NCCL entry points", p);
    }
    ctx->nccl = api;
    return VLGP_OK;
}

#define NCK(call)                                                                                          \
    do {                                                                                                   \
        int r_ = (call);                                                                                   \
        if (r_ != 0)                                                                                       \
            return vlgp_fail(ctx, VLGP_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call,              \
                             ctx->nccl->GetErrorString(r_));                                               \
    } while (0)

int vlgp_allreduce_dev(vlgp_ctx *ctx, double *d_buf, size_t n, int op) {
    if (ctx->n_ranks <= 1 || n == 0) return VLGP_OK;
    // peer-memory mailboxes when enabled (same node): lower latency than an NCCL launch for these small buffers and a
    // fixed summation order; larger buffers go through it in pieces of the payload size
    if (op == 0 && vlgp_p2p_enabled(ctx)) return vlgp_p2p_allreduce(ctx, d_buf, n);
    if (!ctx->comm) return vlgp_fail(ctx, VLGP_ERR_NCCL, "allreduce without a communicator");
    NCK(ctx->nccl->AllReduce(d_buf, d_buf, n, kNcclFloat64, op == 1 ? kNcclMax : kNcclSum, ctx->comm, ctx->stream));
    ctx->counters[3]++;
    return VLGP_OK;
}

void vlgp_comm_destroy(vlgp_ctx *ctx) {
    vlgp_p2p_destroy(ctx);
    if (ctx->shm) vlgp_shm_close(ctx->shm, 0);
    ctx->shm = nullptr;
    if (ctx->comm_m && ctx->nccl) ctx->nccl->CommDestroy(ctx->comm_m);
    ctx->comm_m = nullptr;
    if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->comm);
    ctx->comm = nullptr;
    if (ctx->nccl) {
        // the handle is intentionally not dlclose'd: NCCL keeps background threads alive until process exit
        delete ctx->nccl;
        ctx->nccl = nullptr;
    }
    ctx->n_ranks = 1;
    ctx->rank_id = 0;
}

extern "C" {

int vlgp_comm_unique_id(vlgp_ctx *ctx, const char *libnccl_path, char id[128]) {
    if (!ctx || !id) return VLGP_ERR_ARG;
    int rc = load_nccl(ctx, libnccl_path);
    if (rc) return rc;
    NcclId nid;
    NCK(ctx->nccl->GetUniqueId(&nid));
    memcpy(id, nid.internal, 128);
    return VLGP_OK;
}

int vlgp_comm_init(vlgp_ctx *ctx, const char *libnccl_path, int rank, int n_ranks, const char id[128]) {
    if (!ctx || !id) return VLGP_ERR_ARG;
    REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "comm_init: bad rank %d / %d", rank, n_ranks);
    REQUIRE(ctx->comm == nullptr, "comm_init: communicator already initialised");
    if (n_ranks == 1) return VLGP_OK;
    if (getenv("VLGP_COMM_NO_NCCL")) {
        // ranks of one node that will talk through peer memory and shared memory only (vlgp_comm_enable_p2p must
        // succeed): also lets several ranks share ONE GPU, which NCCL refuses -- the multi-rank path is then testable
        // on a single-GPU box
        ctx->rank_id = rank;
        ctx->n_ranks = n_ranks;
        return VLGP_OK;
    }
    int rc = load_nccl(ctx, libnccl_path);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    NcclId nid;
    memcpy(nid.internal, id, 128);
    void *comm = nullptr;
    NCK(ctx->nccl->CommInitRank(&comm, n_ranks, nid, rank));
    ctx->comm = comm;
    ctx->rank_id = rank;
    ctx->n_ranks = n_ranks;
    // A duplicate communicator for the M-step stream (vlgp_mstep_begin), so that its allreduces can be in flight
    // while the H-step's run on the main stream.  Every rank issues the operations of both in the same order.
    if (ctx->nccl->CommSplit && !getenv("VLGP_NO_COMM_SPLIT")) {
        void *dup = nullptr;
        NCK(ctx->nccl->CommSplit(comm, 0, rank, &dup, nullptr));
        ctx->comm_m = dup;
    }
    return VLGP_OK;
}

int vlgp_comm_allreduce(vlgp_ctx *ctx, double *buf, int n, int op) {
    if (!ctx || !buf) return VLGP_ERR_ARG;
    REQUIRE(n >= 0 && n <= 256, "comm_allreduce: n %d outside [0, 256]", n);
    if (ctx->n_ranks <= 1 || n == 0) return VLGP_OK;
    if (ctx->shm) {
        if (vlgp_shm_allreduce(ctx->shm, buf, n, op) != VLGP_OK)
            return vlgp_fail(ctx, VLGP_ERR_NCCL, "comm_allreduce: shared-memory allreduce timed out (a rank is gone or out of step)");
        return VLGP_OK;
    }
    CK(cudaSetDevice(ctx->device));
    double *stage = ctx->h_pin + 256;          // second half of the 4 KB pinned block
    double *dstage = ctx->d_small + 256;
    memcpy(stage, buf, n * sizeof(double));
    CK(cudaMemcpyAsync(dstage, stage, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int rc = vlgp_allreduce_dev(ctx, dstage, n, op);
    if (rc) return rc;
    CK(cudaMemcpyAsync(stage, dstage, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(buf, stage, n * sizeof(double));
    return VLGP_OK;
}

int vlgp_comm_attach_shm(vlgp_ctx *ctx, void *handle) {
    if (!ctx || !handle) return VLGP_ERR_ARG;
    REQUIRE(ctx->shm == nullptr, "comm_attach_shm: a handle is already attached");
    ctx->shm = handle;
    return VLGP_OK;
}

int vlgp_comm_allreduce_bulk(vlgp_ctx *ctx, double *buf, int64_t n) {
    if (!ctx || !buf) return VLGP_ERR_ARG;
    REQUIRE(n >= 0, "comm_allreduce_bulk: negative length");
    if (ctx->n_ranks <= 1 || n == 0) return VLGP_OK;
    CK(cudaSetDevice(ctx->device));
    double *d = nullptr;
    CK(vlgp_dalloc(ctx, &d, (size_t)n * sizeof(double)));
    CK(cudaMemcpyAsync(d, buf, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int rc = vlgp_allreduce_dev(ctx, d, (size_t)n, 0);
    if (rc) return rc;
    CK(cudaMemcpyAsync(buf, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(vlgp_dfree(ctx, d));
    return VLGP_OK;
}

}   // extern "C"
