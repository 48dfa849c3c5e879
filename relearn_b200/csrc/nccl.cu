// nccl.cu -- data-parallel group.  relearn has no collectives; this replaces the crossbeam thread
// fan-out of src/simulation/train.rs:124-158 across GPUs.  Rollouts shard by lane and need no
// exchange; the update kernels all-reduce their f64 partial sums (gradient, Fisher-vector product,
// loss/KL scalars) through rl_allreduce_f64_inplace().
//
// NCCL is bound at run time (dlopen) so the library has no link-time dependency on a particular
// libnccl and shares the copy torch already loaded when the host process uses torch.distributed.
#include <dlfcn.h>

#include <cstdlib>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(void **, int, nccl_uid, int);
typedef int (*fn_destroy)(void *);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_allgather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*fn_errstr)(int);

struct NcclApi {
    void *handle = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_allgather allgather = nullptr;
    fn_errstr errstr = nullptr;
    bool tried = false;
};

NcclApi &api() {
    static NcclApi a;
    if (!a.tried) {
        a.tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (a.handle) {
            a.get_uid = (fn_get_uid)dlsym(a.handle, "ncclGetUniqueId");
            a.init_rank = (fn_init_rank)dlsym(a.handle, "ncclCommInitRank");
            a.destroy = (fn_destroy)dlsym(a.handle, "ncclCommDestroy");
            a.allreduce = (fn_allreduce)dlsym(a.handle, "ncclAllReduce");
            a.allgather = (fn_allgather)dlsym(a.handle, "ncclAllGather");
            a.errstr = (fn_errstr)dlsym(a.handle, "ncclGetErrorString");
        }
    }
    return a;
}

bool api_ok() {
    NcclApi &a = api();
    return a.handle && a.get_uid && a.init_rank && a.destroy && a.allreduce && a.errstr;
}

constexpr int NCCL_FLOAT64 = 8;  // ncclDataType_t ncclFloat64
constexpr int NCCL_INT8 = 0;     // ncclDataType_t ncclInt8 / ncclChar
constexpr int NCCL_SUM = 0;      // ncclRedOp_t ncclSum

size_t x_data_bytes(int world) { return (size_t)2 * world * RL_X_BLOCKS * RL_X_SLOT * sizeof(double); }
size_t x_flag_bytes(int world) { return (size_t)2 * world * RL_X_BLOCKS * sizeof(unsigned long long); }

// Allocate this rank's mailbox, exchange CUDA IPC handles through the communicator and map every peer's mailbox.
// Any failure leaves x_ok false and the update keeps using ncclAllReduce.
void x_setup(rl_ctx *ctx) {
    const char *e = getenv("RL_XREDUCE");
    if (e && (e[0] == 'n' || e[0] == '0')) return;  // RL_XREDUCE=nccl
    const int world = ctx->world, rank = ctx->rank;
    if (world > RL_X_MAX_RANKS || !api().allgather) return;
    const size_t bytes = x_data_bytes(world) + x_flag_bytes(world) + 256;
    struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
    Msg mine{};
    // The gather and vote buffers come first: a rank that cannot even allocate them must not leave its peers waiting in
    // the collectives below, so it uses the context's scratch (rl_ctx_scratch) and, failing that too, aborts the group
    // set-up on every rank through the communicator's own error (ncclCommAbort is not bound; the allgather then fails
    // on all ranks alike).
    char *gather = nullptr;
    double *vote = nullptr;
    bool own_bufs = cudaMalloc(&gather, (size_t)(world + 1) * sizeof(Msg)) == cudaSuccess && cudaMalloc(&vote, sizeof(double)) == cudaSuccess;
    if (!own_bufs) {
        cudaGetLastError();
        if (gather) cudaFree(gather);
        gather = nullptr;
        void *sc = nullptr;
        if (rl_ctx_scratch(ctx, (size_t)(world + 1) * sizeof(Msg) + 64, &sc) != RL_OK) return;
        gather = static_cast<char *>(sc);
        vote = reinterpret_cast<double *>(gather + (((size_t)(world + 1) * sizeof(Msg) + 15) & ~(size_t)15));
    }
    bool ok = own_bufs && cudaMalloc(&ctx->x_local, bytes) == cudaSuccess && cudaMemset(ctx->x_local, 0, bytes) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.h, ctx->x_local) == cudaSuccess;
    mine.ok = ok ? 1 : 0;
    // every rank takes part in the gather whatever happened locally, so that nobody waits for a missing peer
    std::string all((size_t)world * sizeof(Msg), '\0');
    cudaMemcpyAsync(gather + (size_t)world * sizeof(Msg), &mine, sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream);
    int r = api().allgather(gather + (size_t)world * sizeof(Msg), gather, sizeof(Msg), NCCL_INT8, ctx->nccl_comm, ctx->stream);
    cudaMemcpyAsync(&all[0], gather, (size_t)world * sizeof(Msg), cudaMemcpyDeviceToHost, ctx->stream);
    ok = ok && r == 0 && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    const Msg *msgs = reinterpret_cast<const Msg *>(all.data());
    for (int p = 0; ok && p < world; ++p) ok = msgs[p].ok == 1;
    for (int p = 0; ok && p < world; ++p) {
        void *base = ctx->x_local;
        if (p != rank) {
            ok = cudaIpcOpenMemHandle(&ctx->x_remote[p], msgs[p].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            base = ctx->x_remote[p];
        }
        if (ok) {
            ctx->x.data[p] = reinterpret_cast<double *>(base);
            ctx->x.flag[p] = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(base) + x_data_bytes(world));
        }
    }
    cudaGetLastError();
    // consensus: the mailboxes are used only if every rank mapped every peer
    {
        double host = ok ? 1.0 : 0.0;
        cudaMemcpyAsync(vote, &host, sizeof host, cudaMemcpyHostToDevice, ctx->stream);
        const int rr = api().allreduce(vote, vote, 1, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm, ctx->stream);
        cudaMemcpyAsync(&host, vote, sizeof host, cudaMemcpyDeviceToHost, ctx->stream);
        const bool synced = cudaStreamSynchronize(ctx->stream) == cudaSuccess;
        if (own_bufs) { cudaFree(vote); cudaFree(gather); }
        ok = ok && rr == 0 && synced && host == (double)world;
    }
    if (!ok) return;
    ctx->x.rank = rank;
    ctx->x.world = world;
    ctx->x.error = reinterpret_cast<int *>(reinterpret_cast<char *>(ctx->x_local) + x_data_bytes(world) + x_flag_bytes(world));
    ctx->x_seq = 0;
    ctx->x_ok = true;
}

}  // namespace

void rl_nccl_teardown(rl_ctx *ctx) {
    for (int p = 0; p < RL_X_MAX_RANKS; ++p)
        if (ctx->x_remote[p]) {
            cudaIpcCloseMemHandle(ctx->x_remote[p]);
            ctx->x_remote[p] = nullptr;
        }
    if (ctx->x_local) cudaFree(ctx->x_local);
    ctx->x_local = nullptr;
    ctx->x_ok = false;
    if (ctx->nccl_comm && api_ok()) api().destroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

// In-place f64 sum all-reduce on the context stream; no-op for a single-rank context.
rl_status rl_allreduce_f64_inplace(rl_ctx *ctx, double *buf_dev, size_t n) {
    if (ctx->world <= 1 || n == 0) return RL_OK;
    if (!ctx->nccl_comm) return rl_fail(ctx, RL_ERR_NCCL, "all-reduce requested but the communicator is not initialised");
    int r = api().allreduce(buf_dev, buf_dev, n, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm, ctx->stream);
    if (r != 0) return rl_fail(ctx, RL_ERR_NCCL, "ncclAllReduce failed: %s", api().errstr(r));
    return RL_OK;
}

extern "C" {

rl_status rl_nccl_unique_id(void *out_id128) {
    if (!out_id128) return rl_fail(nullptr, RL_ERR_INVALID_ARG, "rl_nccl_unique_id: NULL argument");
    if (!api_ok()) return rl_fail(nullptr, RL_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
    nccl_uid id;
    int r = api().get_uid(&id);
    if (r != 0) return rl_fail(nullptr, RL_ERR_NCCL, "ncclGetUniqueId failed: %s", api().errstr(r));
    memcpy(out_id128, &id, sizeof id);
    return RL_OK;
}

rl_status rl_ctx_comm_init(rl_ctx *ctx, const void *unique_id128, int32_t rank, int32_t world_size) {
    RL_REQUIRE(ctx, ctx && unique_id128, "rl_ctx_comm_init: NULL argument");
    RL_REQUIRE(ctx, world_size >= 1 && rank >= 0 && rank < world_size, "rl_ctx_comm_init: bad rank/world_size");
    if (world_size == 1) {
        ctx->rank = 0;
        ctx->world = 1;
        return RL_OK;
    }
    if (!api_ok()) return rl_fail(ctx, RL_ERR_NCCL, "libnccl.so.2 could not be loaded");
    RL_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_uid id;
    memcpy(&id, unique_id128, sizeof id);
    void *comm = nullptr;
    int r = api().init_rank(&comm, world_size, id, rank);
    if (r != 0) return rl_fail(ctx, RL_ERR_NCCL, "ncclCommInitRank failed: %s", api().errstr(r));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world_size;
    x_setup(ctx);
    return RL_OK;
}

rl_status rl_ctx_comm_info(rl_ctx *ctx, int32_t *rank, int32_t *world_size) {
    RL_REQUIRE(ctx, ctx, "ctx is NULL");
    if (rank) *rank = ctx->rank;
    if (world_size) *world_size = ctx->world;
    return RL_OK;
}

rl_status rl_ctx_comm_peer_info(rl_ctx *ctx, int32_t *peer_mailboxes, int32_t *timed_out) {
    RL_REQUIRE(ctx, ctx, "ctx is NULL");
    if (peer_mailboxes) *peer_mailboxes = ctx->x_ok ? 1 : 0;
    if (timed_out) {
        *timed_out = 0;
        if (ctx->x_ok) {
            RL_CUDA(ctx, cudaSetDevice(ctx->device));
            RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            int flag = 0;
            RL_CUDA(ctx, cudaMemcpy(&flag, ctx->x.error, sizeof flag, cudaMemcpyDeviceToHost));
            *timed_out = flag;
        }
    }
    return RL_OK;
}

rl_status rl_ctx_allreduce_f64(rl_ctx *ctx, double *buf_dev, size_t n) {
    RL_REQUIRE(ctx, ctx && buf_dev, "rl_ctx_allreduce_f64: NULL argument");
    return rl_allreduce_f64_inplace(ctx, buf_dev, n);
}

}  // extern "C"
