// nccl.cu -- data-parallel group.  relearn has no collectives; this replaces the crossbeam thread
// fan-out of src/simulation/train.rs:124-158 across GPUs.  Rollouts shard by lane and need no
// exchange; the update kernels all-reduce their f64 partial sums (gradient, Fisher-vector product,
// loss/KL scalars) through rl_allreduce_f64_inplace().
//
// NCCL is bound at run time (dlopen) so the library has no link-time dependency on a particular
// libnccl and shares the copy torch already loaded when the host process uses torch.distributed.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(void **, int, nccl_uid, int);
typedef int (*fn_destroy)(void *);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_errstr)(int);

struct NcclApi {
    void *handle = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr;
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
constexpr int NCCL_SUM = 0;      // ncclRedOp_t ncclSum

}  // namespace

void rl_nccl_teardown(rl_ctx *ctx) {
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
    return RL_OK;
}

rl_status rl_ctx_comm_info(rl_ctx *ctx, int32_t *rank, int32_t *world_size) {
    RL_REQUIRE(ctx, ctx, "ctx is NULL");
    if (rank) *rank = ctx->rank;
    if (world_size) *world_size = ctx->world;
    return RL_OK;
}

rl_status rl_ctx_allreduce_f64(rl_ctx *ctx, double *buf_dev, size_t n) {
    RL_REQUIRE(ctx, ctx && buf_dev, "rl_ctx_allreduce_f64: NULL argument");
    return rl_allreduce_f64_inplace(ctx, buf_dev, n);
}

}  // extern "C"
