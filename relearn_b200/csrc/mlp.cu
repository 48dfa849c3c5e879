// mlp.cu -- Mlp module handle (src/torch/modules/ff/mlp.rs) and a plain batched forward.
// The hot loops do not call this kernel: the rollout fuses the policy forward into the step kernel
// (rollout.cu) and the updates fuse forward/backward into their own passes (update.cu).  This is the
// Module::forward entry point used for evaluation and parity checks.
#include "handles.cuh"

namespace {

// One thread per sample; weights staged in shared memory and read as warp-wide broadcasts.
template <int FT, int OT>
__global__ void __launch_bounds__(256) mlp_forward_kernel(MlpView m, const float *__restrict__ x, uint64_t n,
                                                         float *__restrict__ out) {
    extern __shared__ float sw[];
    const uint64_t np = m.n_params;
    for (uint64_t i = threadIdx.x; i < np; i += blockDim.x) sw[i] = m.params[i];
    __syncthreads();
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int F = m.in_dim, H = m.hidden, O = m.out_dim;
    const float *w1 = sw, *b1 = w1 + (size_t)H * F, *w2 = b1 + H, *b2 = w2 + (size_t)O * H;
    float xi[FT], z[OT];
#pragma unroll
    for (int f = 0; f < FT; ++f) xi[f] = f < F ? x[(uint64_t)f * n + s] : 0.0f;
    if (m.n_hidden > 1) {
        rl_mlp_eval_deep(m, sw, xi, z);
        for (int k = 0; k < O; ++k) out[(uint64_t)k * n + s] = z[k];
        return;
    }
#pragma unroll
    for (int k = 0; k < OT; ++k) z[k] = k < O ? b2[k] : 0.0f;
    for (int j = 0; j < H; ++j) {
        float acc = b1[j];
#pragma unroll
        for (int f = 0; f < FT; ++f)
            if (f < F) acc = fmaf(w1[j * F + f], xi[f], acc);
        const float h = rl_activate(m.act, acc);
#pragma unroll
        for (int k = 0; k < OT; ++k)
            if (k < O) z[k] = fmaf(w2[k * H + j], h, z[k]);
    }
#pragma unroll
    for (int k = 0; k < OT; ++k)
        if (k < O) out[(uint64_t)k * n + s] = z[k];
}

}  // namespace

extern "C" {

rl_status rl_mlp_create(rl_ctx *ctx, int32_t in_dim, const int32_t *hidden_sizes, int32_t n_hidden, int32_t out_dim,
                        rl_activation activation, rl_mlp **out) {
    RL_REQUIRE(ctx, ctx && out && hidden_sizes, "rl_mlp_create: NULL argument");
    if (n_hidden < 1 || n_hidden > RL_MLP_MAX_HIDDEN_LAYERS)
        return rl_fail(ctx, RL_ERR_UNSUPPORTED, "rl_mlp_create: one to %d hidden layers are implemented (MlpConfig::hidden_sizes, got %d)",
                       RL_MLP_MAX_HIDDEN_LAYERS, n_hidden);
    RL_REQUIRE(ctx, in_dim >= 1 && in_dim <= 36 && out_dim >= 1 && out_dim <= 32, "rl_mlp_create: dims out of range");
    for (int l = 0; l < n_hidden; ++l)
        RL_REQUIRE(ctx, hidden_sizes[l] >= 1 && hidden_sizes[l] <= (n_hidden == 1 ? 1024 : RL_DEEP_MAXH),
                   "rl_mlp_create: hidden size out of range (<= 1024 with one hidden layer, <= 256 with two or three)");
    RL_CUDA(ctx, cudaSetDevice(ctx->device));
    rl_mlp *m = new (std::nothrow) rl_mlp();
    if (!m) return rl_fail(ctx, RL_ERR_OOM, "rl_mlp_create: host allocation failed");
    m->ctx = ctx;
    m->in_dim = in_dim;
    m->hidden = hidden_sizes[0];
    m->n_hidden = n_hidden;
    for (int l = 0; l < n_hidden; ++l) m->hid[l] = hidden_sizes[l];
    m->out_dim = out_dim;
    m->act = activation;
    m->n_params = rl_mlp::count_layers(in_dim, m->hid, n_hidden, out_dim);
    if (n_hidden > 1 && m->n_params > 48 * 1024) {  // the layer-generic kernels stage the parameters in shared memory
        delete m;
        return rl_fail(ctx, RL_ERR_UNSUPPORTED, "rl_mlp_create: modules with two or three hidden layers are built for <= 48 K parameters");
    }
    cudaError_t e = cudaMalloc((void **)&m->params, m->n_params * sizeof(float));
    if (e != cudaSuccess) {
        delete m;
        return rl_fail(ctx, RL_ERR_OOM, "rl_mlp_create: %s", cudaGetErrorString(e));
    }
    cudaMemsetAsync(m->params, 0, m->n_params * sizeof(float), ctx->stream);
    *out = m;
    return RL_OK;
}

rl_status rl_mlp_destroy(rl_mlp *m) {
    if (!m) return RL_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->params);
    delete m;
    return RL_OK;
}

rl_status rl_mlp_num_params(rl_mlp *m, uint64_t *n) {
    if (!m || !n) return rl_fail(m ? m->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_mlp_num_params: NULL argument");
    *n = m->n_params;
    return RL_OK;
}

rl_status rl_mlp_set_weights(rl_mlp *m, const float *host, uint64_t n) {
    if (!m || !host) return rl_fail(m ? m->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_mlp_set_weights: NULL argument");
    RL_REQUIRE(m->ctx, n == m->n_params, "rl_mlp_set_weights: wrong parameter count");
    RL_CUDA(m->ctx, cudaMemcpyAsync(m->params, host, n * sizeof(float), cudaMemcpyHostToDevice, m->ctx->stream));
    RL_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
    return RL_OK;
}

rl_status rl_mlp_set_weights_async(rl_mlp *m, const float *pinned_host, uint64_t n) {
    if (!m || !pinned_host) return rl_fail(m ? m->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_mlp_set_weights_async: NULL argument");
    RL_REQUIRE(m->ctx, n == m->n_params, "rl_mlp_set_weights_async: wrong parameter count");
    cudaPointerAttributes attr{};
    const cudaError_t e = cudaPointerGetAttributes(&attr, pinned_host);
    if (e != cudaSuccess || attr.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        return rl_fail(m->ctx, RL_ERR_INVALID_ARG, "rl_mlp_set_weights_async: the source must be page-locked host memory (rl_malloc_host)");
    }
    RL_CUDA(m->ctx, cudaMemcpyAsync(m->params, pinned_host, n * sizeof(float), cudaMemcpyHostToDevice, m->ctx->stream));
    return RL_OK;
}

rl_status rl_mlp_get_weights(rl_mlp *m, float *host, uint64_t n) {
    if (!m || !host) return rl_fail(m ? m->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_mlp_get_weights: NULL argument");
    RL_REQUIRE(m->ctx, n == m->n_params, "rl_mlp_get_weights: wrong parameter count");
    RL_CUDA(m->ctx, cudaMemcpyAsync(host, m->params, n * sizeof(float), cudaMemcpyDeviceToHost, m->ctx->stream));
    RL_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
    return RL_OK;
}

rl_status rl_mlp_forward(rl_mlp *m, const float *x_dev, uint64_t n, float *out_dev) {
    if (!m || !x_dev || !out_dev) return rl_fail(m ? m->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_mlp_forward: NULL argument");
    rl_ctx *ctx = m->ctx;
    if (n == 0) return RL_OK;
    const size_t smem = m->n_params * sizeof(float);
    const unsigned block = 256, grid = rl_grid_for(n, block);
    MlpView v = rl_mlp_view(m);
    if (m->in_dim <= 8 && m->out_dim <= 2) {
        RL_CUDA(ctx, cudaFuncSetAttribute(mlp_forward_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RL_LAUNCH(ctx, (mlp_forward_kernel<8, 2>), grid, block, smem, v, x_dev, n, out_dev);
    } else {
        RL_CUDA(ctx, cudaFuncSetAttribute(mlp_forward_kernel<36, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RL_LAUNCH(ctx, (mlp_forward_kernel<36, 32>), grid, block, smem, v, x_dev, n, out_dev);
    }
    return RL_OK;
}

}  // extern "C"
