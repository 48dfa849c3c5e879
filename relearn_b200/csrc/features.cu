// features.cu -- FeatureSpace::batch_features (src/spaces/mod.rs:329-412) for the concrete spaces
// the hot path uses.  Output is the reference layout: f32 [n][num_features] row-major.
#include "handles.cuh"

namespace {

__global__ void encode_interval_kernel(const double *__restrict__ x, uint64_t n, float *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)x[i];  // interval.rs:101-117: [x as f32]
}

// one thread per output element so stores are coalesced
__global__ void encode_onehot_kernel(const int64_t *__restrict__ idx, uint64_t n, uint64_t size, int option,
                                     float *__restrict__ out) {
    const uint64_t width = size + (option ? 1 : 0);
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * width) return;
    const uint64_t row = k / width, col = k - row * width;
    const int64_t v = idx[row];
    float o;
    if (option) {
        // option.rs:88-116: [1, 0...] for None, [0, features(x)...] for Some(x)
        if (col == 0) o = v < 0 ? 1.0f : 0.0f;
        else o = (v >= 0 && (uint64_t)v == col - 1) ? 1.0f : 0.0f;
    } else {
        o = (v >= 0 && (uint64_t)v == col) ? 1.0f : 0.0f;  // index.rs:97-138, boolean.rs:125-139 (size 1: [v])
    }
    out[k] = o;
}

__global__ void encode_boolean_kernel(const int64_t *__restrict__ b, uint64_t n, float *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = b[i] ? 1.0f : 0.0f;
}

}  // namespace

extern "C" rl_status rl_encode_features(rl_ctx *ctx, rl_space_kind kind, uint64_t size, const void *elems_dev,
                                        uint64_t n, float *out_dev) {
    RL_REQUIRE(ctx, ctx && (n == 0 || (elems_dev && out_dev)), "rl_encode_features: NULL argument");
    if (n == 0) return RL_OK;
    const unsigned block = 256;
    switch (kind) {
    case RL_SPACE_INTERVAL:
        RL_LAUNCH(ctx, encode_interval_kernel, rl_grid_for(n, block), block, 0, (const double *)elems_dev, n, out_dev);
        return RL_OK;
    case RL_SPACE_BOOLEAN:
        RL_LAUNCH(ctx, encode_boolean_kernel, rl_grid_for(n, block), block, 0, (const int64_t *)elems_dev, n, out_dev);
        return RL_OK;
    case RL_SPACE_INDEX:
        RL_REQUIRE(ctx, size > 0, "rl_encode_features: index space is empty");
        RL_LAUNCH(ctx, encode_onehot_kernel, rl_grid_for(n * size, block), block, 0, (const int64_t *)elems_dev, n, size, 0,
                  out_dev);
        return RL_OK;
    case RL_SPACE_OPTION_INDEX:
        RL_LAUNCH(ctx, encode_onehot_kernel, rl_grid_for(n * (size + 1), block), block, 0, (const int64_t *)elems_dev, n,
                  size, 1, out_dev);
        return RL_OK;
    }
    return rl_fail(ctx, RL_ERR_INVALID_ARG, "rl_encode_features: unknown space kind");
}
