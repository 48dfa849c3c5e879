// features.cu -- FeatureSpace::batch_features (src/spaces/mod.rs:329-412) for the concrete spaces
// the hot path uses (output in the reference layout: f32 [n][num_features] row-major), and LazyHistoryFeatures'
// packed tensors (src/torch/agents/features.rs) built from a device trajectory.
#include <cub/cub.cuh>

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

// ---- LazyHistoryFeatures (src/torch/agents/features.rs:70-215) in the packed order of PackedStructure (packed.rs:346-420) ----
// Episodes of a lane are the runs of stored steps that end with a non-Continue successor (the trajectory is already
// finalised: buffers/mod.rs:237-261), listed lane after lane like the reference walks its buffers.

// pass 0 (fill == false): episodes per lane; pass 1: (lane, start, sort key) of every episode at base[lane] + k
__global__ void __launch_bounds__(128)
    pack_episodes_kernel(const uint8_t *__restrict__ succ, uint64_t T, uint64_t E, bool fill, uint32_t *__restrict__ count,
                         const uint32_t *__restrict__ base, uint32_t *__restrict__ ep_lane, uint32_t *__restrict__ ep_start,
                         uint32_t *__restrict__ key, uint32_t *__restrict__ val) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    uint32_t n = 0, start = 0;
    for (uint64_t t = 0; t < T; ++t) {
        const uint8_t sc = succ[t * E + e];
        if (sc == RL_PAD) break;
        if (sc != RL_CONTINUE) {
            if (fill) {
                const uint32_t k = base[e] + n;
                ep_lane[k] = (uint32_t)e;
                ep_start[k] = start;
                key[k] = (uint32_t)T - ((uint32_t)t + 1u - start);  // ascending key = descending length
                val[k] = k;
            }
            n += 1;
            start = (uint32_t)t + 1u;
        }
    }
    if (!fill) count[e] = n;
}

// batch_sizes[t] = episodes longer than t (the sorted keys are T - len, ascending); ext: episodes of len + 1 > t
__global__ void pack_batch_sizes_kernel(const uint32_t *__restrict__ key_sorted, uint32_t n_eps, uint32_t T,
                                        unsigned long long *__restrict__ batch, unsigned long long *__restrict__ ext_batch) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > T) return;
    auto longer_than = [&](uint32_t limit) {  // #episodes with len > limit  <=>  key < T - limit
        if (limit >= T) return 0u;
        const uint32_t bound = T - limit;
        uint32_t lo = 0, hi = n_eps;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (key_sorted[mid] < bound) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    if (t < T) batch[t] = longer_than(t);
    ext_batch[t] = t == 0 ? n_eps : longer_than(t - 1);
}

// The [T][F][E] planes hold a slot's bytes in F + 3 different rows; gathering them by destination would read F + 3
// sectors per slot (measured: 7.9 ms for 33 M steps; scattering them in source order instead: 11.6 ms).  So the stored
// slots are first copied into records [T][E][RW words] = (F observation features, reward, action | successor << 8) with
// coalesced loads and stores, and the gather reads one or two sectors per slot (0.73 + 2.06 ms).  Routing the rows
// through shared memory for fully coalesced stores was measured slower (2.9 ms: the occupancy it costs hides less of the
// scattered reads' latency).
__host__ __device__ inline int pack_record_words(int F) { return (F + 2 + 3) / 4 * 4; }

__global__ void __launch_bounds__(256)
    pack_records_kernel(const float *__restrict__ obs, const uint8_t *__restrict__ action, const float *__restrict__ reward,
                        const uint8_t *__restrict__ succ, uint64_t T, uint64_t E, int F, float *__restrict__ rec) {
    const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= T * E) return;
    const uint8_t sc = succ[n];
    if (sc == RL_PAD) return;
    const uint64_t t = n / E, e = n - t * E;
    const int RW = pack_record_words(F);
    float *r = rec + n * RW;
    for (int f = 0; f < F; ++f) r[f] = obs[(t * F + f) * E + e];
    r[F] = reward[n];
    r[F + 1] = __uint_as_float((uint32_t)action[n] | ((uint32_t)sc << 8));
}

__global__ void __launch_bounds__(256)
    pack_gather_kernel(const uint32_t *__restrict__ key_sorted, const uint32_t *__restrict__ order, const uint32_t *__restrict__ ep_lane,
                       const uint32_t *__restrict__ ep_start, uint32_t n_eps, uint32_t T, uint64_t E, int F,
                       const unsigned long long *__restrict__ off, const unsigned long long *__restrict__ ext_off,
                       const float *__restrict__ rec, const float *__restrict__ next_obs, float *__restrict__ out_obs,
                       float *__restrict__ out_ext, uint8_t *__restrict__ out_invalid, long long *__restrict__ out_action,
                       float *__restrict__ out_reward) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    if (r >= n_eps) return;
    const uint32_t len = T - key_sorted[r];
    if (t > len) return;
    const uint32_t k = order[r];
    const uint64_t lane = ep_lane[k], step = (uint64_t)ep_start[k] + t;
    const int RW = pack_record_words(F);
    if (t < len) {
        const uint64_t dst = off[t] + r, edst = ext_off[t] + r;
        const float4 *src = reinterpret_cast<const float4 *>(rec + (step * E + lane) * RW);
        for (int q = 0; q < RW / 4; ++q) {
            const float4 v = __ldg(src + q);
            const float w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int f = 4 * q + c;
                if (f < F) {
                    if (out_obs) out_obs[dst * F + f] = w[c];
                    if (out_ext) out_ext[edst * F + f] = w[c];
                } else if (f == F) {
                    if (out_reward) out_reward[dst] = w[c];
                } else if (f == F + 1) {
                    if (out_action) out_action[dst] = (long long)(__float_as_uint(w[c]) & 0xffu);
                }
            }
        }
        if (out_invalid) out_invalid[edst] = 0;
    } else if (out_ext || out_invalid) {
        // ExtendedEpisodeObservations (features.rs:217-262): after the steps' observations the successor observation --
        // the stored next observation on Interrupt, none (invalid, zeros) on Terminate
        const uint64_t dst = ext_off[t] + r, last = step - 1;
        const uint32_t meta = __float_as_uint(rec[(last * E + lane) * RW + F + 1]);
        const bool invalid = (meta >> 8) != RL_INTERRUPT;
        if (out_ext)
            for (int f = 0; f < F; ++f) out_ext[dst * F + f] = invalid ? 0.0f : next_obs[(last * F + f) * E + lane];
        if (out_invalid) out_invalid[dst] = invalid ? 1 : 0;
    }
}

}  // namespace

extern "C" rl_status rl_pack_history(rl_traj *traj, float *obs_dev, float *ext_obs_dev, uint8_t *ext_invalid_dev, int64_t *action_dev,
                                     float *reward_dev, int64_t *batch_sizes_dev, int64_t *ext_batch_sizes_dev, rl_packed_info *info) {
    if (!traj || !info) return rl_fail(traj ? traj->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_pack_history: NULL argument");
    rl_ctx *ctx = traj->ctx;
    const uint64_t T = traj->used_T ? traj->used_T : traj->T, E = traj->E;
    RL_REQUIRE(ctx, T < (1ull << 31) && E < (1ull << 32) && T * E < (1ull << 32), "rl_pack_history: trajectory too large");
    // scratch: count | base [E + 1] | ep_lane, ep_start, key, val, key_sorted, order [T E each] | batch, ext_batch, off, ext_off [T + 1 each]
    // | records [T E][RW]
    const size_t cap = (size_t)T * E;
    size_t sort_bytes = 0, scan_bytes = 0, scan2_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (int)cap, 0, 32, ctx->stream);
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)E, ctx->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, scan2_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)T + 1, ctx->stream);
    const size_t tmp_bytes = (std::max(sort_bytes, std::max(scan_bytes, scan2_bytes)) + 255) / 256 * 256;
    const size_t u32_words = 2 * (E + 1) + 6 * cap;
    const size_t rec_bytes = (cap * pack_record_words((int)traj->F) * sizeof(float) + 255) / 256 * 256;
    const size_t bytes = tmp_bytes + (u32_words * 4 + 255) / 256 * 256 + 4 * (T + 1) * sizeof(unsigned long long) + 256 + rec_bytes;
    unsigned char *scratch;
    RL_TRY(rl_ctx_scratch(ctx, bytes, (void **)&scratch));
    void *tmp = scratch;
    uint32_t *count = reinterpret_cast<uint32_t *>(scratch + tmp_bytes), *base = count + (E + 1);
    uint32_t *ep_lane = base + (E + 1), *ep_start = ep_lane + cap, *key = ep_start + cap, *val = key + cap;
    uint32_t *key_sorted = val + cap, *order = key_sorted + cap;
    unsigned long long *batch = reinterpret_cast<unsigned long long *>(scratch + tmp_bytes + (u32_words * 4 + 255) / 256 * 256);
    unsigned long long *ext_batch = batch + (T + 1), *off = ext_batch + (T + 1), *ext_off = off + (T + 1);
    float *rec = reinterpret_cast<float *>(scratch + (bytes - rec_bytes) / 256 * 256);  // 16-byte aligned records

    RL_LAUNCH(ctx, pack_episodes_kernel, rl_grid_for(E, 128), 128, 0, traj->succ, T, E, false, count, base, ep_lane, ep_start, key, val);
    RL_CUDA(ctx, cudaMemsetAsync(base, 0, sizeof(uint32_t), ctx->stream));
    size_t tb = tmp_bytes;
    RL_CUDA(ctx, cub::DeviceScan::InclusiveSum(tmp, tb, count, base + 1, (int)E, ctx->stream));
    uint32_t n_eps = 0;
    RL_CUDA(ctx, cudaMemcpyAsync(&n_eps, base + E, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RL_LAUNCH(ctx, pack_episodes_kernel, rl_grid_for(E, 128), 128, 0, traj->succ, T, E, true, count, base, ep_lane, ep_start, key, val);
    int bits = 1;
    while ((1ull << bits) <= T) ++bits;
    if (n_eps > 0) {
        tb = tmp_bytes;
        RL_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, tb, key, key_sorted, val, order, (int)n_eps, 0, bits, ctx->stream));  // stable
    }
    RL_LAUNCH(ctx, pack_batch_sizes_kernel, rl_grid_for(T + 1, 256), 256, 0, key_sorted, n_eps, (uint32_t)T, batch, ext_batch);
    tb = tmp_bytes;
    RL_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tb, batch, off, (int)T, ctx->stream));
    tb = tmp_bytes;
    RL_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tb, ext_batch, ext_off, (int)T + 1, ctx->stream));
    uint32_t max_len = 0;
    if (n_eps > 0) {
        uint32_t k0 = 0;
        RL_CUDA(ctx, cudaMemcpyAsync(&k0, key_sorted, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        max_len = (uint32_t)T - k0;
        RL_LAUNCH(ctx, pack_records_kernel, rl_grid_for(T * E, 256), 256, 0, traj->obs, traj->action, traj->reward, traj->succ, T, E,
                  (int)traj->F, rec);
        RL_LAUNCH(ctx, pack_gather_kernel, dim3(rl_grid_for(n_eps, 256), max_len + 1), 256, 0, key_sorted, order, ep_lane, ep_start, n_eps,
                  (uint32_t)T, E, (int)traj->F, off, ext_off, rec, traj->next_obs, obs_dev, ext_obs_dev, ext_invalid_dev,
                  (long long *)action_dev, reward_dev);
    }
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "batch sizes are copied as 64-bit words");
    if (batch_sizes_dev && max_len > 0)
        RL_CUDA(ctx, cudaMemcpyAsync(batch_sizes_dev, batch, (size_t)max_len * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    if (ext_batch_sizes_dev && n_eps > 0)
        RL_CUDA(ctx, cudaMemcpyAsync(ext_batch_sizes_dev, ext_batch, ((size_t)max_len + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice,
                                     ctx->stream));
    unsigned long long n_steps = 0;
    if (n_eps > 0) {
        unsigned long long last_off = 0, last_batch = 0;
        RL_CUDA(ctx, cudaMemcpyAsync(&last_off, off + (max_len - 1), sizeof(last_off), cudaMemcpyDeviceToHost, ctx->stream));
        RL_CUDA(ctx, cudaMemcpyAsync(&last_batch, batch + (max_len - 1), sizeof(last_batch), cudaMemcpyDeviceToHost, ctx->stream));
        RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        n_steps = last_off + last_batch;
    }
    info->num_steps = n_steps;
    info->num_episodes = n_eps;
    info->max_len = max_len;
    return RL_OK;
}

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
