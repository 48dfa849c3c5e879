// scan.cu -- reverse discounted scans over trajectories.
//
// PackedTensor::discounted_cumsum_from_end (src/torch/packed.rs:280-342) is a CPU ndarray loop in the
// reference (GPU -> CPU -> GPU round trip per call).  Here trajectories are time-major [T][E], so a
// thread owns one lane, walks it from the end, and adjacent threads touch adjacent addresses: every
// load/store of a warp is one coalesced 128 B (f32) or 32 B (u8) segment.  The recurrence is
// evaluated exactly as the reference does -- y_t = x_t + (y_{t+1} * d), multiply then add, f32 --
// so results are bit-identical to the sequential loop.
//
// K3  gae_scan_kernel: HBM-bound, 17 algorithmic bytes per step (reward 4 + V 4 + succ 1 in,
//     advantage 4 + reward-to-go 4 out; + 4 for V(next) on interrupted steps).
#include "handles.cuh"

namespace {

constexpr int SCAN_UNROLL = 8;

// y[t][e] = x[t][e] + d * y[t+1][e]; restart after steps whose successor is not Continue; PAD -> 0.
__global__ void __launch_bounds__(128) cumsum_kernel(const float *__restrict__ x, const uint8_t *__restrict__ succ,
                                                     uint64_t T, uint64_t E, float d, float *__restrict__ y) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float carry = 0.0f;
    int64_t t = (int64_t)T - 1;
    // batches of independent loads first, then the dependent recurrence
    for (; t >= SCAN_UNROLL - 1; t -= SCAN_UNROLL) {
        float xv[SCAN_UNROLL];
        uint8_t sv[SCAN_UNROLL];
#pragma unroll
        for (int k = 0; k < SCAN_UNROLL; ++k) {
            xv[k] = __ldg(x + (uint64_t)(t - k) * E + e);
            sv[k] = __ldg(succ + (uint64_t)(t - k) * E + e);
        }
#pragma unroll
        for (int k = 0; k < SCAN_UNROLL; ++k) {
            if (sv[k] != RL_CONTINUE) carry = 0.0f;
            carry = __fadd_rn(xv[k], __fmul_rn(carry, d));  // packed.rs:336  *a += *b * discount
            if (sv[k] == RL_PAD) carry = 0.0f;
            y[(uint64_t)(t - k) * E + e] = carry;
        }
    }
    for (; t >= 0; --t) {
        const float xv = x[(uint64_t)t * E + e];
        const uint8_t sv = succ[(uint64_t)t * E + e];
        if (sv != RL_CONTINUE) carry = 0.0f;
        carry = __fadd_rn(xv, __fmul_rn(carry, d));
        if (sv == RL_PAD) carry = 0.0f;
        y[(uint64_t)t * E + e] = carry;
    }
}

// Packed (time-major ragged, longest first) in-place scan: thread i owns sequence i.
__global__ void packed_cumsum_kernel(float *__restrict__ x, const uint64_t *__restrict__ offsets,
                                     const uint64_t *__restrict__ batch_sizes, uint64_t n_batches, float d) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_batches == 0 || i >= batch_sizes[0]) return;
    uint64_t len = 0;
    while (len < n_batches && batch_sizes[len] > i) ++len;
    float carry = 0.0f;
    for (int64_t t = (int64_t)len - 1; t >= 0; --t) {
        carry = __fadd_rn(x[offsets[t] + i], __fmul_rn(carry, d));
        x[offsets[t] + i] = carry;
    }
}

// State values of every stored observation, and of the successor observation on interrupted steps
// (eval_extended_state_values, critics/mod.rs:116-131).  One thread per (t, lane) slot.
__global__ void __launch_bounds__(256) value_forward_kernel(MlpView m, const float *__restrict__ obs,
                                                           const float *__restrict__ next_obs,
                                                           const uint8_t *__restrict__ succ, uint64_t T, uint64_t E,
                                                           float *__restrict__ v, float *__restrict__ v_next) {
    extern __shared__ float sw[];
    const uint64_t np = m.n_params;
    for (uint64_t i = threadIdx.x; i < np; i += blockDim.x) sw[i] = m.params[i];
    __syncthreads();
    const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= T * E) return;
    const uint8_t sc = succ[n];
    if (sc == RL_PAD) {
        v[n] = 0.0f;
        return;
    }
    const uint64_t t = n / E, e = n - t * E;
    const int F = m.in_dim, H = m.hidden;
    const float *w1 = sw, *b1 = w1 + (size_t)H * F, *w2 = b1 + H, *b2 = w2 + H;
    constexpr int FT = 36;
    float xi[FT], xn[FT];
    const bool intr = sc == RL_INTERRUPT;
#pragma unroll
    for (int f = 0; f < FT; ++f) {
        xi[f] = f < F ? obs[(t * F + f) * E + e] : 0.0f;
        xn[f] = (f < F && intr) ? next_obs[(t * F + f) * E + e] : 0.0f;
    }
    if (m.n_hidden > 1) {  // MlpConfig::hidden_sizes with two or three entries
        float zd;
        rl_mlp_eval_deep(m, sw, xi, &zd);
        v[n] = zd;
        if (intr) {
            rl_mlp_eval_deep(m, sw, xn, &zd);
            v_next[n] = zd;
        }
        return;
    }
    float z = b2[0], zn = b2[0];
    for (int j = 0; j < H; ++j) {
        float acc = b1[j], accn = b1[j];
#pragma unroll
        for (int f = 0; f < FT; ++f)
            if (f < F) {
                const float w = w1[j * F + f];
                acc = fmaf(w, xi[f], acc);
                accn = fmaf(w, xn[f], accn);
            }
        z = fmaf(w2[j], rl_activate(m.act, acc), z);
        zn = fmaf(w2[j], rl_activate(m.act, accn), zn);
    }
    v[n] = z;
    if (intr) v_next[n] = zn;
}


// Same map for the default critic (5 -> 128 -> 1 ReLU, MlpConfig defaults): thread per slot, hidden units as
// pairs (q, q + 64) with packed FP32, weights in shared memory as float4 planes read by warp-wide broadcast:
//   plane 0: w1[.][0] pair, w1[.][1] pair   plane 1: w1[.][2] pair, w1[.][3] pair
//   plane 2: w1[.][4] pair, b1 pair         plane 3: w2 pair, 0, 0
// 4 LDS.128 + 6 FFMA2 + 2 FMNMX per pair instead of ~40 scalar instructions in the generic kernel.
__global__ void __launch_bounds__(256) value_forward_pairs_kernel(MlpView m, const float *__restrict__ obs,
                                                                 const float *__restrict__ next_obs,
                                                                 const uint8_t *__restrict__ succ, uint64_t T, uint64_t E,
                                                                 float *__restrict__ v, float *__restrict__ v_next) {
    constexpr int H = 128, NPAIR = 64;
    __shared__ float4 sw4[4 * NPAIR];
    __shared__ float sb2;
    const int F = m.in_dim;
    const float *w1 = m.w1(), *b1 = m.b1(), *w2 = m.w2();
    for (int i = threadIdx.x; i < 4 * NPAIR; i += blockDim.x) {
        const int c = i / NPAIR, q = i - c * NPAIR, j0 = q, j1 = q + NPAIR;
        auto W1 = [&](int j, int f) { return f < F ? w1[j * F + f] : 0.0f; };
        float4 val;
        if (c == 0) val = make_float4(W1(j0, 0), W1(j1, 0), W1(j0, 1), W1(j1, 1));
        else if (c == 1) val = make_float4(W1(j0, 2), W1(j1, 2), W1(j0, 3), W1(j1, 3));
        else if (c == 2) val = make_float4(W1(j0, 4), W1(j1, 4), b1[j0], b1[j1]);
        else val = make_float4(w2[j0], w2[j1], 0.0f, 0.0f);
        sw4[i] = val;
    }
    if (threadIdx.x == 0) sb2 = m.b2()[0];
    __syncthreads();
    const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= T * E) return;
    const uint8_t sc = succ[n];
    if (sc == RL_PAD) {
        v[n] = 0.0f;
        return;
    }
    const uint64_t t = n / E, e = n - t * E;
    auto forward = [&](const float *src) {
        float x[5];
#pragma unroll
        for (int f = 0; f < 5; ++f) x[f] = f < F ? __ldg(src + (t * F + f) * E + e) : 0.0f;
        const float2 o0 = make_float2(x[0], x[0]), o1 = make_float2(x[1], x[1]), o2 = make_float2(x[2], x[2]);
        const float2 o3 = make_float2(x[3], x[3]), o4 = make_float2(x[4], x[4]);
        float2 za = make_float2(0.0f, 0.0f), zb = make_float2(0.0f, 0.0f);
#pragma unroll 8
        for (int q = 0; q < NPAIR; ++q) {
            const float4 A = sw4[q], B = sw4[NPAIR + q], Cw = sw4[2 * NPAIR + q], D = sw4[3 * NPAIR + q];
            float2 pre = make_float2(Cw.z, Cw.w);
            pre = __ffma2_rn(make_float2(A.x, A.y), o0, pre);
            pre = __ffma2_rn(make_float2(A.z, A.w), o1, pre);
            pre = __ffma2_rn(make_float2(B.x, B.y), o2, pre);
            pre = __ffma2_rn(make_float2(B.z, B.w), o3, pre);
            pre = __ffma2_rn(make_float2(Cw.x, Cw.y), o4, pre);
            const float2 h = make_float2(fmaxf(pre.x, 0.0f), fmaxf(pre.y, 0.0f));
            if (q & 1) zb = __ffma2_rn(make_float2(D.x, D.y), h, zb);
            else za = __ffma2_rn(make_float2(D.x, D.y), h, za);
        }
        return (za.x + za.y) + (zb.x + zb.y) + sb2;
    };
    v[n] = forward(obs);
    if (sc == RL_INTERRUPT) v_next[n] = forward(next_obs);
}

// temporal_differences + gae + reward_to_go in one backward pass per lane
// (critics/mod.rs:101-105,158-199).  delta = (r + gamma * V_next) - V, f32, unfused.
template <bool HAS_V>
__global__ void __launch_bounds__(128)
    gae_scan_kernel(const float *__restrict__ reward, const float *__restrict__ v, const float *__restrict__ v_next,
                    const uint8_t *__restrict__ succ, uint64_t T, uint64_t E, float gamma, float gl,
                    float *__restrict__ adv, float *__restrict__ rtg) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float carry_a = 0.0f, carry_r = 0.0f, v_after = 0.0f;
    int64_t t = (int64_t)T - 1;
    for (; t >= 0; t -= SCAN_UNROLL) {
        float rv[SCAN_UNROLL], vv[SCAN_UNROLL];
        uint8_t sv[SCAN_UNROLL];
#pragma unroll
        for (int k = 0; k < SCAN_UNROLL; ++k) {
            const int64_t tt = t - k;
            const bool in = tt >= 0;
            const uint64_t idx = (uint64_t)(in ? tt : 0) * E + e;
            sv[k] = in ? __ldg(succ + idx) : (uint8_t)RL_PAD;
            rv[k] = in ? __ldg(reward + idx) : 0.0f;
            vv[k] = (HAS_V && in) ? __ldg(v + idx) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < SCAN_UNROLL; ++k) {
            const int64_t tt = t - k;
            if (tt < 0) break;
            const uint64_t idx = (uint64_t)tt * E + e;
            if (sv[k] == RL_PAD) {
                carry_a = carry_r = v_after = 0.0f;
                if (adv) adv[idx] = 0.0f;
                if (rtg) rtg[idx] = 0.0f;
                continue;
            }
            float next = v_after;  // Continue: V of the following step's observation
            if (sv[k] != RL_CONTINUE) {
                carry_a = carry_r = 0.0f;
                next = (HAS_V && sv[k] == RL_INTERRUPT) ? v_next[idx] : 0.0f;  // features.rs:139-185
            }
            const float delta = __fsub_rn(__fadd_rn(rv[k], __fmul_rn(gamma, next)), vv[k]);
            carry_a = __fadd_rn(delta, __fmul_rn(carry_a, gl));
            carry_r = __fadd_rn(rv[k], __fmul_rn(carry_r, gamma));
            if (adv) adv[idx] = carry_a;
            if (rtg) rtg[idx] = carry_r;
            v_after = vv[k];
        }
    }
}

}  // namespace

extern "C" {

rl_status rl_discounted_cumsum(rl_ctx *ctx, const float *x_dev, const uint8_t *succ_dev, uint64_t steps,
                               uint64_t lanes, float discount, float *y_dev) {
    RL_REQUIRE(ctx, ctx && x_dev && succ_dev && y_dev, "rl_discounted_cumsum: NULL argument");
    if (steps == 0 || lanes == 0) return RL_OK;
    const unsigned block = 128, grid = rl_grid_for(lanes, block);
    RL_LAUNCH(ctx, cumsum_kernel, grid, block, 0, x_dev, succ_dev, steps, lanes, discount, y_dev);
    return RL_OK;
}

rl_status rl_discounted_cumsum_packed(rl_ctx *ctx, float *x_host, uint64_t n, const uint64_t *batch_sizes,
                                      uint64_t n_batches, float discount) {
    RL_REQUIRE(ctx, ctx && x_host && batch_sizes, "rl_discounted_cumsum_packed: NULL argument");
    if (n == 0 || n_batches == 0) return RL_OK;
    uint64_t total = 0;
    for (uint64_t t = 0; t < n_batches; ++t) {
        RL_REQUIRE(ctx, t == 0 || batch_sizes[t] <= batch_sizes[t - 1], "batch sizes must be non-increasing");
        total += batch_sizes[t];
    }
    RL_REQUIRE(ctx, total == n, "batch sizes do not match array first dimension length");  // packed.rs:338-341
    const size_t bytes = n * sizeof(float) + 2 * n_batches * sizeof(uint64_t);
    char *buf;
    RL_TRY(rl_ctx_scratch(ctx, bytes + 64, (void **)&buf));
    uint64_t *d_off = (uint64_t *)buf;
    uint64_t *d_bs = d_off + n_batches;
    float *d_x = (float *)(d_bs + n_batches);
    uint64_t *h;
    RL_TRY(rl_ctx_pinned(ctx, 2 * n_batches * sizeof(uint64_t), (void **)&h));
    uint64_t off = 0;
    for (uint64_t t = 0; t < n_batches; ++t) {
        h[t] = off;
        h[n_batches + t] = batch_sizes[t];
        off += batch_sizes[t];
    }
    RL_CUDA(ctx, cudaMemcpyAsync(d_off, h, 2 * n_batches * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    RL_CUDA(ctx, cudaMemcpyAsync(d_x, x_host, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned block = 128, grid = rl_grid_for(batch_sizes[0], block);
    RL_LAUNCH(ctx, packed_cumsum_kernel, grid, block, 0, d_x, d_off, d_bs, n_batches, discount);
    RL_CUDA(ctx, cudaMemcpyAsync(x_host, d_x, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RL_OK;
}

rl_status rl_gae(rl_traj *traj, rl_mlp *value_fn, float gamma, float lambda, float *adv_dev, float *rtg_dev) {
    if (!traj) return rl_fail(nullptr, RL_ERR_INVALID_ARG, "rl_gae: traj is NULL");
    rl_ctx *ctx = traj->ctx;
    const uint64_t T = traj->used_T ? traj->used_T : traj->T, E = traj->E;
    const float gl = lambda * gamma;  // f32 product, critics/mod.rs:198
    const unsigned block = 128, grid = rl_grid_for(E, block);
    if (value_fn) {
        RL_REQUIRE(ctx, value_fn->in_dim == (int)traj->F && value_fn->out_dim == 1,
                   "rl_gae: value function must map num_features -> 1");
        float *v;
        RL_TRY(rl_ctx_scratch(ctx, 2 * T * E * sizeof(float), (void **)&v));
        float *v_next = v + T * E;
        if (value_fn->n_hidden == 1 && value_fn->hidden == 128 && value_fn->in_dim <= 5 && value_fn->act == RL_ACT_RELU) {
            RL_LAUNCH(ctx, value_forward_pairs_kernel, rl_grid_for(T * E, 256), 256, 0, rl_mlp_view(value_fn), traj->obs,
                      traj->next_obs, traj->succ, T, E, v, v_next);
        } else {
            const size_t smem = value_fn->n_params * sizeof(float);
            RL_CUDA(ctx, cudaFuncSetAttribute(value_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            RL_LAUNCH(ctx, value_forward_kernel, rl_grid_for(T * E, 256), 256, smem, rl_mlp_view(value_fn), traj->obs,
                      traj->next_obs, traj->succ, T, E, v, v_next);
        }
        RL_LAUNCH(ctx, gae_scan_kernel<true>, grid, block, 0, traj->reward, v, v_next, traj->succ, T, E, gamma, gl,
                  adv_dev, rtg_dev);
    } else {
        RL_LAUNCH(ctx, gae_scan_kernel<false>, grid, block, 0, traj->reward, nullptr, nullptr, traj->succ, T, E, gamma,
                  gl, adv_dev, rtg_dev);
    }
    return RL_OK;
}

// Critic::advantages / reward_to_go with a recurrent state-value module (ValuesOpt<Chain<Gru, Linear>>,
// rl2-bandits.rs:412-419): V from SeqPacked over the stored episodes, V(next) of interrupted steps from one more
// step of the same sequence, then the same scan as rl_gae.
rl_status rl_gae_seq(rl_traj *traj, rl_grunet *value_fn, float gamma, float lambda, float *adv_dev, float *rtg_dev) {
    if (!traj || !value_fn) return rl_fail(traj ? traj->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_gae_seq: NULL argument");
    rl_ctx *ctx = traj->ctx;
    const rl_grunet_view net = rl_grunet_view_of(value_fn);
    RL_REQUIRE(ctx, net.in_dim == (int)traj->F && net.out_dim == 1, "rl_gae_seq: value function must map num_features -> 1");
    const uint64_t T = traj->used_T ? traj->used_T : traj->T, E = traj->E;
    const float gl = lambda * gamma;
    float *v;
    RL_TRY(rl_ctx_scratch(ctx, 2 * T * E * sizeof(float), (void **)&v));
    float *v_next = v + T * E;
    RL_TRY(rl_grunet_seq_enqueue(value_fn, traj, v, v_next));
    RL_LAUNCH(ctx, gae_scan_kernel<true>, rl_grid_for(E, 128), 128, 0, traj->reward, v, v_next, traj->succ, T, E, gamma, gl,
              adv_dev, rtg_dev);
    return RL_OK;
}

}  // extern "C"
