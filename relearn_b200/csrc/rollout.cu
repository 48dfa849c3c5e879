// rollout.cu -- trajectories and the fused rollout kernels.
//
// One launch = Agent::actor + Steps + TakeAlignedSteps + write_experience + OnlineStepsSummary for
// every lane (src/simulation/steps.rs:113-168, take_steps.rs:18-89, agents/buffers/vec.rs:113-141,
// buffers/mod.rs:237-261, simulation/summary.rs:198-216, train.rs:98-158).  Environment state lives
// in registers for the whole rollout, the policy MLP is evaluated inside the step loop, and the only
// HBM traffic is the trajectory write stream (26 B per env-step for CartPole: obs 20 + action 1 +
// reward 4 + succ 1).
//
//  K2a rollout_kernel<EnvT>            one thread per lane; weights in shared memory (broadcast LDS)
//  K2b rollout_cartpole_coop_kernel<L> L threads per lane split the 128 hidden units and keep their
//                                      slice of the weights in registers; logits are combined by
//                                      an xor-shuffle butterfly.  Used when there are too few
//                                      lanes to fill the GPU with one thread each (E = 4096).
#include "handles.cuh"
#include "tcgen05.cuh"

#include <cstdio>
#include <cstdlib>

namespace {

enum { ST_STEPS = 0, ST_R, ST_R2, ST_EPS, ST_ER, ST_ER2, ST_EL, ST_EL2, ST_STORED_STEPS, ST_STORED_EPS, ST_COUNT };

struct RolloutArgs {
    uint64_t E, lane_offset, Tcap;
    NoiseSource noise;
    uint32_t min_steps, slack;
    float *obs, *reward, *next_obs;
    uint8_t *action, *succ;
    uint32_t *lane_len;
    uint8_t *lane_flags;
    int actor_kind;
    MlpView net;
    const uint8_t *actions;
    double eps;
    int training;
    const double *qtable;
    uint64_t qtable_lane_stride;  // S * A with one table per lane, 0 with one table shared by every lane (tabular.rs:148-156)
    const double *ucb_mean;       // UCB1: tables (ucb.rs:96-103), per-lane strides as for qtable (visits: S or 0)
    const unsigned long long *ucb_count, *ucb_visits;
    double ucb_rate;
    int S, A, F;
    double *partials;  // f64 [gridDim.x][ST_COUNT]
    int sm_count;      // K2w: CTAs past the first per SM put their dynamics warp on another sub-partition
    int dyn_first, dyn_second;  // K2w: warp index of the dynamics warp in the first / later CTAs of an SM
    int aux_first, aux_second;  // K2v / K2y: warp index of the aux (noise) warp
    int head_first, head_second;  // K2y: warp index of the head warp
    uint64_t role_table;          // K2q: four bits per hardware warp (policy index 0..7, QK_ROLE_DYN / _AUX / _IDLE)
    int policy_warps;             // K2q: policy warps per CTA (4 envs each)
};

struct LaneStats {
    double v[ST_COUNT];
    double cur_reward;
    uint32_t cur_len;
    __device__ void init() {
#pragma unroll
        for (int i = 0; i < ST_COUNT; ++i) v[i] = 0.0;
        cur_reward = 0.0;
        cur_len = 0;
    }
    // OnlineStepsSummary::push (summary.rs:198-216), kept as raw f64 sums (count, sum, sum of squares)
    __device__ void push(float reward, int succ) {
        const double r = (double)reward;
        v[ST_STEPS] += 1.0; v[ST_R] += r; v[ST_R2] += r * r;
        cur_len += 1;
        cur_reward += r;
        if (succ != RL_CONTINUE) {
            const double L = (double)cur_len;
            v[ST_EPS] += 1.0; v[ST_ER] += cur_reward; v[ST_ER2] += cur_reward * cur_reward;
            v[ST_EL] += L; v[ST_EL2] += L * L;
            cur_reward = 0.0;
            cur_len = 0;
        }
    }
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// Deterministic block reduction of the per-lane statistics into partials[blockIdx.x][*].
__device__ __forceinline__ void block_reduce_stats(const LaneStats &st, bool contributes, double *partials) {
    __shared__ double red[32][ST_COUNT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) {
        const double s = warp_sum(contributes ? st.v[i] : 0.0);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < ST_COUNT) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * ST_COUNT + threadIdx.x] = s;
    }
}

// Categorical::new + sample (categorical.rs:29-33,52-54) as inverse CDF over exp(log_softmax(z)).
template <int MAXA>
__device__ __forceinline__ uint32_t categorical_sample(const float *z, int A, float u) {
    float m = z[0];
#pragma unroll
    for (int k = 1; k < MAXA; ++k)
        if (k < A) m = fmaxf(m, z[k]);
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < MAXA; ++k)
        if (k < A) sum += expf(z[k] - m);
    const float lse = m + logf(sum);
    float c = 0.0f;
    uint32_t a = (uint32_t)(A - 1);
    bool found = false;
#pragma unroll
    for (int k = 0; k < MAXA; ++k)
        if (k < A && !found) {
            c += expf(z[k] - lse);
            if (u < c) { a = (uint32_t)k; found = true; }
        }
    return a;
}

template <int MAXA>
__device__ __forceinline__ uint32_t argmax_first(const float *z, int A) {
    uint32_t best = 0;
    float bv = z[0];
#pragma unroll
    for (int k = 1; k < MAXA; ++k)
        if (k < A && z[k] > bv) { bv = z[k]; best = (uint32_t)k; }
    return best;
}

// Weights in shared memory.  PACK8: per hidden unit 8 floats {w1[j][0..4], b1[j], w2[0][j], w2[1][j]}
// (F <= 5, A <= 2: two broadcast LDS.128 per unit).  Otherwise the reference layout verbatim.
template <bool PACK8>
__device__ void stage_weights(const MlpView &m, float *sw) {
    if (m.params == nullptr) return;
    const int F = m.in_dim, H = m.hidden, A = m.out_dim;
    if (PACK8 && m.n_hidden == 1) {
        for (int i = threadIdx.x; i < H * 8; i += blockDim.x) {
            const int j = i >> 3, c = i & 7;
            float v = 0.0f;
            if (c < 5) v = c < F ? m.w1()[j * F + c] : 0.0f;
            else if (c == 5) v = m.b1()[j];
            else v = (c - 6) < A ? m.w2()[(c - 6) * H + j] : 0.0f;
            sw[i] = v;
        }
        if (threadIdx.x < 2) sw[H * 8 + threadIdx.x] = (int)threadIdx.x < A ? m.b2()[threadIdx.x] : 0.0f;
    } else {
        const uint64_t np = m.n_params;
        for (uint64_t i = threadIdx.x; i < np; i += blockDim.x) sw[i] = m.params[i];
    }
}

template <class EnvT, bool PACK8>
__device__ __forceinline__ void mlp_logits(const MlpView &m, const float *sw, const float *obs, float *z) {
    const int F = m.in_dim, H = m.hidden, A = m.out_dim;
    if (m.n_hidden > 1) {  // MlpConfig::hidden_sizes with two or three entries: flat layout, layer-generic forward
        rl_mlp_eval_deep(m, sw, obs, z);
        return;
    }
    if constexpr (PACK8) {
        float z0 = sw[H * 8], z1 = sw[H * 8 + 1];
        const float4 *w4 = reinterpret_cast<const float4 *>(sw);
#pragma unroll 4
        for (int j = 0; j < H; ++j) {
            const float4 a = w4[2 * j], b = w4[2 * j + 1];
            float acc = b.y;
            acc = fmaf(a.x, obs[0], acc);
            acc = fmaf(a.y, obs[1], acc);
            acc = fmaf(a.z, obs[2], acc);
            acc = fmaf(a.w, obs[3], acc);
            acc = fmaf(b.x, obs[4], acc);
            const float h = rl_activate(m.act, acc);
            z0 = fmaf(b.z, h, z0);
            z1 = fmaf(b.w, h, z1);
        }
        z[0] = z0;
        z[1] = z1;
    } else {
        const float *w1 = sw, *b1 = w1 + (size_t)H * F, *w2 = b1 + H, *b2 = w2 + (size_t)A * H;
#pragma unroll
        for (int k = 0; k < EnvT::MAXA; ++k) z[k] = k < A ? b2[k] : 0.0f;
        for (int j = 0; j < H; ++j) {
            float acc = b1[j];
#pragma unroll
            for (int f = 0; f < EnvT::MAXF; ++f)
                if (f < F) acc = fmaf(w1[j * F + f], obs[f], acc);
            const float h = rl_activate(m.act, acc);
#pragma unroll
            for (int k = 0; k < EnvT::MAXA; ++k)
                if (k < A) z[k] = fmaf(w2[k * H + j], h, z[k]);
        }
    }
}

// True when the actor evaluates its network on this step (uniform over the grid, so that the
// cooperative kernel can shuffle logits without divergence).
__device__ __forceinline__ bool actor_needs_logits(const RolloutArgs &a) {
    return a.actor_kind == RL_ACTOR_CATEGORICAL_POLICY || (a.actor_kind == RL_ACTOR_EPS_GREEDY_Q && a.eps < 1.0);
}

// Actor::act for every actor kind (steps.rs:126-128); z holds the network output when needed.
template <class EnvT, bool REPLAY>
__device__ __forceinline__ uint32_t actor_act(const RolloutArgs &a, const typename EnvT::Params &p,
                                              const typename EnvT::State &s, LaneNoise<REPLAY> &nz, uint64_t e,
                                              uint32_t t, const float *z) {
    switch (a.actor_kind) {
    case RL_ACTOR_REPLAY_ACTIONS: return a.actions[(uint64_t)t * a.E + e];
    case RL_ACTOR_RANDOM: return rl_gen_range<REPLAY, RL_STREAM_ACTOR>(nz, (uint32_t)a.A);
    case RL_ACTOR_CATEGORICAL_POLICY: {  // policies/actor.rs:42-55
        const float u = rl_u32_to_f32(nz.template next_u32<RL_STREAM_ACTOR>());
        return categorical_sample<EnvT::MAXA>(z, a.A, u);
    }
    case RL_ACTOR_EPS_GREEDY_Q: {  // dqn.rs:360-379
        if (rl_gen_bool<REPLAY, RL_STREAM_ACTOR>(nz, a.eps)) return rl_gen_range<REPLAY, RL_STREAM_ACTOR>(nz, (uint32_t)a.A);
        return argmax_first<EnvT::MAXA>(z, a.A);
    }
    case RL_ACTOR_TABULAR_EPS_GREEDY: {  // tabular.rs:222-232
        if (a.training && rl_u64_to_f64(nz.template next_u64<RL_STREAM_ACTOR>()) < a.eps)
            return rl_gen_range<REPLAY, RL_STREAM_ACTOR>(nz, (uint32_t)a.A);
        const double *row = a.qtable + (uint64_t)e * a.qtable_lane_stride + (uint64_t)EnvT::observe_index(p, s) * a.A;
        uint32_t best = 0;
        double bv = row[0];
        for (int k = 1; k < a.A; ++k)
            if (row[k] > bv) { bv = row[k]; best = (uint32_t)k; }
        return best;
    }
    case RL_ACTOR_UCB1: {  // ucb.rs:214-243; argmax_by keeps the LAST maximal element (utils/iter/cmp.rs:58-76)
        const uint32_t o = EnvT::observe_index(p, s);
        const unsigned long long *cnt = a.ucb_count + (uint64_t)e * a.qtable_lane_stride + (uint64_t)o * a.A;
        uint32_t best = 0;
        if (a.training) {
            const double *mean = a.ucb_mean + (uint64_t)e * a.qtable_lane_stride + (uint64_t)o * a.A;
            const uint64_t vstride = a.qtable_lane_stride ? (uint64_t)a.S : 0;
            const double lsv = __dmul_rn(2.0, log((double)a.ucb_visits[(uint64_t)e * vstride + o]));
            double bv = 0.0;
            for (int k = 0; k < a.A; ++k) {
                const double u = __dadd_rn(__dmul_rn(__dsqrt_rn(__ddiv_rn(lsv, (double)cnt[k])), a.ucb_rate), mean[k]);
                if (k == 0 || u >= bv) { bv = u; best = (uint32_t)k; }
            }
        } else {
            unsigned long long bc = 0;
            for (int k = 0; k < a.A; ++k)
                if (k == 0 || cnt[k] >= bc) { bc = cnt[k]; best = (uint32_t)k; }
        }
        return best;
    }
    }
    return 0;
}

__device__ __forceinline__ float pick5(const float *v, int i) {
    return i == 0 ? v[0] : i == 1 ? v[1] : i == 2 ? v[2] : i == 3 ? v[3] : v[4];
}

// ------------------------------------------------------------------------------------------------
// K2a: one thread per lane
// ------------------------------------------------------------------------------------------------
template <class EnvT, bool REPLAY>
__global__ void __launch_bounds__(128) rollout_kernel(typename EnvT::Params p, RolloutArgs a) {
    constexpr bool PACK8 = EnvT::MAXF <= 5 && EnvT::MAXA <= 2;
    extern __shared__ __align__(16) float sw[];
    stage_weights<PACK8>(a.net, sw);
    __syncthreads();

    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < a.E;
    const bool needs_logits = actor_needs_logits(a);
    LaneStats st;
    st.init();
    if (valid) {
        const int F = a.F;
        LaneNoise<REPLAY> nz;
        nz.init(a.noise, a.lane_offset + e, e);
        const uint32_t t0 = a.noise.step_counter;
        typename EnvT::State s;
        float obs[EnvT::MAXF], last_obs[EnvT::MAXF];
#pragma unroll
        for (int f = 0; f < EnvT::MAXF; ++f) obs[f] = last_obs[f] = 0.0f;
        uint32_t n = a.min_steps ? a.min_steps + a.slack : 0;  // take_steps.rs:20-31
        uint32_t i = 0;
        int succ_last = RL_TERMINATE, succ_prev = RL_TERMINATE;
        if (n > 0) {  // train.rs:135: every period starts fresh episodes
            nz.set_step(t0);
            EnvT::template reset<REPLAY>(p, s, nz);
            EnvT::observe(p, s, obs);
        }
        while (n > 0) {
            nz.set_step(t0 + i);
            float z[EnvT::MAXA];
            if (needs_logits) mlp_logits<EnvT, PACK8>(a.net, sw, obs, z);
            const uint32_t action = actor_act<EnvT, REPLAY>(a, p, s, nz, e, i, z);
#pragma unroll
            for (int f = 0; f < EnvT::MAXF; ++f)
                if (f < F) {
                    a.obs[((uint64_t)i * F + f) * a.E + e] = obs[f];
                    last_obs[f] = obs[f];
                }
            float r;
            const int sc = EnvT::template step<REPLAY>(p, s, action, nz, r);
            if (sc == RL_INTERRUPT) {
                EnvT::observe(p, s, obs);
#pragma unroll
                for (int f = 0; f < EnvT::MAXF; ++f)
                    if (f < F) a.next_obs[((uint64_t)i * F + f) * a.E + e] = obs[f];
            }
            if (sc != RL_CONTINUE) {
                nz.set_step(t0 + i + 1);
                EnvT::template reset<REPLAY>(p, s, nz);
            }
            EnvT::observe(p, s, obs);
            a.action[(uint64_t)i * a.E + e] = (uint8_t)action;
            a.reward[(uint64_t)i * a.E + e] = r;
            a.succ[(uint64_t)i * a.E + e] = (uint8_t)sc;
            st.push(r, sc);
            succ_prev = succ_last;
            succ_last = sc;
            i += 1;
            n -= 1;
            if (sc != RL_CONTINUE && n <= a.slack) n = 0;  // take_steps.rs:83-88
        }
        // VecBuffer::end_experience -> finalize_last_episode (buffers/mod.rs:237-261)
        uint32_t len = i;
        uint32_t flags = 0;
        double eps = st.v[ST_EPS];
        if (i > 0 && succ_last == RL_CONTINUE) {
            len = i - 1;
            flags = 1;
            a.succ[(uint64_t)len * a.E + e] = RL_PAD;
            if (len > 0 && succ_prev == RL_CONTINUE) {
                flags = 3;
                a.succ[(uint64_t)(len - 1) * a.E + e] = RL_INTERRUPT;
#pragma unroll
                for (int f = 0; f < EnvT::MAXF; ++f)
                    if (f < F) a.next_obs[((uint64_t)(len - 1) * F + f) * a.E + e] = last_obs[f];
                eps += 1.0;
            }
        }
        a.lane_len[e] = len;
        a.lane_flags[e] = (uint8_t)flags;
        st.v[ST_STORED_STEPS] = (double)len;
        st.v[ST_STORED_EPS] = eps;
        nz.finish(a.noise, e);
    }
    block_reduce_stats(st, valid, a.partials);
}

// ------------------------------------------------------------------------------------------------
// K2g: one WARP per lane, for the network actors on modules the specialised kernels do not serve (any MlpConfig: one to
// three hidden layers, any activation) while the lanes are few: K2a evaluates the module serially in one thread (4096 envs
// are 128 warps on 148 SMs: 5-64-64-2 tanh: 16 ms per 256-step period), here the 32 threads of an env's warp split every
// layer's units (weights transposed with an odd pitch in shared memory, activations of deeper layers in a per-warp
// buffer, outputs combined by a butterfly so that every thread holds the same logits) and all of them carry the env, the
// noise state and the actor redundantly -- same values in every thread, no divergence; thread 0 stores.
// ------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 128, WG_MAX_ENVS = 65536;
// GL threads of a warp per env (32 / GL envs per warp): the redundant part of a step is shared by fewer threads as GL shrinks,
// the module's serial part per thread grows.
__host__ __device__ inline size_t rollout_warp_smem_bytes(const DeepLayout &d, int nwarps, int GL) {
    return ((size_t)((d.P_pad + 3) & ~3) + (size_t)nwarps * (32 / GL) * 2 * ((d.maxH + 31) / 32 * 32)) * sizeof(float);
}

template <class EnvT, int GL>
__device__ __forceinline__ void mlp_logits_warp(const MlpView &m, const DeepLayout &d, const float *__restrict__ th, float *__restrict__ hb,
                                                int HB, const float *obs, float *z) {
    const int sub = threadIdx.x & (GL - 1), L = m.n_hidden, A = m.out_dim;
    const float *vin = nullptr;  // layer 0 reads the observation from registers
    for (int l = 0; l < L; ++l) {
        const int n_in = d.in[l], n_out = d.out[l], ld = d.ld[l];
        const float *Wt = th + d.off_w[l], *bb = th + d.off_b[l];
        float *hout = hb + (l & 1) * HB;
        for (int j = sub; j < n_out; j += GL) {
            float a0 = bb[j], a1 = 0.0f;
            if (l == 0) {
#pragma unroll
                for (int f = 0; f < EnvT::MAXF; ++f)
                    if (f < n_in) a0 = fmaf(Wt[f * ld + j], obs[f], a0);
            } else {
                int f = 0;
                for (; f + 3 < n_in; f += 4) {  // four inputs per broadcast LDS.128 (the buffers are 16-byte aligned)
                    const float4 v = *reinterpret_cast<const float4 *>(vin + f);
                    a0 = fmaf(Wt[f * ld + j], v.x, a0);
                    a1 = fmaf(Wt[(f + 1) * ld + j], v.y, a1);
                    a0 = fmaf(Wt[(f + 2) * ld + j], v.z, a0);
                    a1 = fmaf(Wt[(f + 3) * ld + j], v.w, a1);
                }
                for (; f + 1 < n_in; f += 2) {
                    a0 = fmaf(Wt[f * ld + j], vin[f], a0);
                    a1 = fmaf(Wt[(f + 1) * ld + j], vin[f + 1], a1);
                }
                if (f < n_in) a0 = fmaf(Wt[f * ld + j], vin[f], a0);
            }
            hout[j] = rl_activate(m.act, a0 + a1);
        }
        __syncwarp();
        vin = hout;
    }
    const int n_in = d.in[L], ld = d.ld[L];
    const float *Wt = th + d.off_w[L], *bb = th + d.off_b[L];
    float pz[EnvT::MAXA];
#pragma unroll
    for (int k = 0; k < EnvT::MAXA; ++k) pz[k] = 0.0f;
    for (int j = sub; j < n_in; j += GL) {
        const float h = vin[j];
#pragma unroll
        for (int k = 0; k < EnvT::MAXA; ++k)
            if (k < A) pz[k] = fmaf(Wt[j * ld + k], h, pz[k]);
    }
#pragma unroll
    for (int k = 0; k < EnvT::MAXA; ++k) {
        float v = pz[k];
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);  // stays inside the env's GL threads
        z[k] = k < A ? v + bb[k] : 0.0f;
    }
    __syncwarp();  // the buffers are reused by the next step
}

template <class EnvT, bool REPLAY, int GL>
__global__ void __launch_bounds__(WG_THREADS) rollout_warp_kernel(typename EnvT::Params p, RolloutArgs a) {
    extern __shared__ __align__(16) float sw[];
    const DeepLayout d = rl_mlp_layout(a.net.in_dim, a.net.n_hidden, a.net.hid, a.net.out_dim);
    const int HB = (d.maxH + 31) / 32 * 32;
    for (int i = threadIdx.x; i < d.P; i += blockDim.x) sw[deep_pidx(d, i)] = a.net.params[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane & (GL - 1), grp = threadIdx.x / GL;  // grp: env slot in the CTA
    float *hb = sw + ((d.P_pad + 3) & ~3) + (size_t)grp * 2 * HB;

    const uint64_t e = (uint64_t)blockIdx.x * (WG_THREADS / GL) + grp;
    const bool valid = e < a.E;
    const uint64_t e_safe = valid ? e : 0;
    LaneStats st;
    st.init();
    const int F = a.F;
    LaneNoise<REPLAY> nz;
    nz.init(a.noise, a.lane_offset + e_safe, e_safe);
    const uint32_t t0 = a.noise.step_counter;
    typename EnvT::State s;
    float obs[EnvT::MAXF], last_obs[EnvT::MAXF];
#pragma unroll
    for (int f = 0; f < EnvT::MAXF; ++f) obs[f] = last_obs[f] = 0.0f;
    uint32_t n = (valid && a.min_steps) ? a.min_steps + a.slack : 0;  // take_steps.rs:20-31
    uint32_t i = 0;
    int succ_last = RL_TERMINATE, succ_prev = RL_TERMINATE;
    if (n > 0) {  // train.rs:135: every period starts fresh episodes
        nz.set_step(t0);
        EnvT::template reset<REPLAY>(p, s, nz);
        EnvT::observe(p, s, obs);
    }
    // The envs of a warp may stop at different steps (slack): the warp leaves the loop together, the module is evaluated for
    // every env of the warp (it holds the warp's shuffles and barriers), everything else only for the envs still collecting.
    while (__any_sync(0xffffffffu, n > 0)) {
        float z[EnvT::MAXA];
        mlp_logits_warp<EnvT, GL>(a.net, d, sw, hb, HB, obs, z);
        if (n > 0) {
            nz.set_step(t0 + i);
            const uint32_t action = actor_act<EnvT, REPLAY>(a, p, s, nz, e, i, z);
            if (sub == 0) {
#pragma unroll
                for (int f = 0; f < EnvT::MAXF; ++f)
                    if (f < F) a.obs[((uint64_t)i * F + f) * a.E + e] = obs[f];
            }
#pragma unroll
            for (int f = 0; f < EnvT::MAXF; ++f) last_obs[f] = obs[f];
            float r;
            const int sc = EnvT::template step<REPLAY>(p, s, action, nz, r);
            if (sc == RL_INTERRUPT) {
                EnvT::observe(p, s, obs);
                if (sub == 0) {
#pragma unroll
                    for (int f = 0; f < EnvT::MAXF; ++f)
                        if (f < F) a.next_obs[((uint64_t)i * F + f) * a.E + e] = obs[f];
                }
            }
            if (sc != RL_CONTINUE) {
                nz.set_step(t0 + i + 1);
                EnvT::template reset<REPLAY>(p, s, nz);
            }
            EnvT::observe(p, s, obs);
            if (sub == 0) {
                a.action[(uint64_t)i * a.E + e] = (uint8_t)action;
                a.reward[(uint64_t)i * a.E + e] = r;
                a.succ[(uint64_t)i * a.E + e] = (uint8_t)sc;
            }
            st.push(r, sc);
            succ_prev = succ_last;
            succ_last = sc;
            i += 1;
            n -= 1;
            if (sc != RL_CONTINUE && n <= a.slack) n = 0;  // take_steps.rs:83-88
        }
    }
    if (valid) {
        // VecBuffer::end_experience -> finalize_last_episode (buffers/mod.rs:237-261)
        uint32_t len = i;
        uint32_t flags = 0;
        double eps = st.v[ST_EPS];
        if (i > 0 && succ_last == RL_CONTINUE) {
            len = i - 1;
            flags = 1;
            if (sub == 0) a.succ[(uint64_t)len * a.E + e] = RL_PAD;
            if (len > 0 && succ_prev == RL_CONTINUE) {
                flags = 3;
                if (sub == 0) {
                    a.succ[(uint64_t)(len - 1) * a.E + e] = RL_INTERRUPT;
#pragma unroll
                    for (int f = 0; f < EnvT::MAXF; ++f)
                        if (f < F) a.next_obs[((uint64_t)(len - 1) * F + f) * a.E + e] = last_obs[f];
                }
                eps += 1.0;
            }
        }
        if (sub == 0) {
            a.lane_len[e] = len;
            a.lane_flags[e] = (uint8_t)flags;
            nz.finish(a.noise, e);
        }
        st.v[ST_STORED_STEPS] = (double)len;
        st.v[ST_STORED_EPS] = eps;
    }
    block_reduce_stats(st, valid && sub == 0, a.partials);
}

// ------------------------------------------------------------------------------------------------
// K2b: CartPole, LANES threads per lane, H = 128 hidden units split across them, weights in registers.
// Every thread of a group carries the full (redundant) f64 physics so that no state is exchanged;
// only the two partial logits cross lanes (log2(LANES) xor-shuffles each).
// ------------------------------------------------------------------------------------------------
template <int LANES, bool REPLAY>
__global__ void __launch_bounds__(128) rollout_cartpole_coop_kernel(CartPoleEnv::Params p, RolloutArgs a) {
    using EnvT = CartPoleEnv;
    constexpr int H = 128, UNITS = H / LANES;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t e = gtid / LANES;
    const int sub = (int)(gtid % LANES);
    const bool valid = e < a.E;
    const int F = a.F;

    // this thread's slice of the network: units sub, sub + LANES, ...
    float w1[UNITS][5], b1[UNITS], w2a[UNITS], w2b[UNITS];
    float b2a = 0.0f, b2b = 0.0f;
    const bool has_net = a.net.params != nullptr;
    if (has_net) {
#pragma unroll
        for (int u = 0; u < UNITS; ++u) {
            const int j = sub + u * LANES;
#pragma unroll
            for (int f = 0; f < 5; ++f) w1[u][f] = f < F ? a.net.w1()[j * F + f] : 0.0f;
            b1[u] = a.net.b1()[j];
            w2a[u] = a.net.w2()[j];
            w2b[u] = a.net.w2()[H + j];
        }
        b2a = a.net.b2()[0];
        b2b = a.net.b2()[1];
    }

    const bool needs_logits = actor_needs_logits(a);
    LaneStats st;
    st.init();
    LaneNoise<REPLAY> nz;
    const uint64_t e_safe = valid ? e : 0;  // out-of-range threads shadow lane 0 without storing anything
    nz.init(a.noise, a.lane_offset + e_safe, e_safe);
    const uint32_t t0 = a.noise.step_counter;
    EnvT::State s;
    float obs[5] = {0, 0, 0, 0, 0}, last_obs[5] = {0, 0, 0, 0, 0};
    uint32_t n = (valid && a.min_steps) ? a.min_steps + a.slack : 0;
    uint32_t i = 0;
    int succ_last = RL_TERMINATE, succ_prev = RL_TERMINATE;
    s.x = s.xd = s.th = s.thd = 0.0;
    s.meta = 0x80000000u | p.max_steps;
    if (n > 0) {
        nz.set_step(t0);
        EnvT::reset<REPLAY>(p, s, nz);
        EnvT::observe(p, s, obs);
    }
    // The loop is warp-uniform: every thread of the warp takes part in the logit shuffles until the
    // last lane of the warp has finished (lanes only finish at different times when slack > 0).
    while (__any_sync(0xffffffffu, n > 0)) {
        float z[2] = {0.0f, 0.0f};
        if (needs_logits) {
            float z0 = 0.0f, z1 = 0.0f;
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
                float acc = b1[u];
#pragma unroll
                for (int f = 0; f < 5; ++f) acc = fmaf(w1[u][f], obs[f], acc);
                const float h = rl_activate(a.net.act, acc);
                z0 = fmaf(w2a[u], h, z0);
                z1 = fmaf(w2b[u], h, z1);
            }
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) {
                z0 += __shfl_xor_sync(0xffffffffu, z0, o);
                z1 += __shfl_xor_sync(0xffffffffu, z1, o);
            }
            z[0] = z0 + b2a;
            z[1] = z1 + b2b;
        }
        if (n > 0) {
            nz.set_step(t0 + i);
            const uint32_t action = actor_act<EnvT, REPLAY>(a, p, s, nz, e, i, z);
#pragma unroll
            for (int f = 0; f < 5; ++f) last_obs[f] = obs[f];
            float r;
            const int sc = EnvT::step<REPLAY>(p, s, action, nz, r);
            // trajectory writes are spread over the group: thread `sub` owns one column of the record
            if (sub < F) a.obs[((uint64_t)i * F + sub) * a.E + e] = pick5(last_obs, sub);
            if (sc == RL_INTERRUPT) {
                EnvT::observe(p, s, obs);
                if (sub < F) a.next_obs[((uint64_t)i * F + sub) * a.E + e] = pick5(obs, sub);
            }
            if (sc != RL_CONTINUE) {
                nz.set_step(t0 + i + 1);
                EnvT::reset<REPLAY>(p, s, nz);
            }
            EnvT::observe(p, s, obs);
            if (sub == 5) a.action[(uint64_t)i * a.E + e] = (uint8_t)action;
            if (sub == 6) a.reward[(uint64_t)i * a.E + e] = r;
            if (sub == 7) a.succ[(uint64_t)i * a.E + e] = (uint8_t)sc;
            st.push(r, sc);
            succ_prev = succ_last;
            succ_last = sc;
            i += 1;
            n -= 1;
            if (sc != RL_CONTINUE && n <= a.slack) n = 0;
        }
    }
    if (valid) {
        uint32_t len = i;
        uint32_t flags = 0;
        double eps = st.v[ST_EPS];
        if (i > 0 && succ_last == RL_CONTINUE) {
            len = i - 1;
            flags = 1;
            // same thread as the in-loop store of these addresses, so program order applies
            if (sub == 7) a.succ[(uint64_t)len * a.E + e] = RL_PAD;
            if (len > 0 && succ_prev == RL_CONTINUE) {
                flags = 3;
                if (sub == 7) a.succ[(uint64_t)(len - 1) * a.E + e] = RL_INTERRUPT;
                if (sub < F) a.next_obs[((uint64_t)(len - 1) * F + sub) * a.E + e] = pick5(last_obs, sub);
                eps += 1.0;
            }
        }
        if (sub == 0) {
            a.lane_len[e] = len;
            a.lane_flags[e] = (uint8_t)flags;
        }
        st.v[ST_STORED_STEPS] = (double)len;
        st.v[ST_STORED_EPS] = eps;
        if (sub == 0) nz.finish(a.noise, e);
    }
    block_reduce_stats(st, valid && sub == 0, a.partials);
}


// ------------------------------------------------------------------------------------------------
// K2c: CartPole + 5->128->2 ReLU network, LANES threads per env (LANES = 1 .. 32), 32 / LANES envs per warp.
//
// The step chain  policy -> sample -> f64 physics -> observe  is strictly sequential per env, so the
// kernel is shaped by two costs: the physics (~the same instruction count per WARP however many envs
// the warp carries) and the 128-unit hidden layer (per env).  Packing 32 / LANES envs into a warp
// amortises the physics; splitting the hidden layer over LANES threads keeps enough warps in flight
// when envs are scarce (E = 4096: LANES = 8 -> 1024 warps, one wave over 148 SMs).
//
// Hidden units are processed as PAIRS with packed FP32 (FFMA2): pair q = units (q, q + 64).  Weights
// live in shared memory as four float4 planes [4][64] so that the LANES threads of a group read
// LANES consecutive 16-byte words (conflict-free) and the 32 / LANES groups read the same words
// (broadcast): 4 LDS.128 + 7 FFMA2 + 2 FMNMX per pair.
//   plane 0: w1[q][0] w1[q+64][0] w1[q][1] w1[q+64][1]     plane 2: w1[.][4] pair, b1 pair
//   plane 1: w1[.][2] pair, w1[.][3] pair                  plane 3: (w2[1][.] - w2[0][.]) pair, unused
// ------------------------------------------------------------------------------------------------
constexpr int GK_H = 128, GK_PAIRS = 64;

// Two-action categorical sampling as ONE comparison on the step chain.  Categorical::new + sample (categorical.rs:29-33,
// 52-54) takes action 0 iff u < exp(log_softmax(z))[0] = 1 / (1 + exp(z1 - z0)), i.e. iff z1 - z0 < log((1 - u) / u).
// The threshold depends only on the uniform, which is known before the policy is evaluated, so it is computed off the
// dependent chain (an IEEE divide and a logf instead of an exp and a reciprocal ON it).  Identical in exact arithmetic;
// decisions can differ only for u within ~1e-7 of the boundary, the near-tie class the logit summation order already allows.
// u = 0 gives +inf (action 0 for every finite logit difference, like u < p0); a NaN difference compares false (action 1).
__device__ __forceinline__ float rl_logit_threshold(float u) { return logf(__fdiv_rn(1.0f - u, u)); }

__device__ __forceinline__ void stage_pair_weights(const MlpView &m, float4 *sw4, float *tail, const CartPoleEnv::Params &p,
                                                   int rem_entries) {
    const int F = m.in_dim;
    const float *w1 = m.w1(), *b1 = m.b1(), *w2 = m.w2();
    for (int i = threadIdx.x; i < 4 * GK_PAIRS; i += blockDim.x) {
        const int c = i / GK_PAIRS, q = i - c * GK_PAIRS, j0 = q, j1 = q + GK_PAIRS;
        auto W1 = [&](int j, int f) { return f < F ? w1[j * F + f] : 0.0f; };
        float4 v;
        if (c == 0) v = make_float4(W1(j0, 0), W1(j1, 0), W1(j0, 1), W1(j1, 1));
        else if (c == 1) v = make_float4(W1(j0, 2), W1(j1, 2), W1(j0, 3), W1(j1, 3));
        else if (c == 2) v = make_float4(W1(j0, 4), W1(j1, 4), b1[j0], b1[j1]);
        else v = make_float4(__fsub_rn(w2[GK_H + j0], w2[j0]), __fsub_rn(w2[GK_H + j1], w2[j1]), 0.0f, 0.0f);
        sw4[i] = v;
    }
    // a two-action actor only needs z_1 - z_0 = sum_j (w2_1j - w2_0j) relu(pre_j) + (b2_1 - b2_0)
    if (threadIdx.x == 0) tail[0] = __fsub_rn(m.b2()[1], m.b2()[0]);
    if (threadIdx.x == 1) tail[1] = 0.0f;
    // StepLimitObs::remaining = steps_remaining as f64 / max_steps as f64, then `as f32` (step_limit.rs:194-200,
    // interval.rs:114): exact table instead of an f64 division per step
    for (int i = threadIdx.x; i < rem_entries; i += blockDim.x)
        tail[2 + i] = (float)__ddiv_rn((double)i, (double)p.max_steps);
}

constexpr int GK_REM_TABLE_MAX = 2048;

template <int LANES, bool REPLAY, int AK>
__global__ void __launch_bounds__(128) rollout_cartpole_group_kernel(CartPoleEnv::Params p, RolloutArgs a) {
    using EnvT = CartPoleEnv;
    constexpr int PPL = GK_PAIRS / LANES;  // pairs per thread
    extern __shared__ __align__(16) unsigned char gk_smem[];
    float4 *sw4 = reinterpret_cast<float4 *>(gk_smem);
    float *tail = reinterpret_cast<float *>(sw4 + 4 * GK_PAIRS);
    const bool rem_table = p.max_steps != 0 && p.max_steps < GK_REM_TABLE_MAX;
    stage_pair_weights(a.net, sw4, tail, p, rem_table ? (int)p.max_steps + 1 : 0);
    __syncthreads();
    const float b2d = tail[0];
    const float *rem = tail + 2;

    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t e = gtid / LANES;
    const int sub = (int)(gtid % LANES);
    const bool valid = e < a.E;
    const int F = a.F;
    const bool needs_logits = AK == RL_ACTOR_CATEGORICAL_POLICY || a.eps < 1.0;
    constexpr bool REGW = PPL <= 8;  // LANES >= 8: at most 128 weight registers per thread
    float4 wA[REGW ? PPL : 1], wB[REGW ? PPL : 1], wC[REGW ? PPL : 1];
    float2 wD[REGW ? PPL : 1];
    if constexpr (REGW) {
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int q = sub + LANES * u;
            wA[u] = sw4[q]; wB[u] = sw4[GK_PAIRS + q]; wC[u] = sw4[2 * GK_PAIRS + q];
            wD[u] = make_float2(sw4[3 * GK_PAIRS + q].x, sw4[3 * GK_PAIRS + q].y);
        }
    }

    auto remaining_feature = [&](uint32_t r) {
        return p.max_steps == 0 ? 0.0f : rem_table ? rem[r] : (float)__ddiv_rn((double)r, (double)p.max_steps);
    };
    auto observe = [&](const EnvT::State &s, float *obs) {
        obs[0] = (float)s.x; obs[1] = (float)s.xd; obs[2] = (float)s.th; obs[3] = (float)s.thd;
        obs[4] = remaining_feature(s.meta & 0x7FFFFFFFu);
    };
    const float rem_full = remaining_feature(p.max_steps);

    LaneNoise<REPLAY> nz;
    const uint64_t e_safe = valid ? e : 0;  // out-of-range threads shadow lane 0 without storing anything
    nz.init(a.noise, a.lane_offset + e_safe, e_safe);
    const uint32_t t0 = a.noise.step_counter;
    EnvT::State s;
    float obs[5] = {0, 0, 0, 0, 0}, last_obs[5] = {0, 0, 0, 0, 0};
    uint32_t n = (valid && a.min_steps) ? a.min_steps + a.slack : 0;
    uint32_t i = 0;
    int succ_last = RL_TERMINATE, succ_prev = RL_TERMINATE;
    s.x = s.xd = s.th = s.thd = 0.0;
    s.meta = 0x80000000u | p.max_steps;
    if (n > 0) {
        nz.set_step(t0);
        EnvT::reset<REPLAY>(p, s, nz);
        observe(s, obs);
    }
    // Column k of the step record (obs 0..4, action, reward, succ) is stored by thread k % LANES of the group.
    bool owns[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) owns[k] = valid && sub == k % LANES && (k >= 5 || k < F);
    const uint64_t FE = (uint64_t)F * a.E;
    uint64_t io = e_safe, is = e_safe;  // running offsets: obs planes advance by F*E per step, the others by E
    // OnlineStepsSummary::push (summary.rs:198-216) as raw f64 sums, branch-free
    // (CartPole's reward is the constant 1.0, cartpole.rs:140: step sums follow from the step count and an
    // episode's return equals its length; both are exact in f64)
    double n_eps = 0.0, sum_el = 0.0, sum_el2 = 0.0;
    uint32_t cur_len = 0;
    uint32_t shared_word = 0;  // Philox: this thread's actor word for step (i rounded down to LANES) + sub

    // The step loop is warp-uniform (all threads take part in the logit shuffles until the warp's last env is
    // done; threads whose env finished earlier keep stepping a dead state with every side effect masked) and,
    // apart from the rare Interrupt and reset paths, one basic block -- so the compiler interleaves the Philox
    // draw, the stores and the statistics with the dependent policy -> sample -> physics chain.
    while (__any_sync(0xffffffffu, n > 0)) {
        const bool active = n > 0;
        if (!REPLAY) nz.set_step(t0 + i);
        // the `remaining` feature of the next observation is known before the step: one less if the episode
        // continues, the full limit after a reset.  Looked up here, off the dependent chain.
        const uint32_t r_now = s.meta & 0x7FFFFFFFu;
        const float rem_cont = remaining_feature(r_now > 0 ? r_now - 1 : 0);
        // Speculative dynamics (LANES >= 2): the threads of a group used to compute the same f64 step redundantly after
        // the action was known.  Instead the lower half of the group steps a copy of the state with action 0 and the
        // upper half with action 1 BEFORE the policy is evaluated -- the two instruction streams are independent, so
        // the scheduler interleaves them -- and the sampled action only selects which half's result is taken (one
        // round of shuffles).  Same operations on the same operands as before: results are bit-identical.
        constexpr bool SPEC = LANES >= 2;
        EnvT::State cand = s;
        int cand_sc = RL_CONTINUE;
        if constexpr (SPEC) cand_sc = EnvT::step_fast(p, cand, sub >= LANES / 2 ? 1u : 0u);
        // the actor's uniform and its logit-space threshold: independent of the policy, issued ahead of it
        float theta = 0.0f;
        if (AK == RL_ACTOR_CATEGORICAL_POLICY) {
            uint32_t w = 0;
            if constexpr (REPLAY) {
                if (active) w = nz.template next_u32<RL_STREAM_ACTOR>();
            } else if constexpr (LANES == 1) {
                w = nz.template next_u32<RL_STREAM_ACTOR>();
            } else {
                // The LANES threads of a group would all compute the same Philox block; instead thread `sub`
                // computes the word of step (i rounded down to a multiple of LANES) + sub once every LANES
                // steps and the group picks the current one by shuffle (same words, 1/LANES of the work).
                // Valid because a warp's active envs advance in lockstep (i is warp-uniform among them).
                const uint32_t phase = i & (LANES - 1);
                if (__any_sync(0xffffffffu, active && phase == 0))
                    shared_word = (uint32_t)rl_philox_slot_impl(nz.seed, nz.lane, t0 + i - phase + sub, RL_STREAM_ACTOR, 0);
                w = __shfl_sync(0xffffffffu, shared_word, (threadIdx.x & 31 & ~(LANES - 1)) + phase);
            }
            theta = rl_logit_threshold(rl_u32_to_f32(w));
        }
        float d = 0.0f;  // z_1 - z_0
        if (needs_logits) {
            const float2 o0 = make_float2(obs[0], obs[0]), o1 = make_float2(obs[1], obs[1]), o2 = make_float2(obs[2], obs[2]);
            const float2 o3 = make_float2(obs[3], obs[3]), o4 = make_float2(obs[4], obs[4]);
            float2 za = make_float2(0.0f, 0.0f);
            if constexpr (REGW) {
                // this thread's pairs live in registers for the whole rollout: no shared-memory latency on the chain
                float2 pre[PPL];
#pragma unroll
                for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wA[u].x, wA[u].y), o0, make_float2(wC[u].z, wC[u].w));
#pragma unroll
                for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wA[u].z, wA[u].w), o1, pre[u]);
#pragma unroll
                for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wB[u].x, wB[u].y), o2, pre[u]);
#pragma unroll
                for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wB[u].z, wB[u].w), o3, pre[u]);
#pragma unroll
                for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wC[u].x, wC[u].y), o4, pre[u]);
                float2 zc = make_float2(0.0f, 0.0f);  // second accumulator: shorter chains
#pragma unroll
                for (int u = 0; u < PPL; ++u) {
                    const float2 h = make_float2(fmaxf(pre[u].x, 0.0f), fmaxf(pre[u].y, 0.0f));
                    if (u & 1) zc = __ffma2_rn(wD[u], h, zc);
                    else za = __ffma2_rn(wD[u], h, za);
                }
                za = __fadd2_rn(za, zc);
            } else {
#pragma unroll(PPL < 8 ? PPL : 8)
                for (int u = 0; u < PPL; ++u) {
                    const int q = sub + LANES * u;
                    const float4 A = sw4[q], B = sw4[GK_PAIRS + q], Cw = sw4[2 * GK_PAIRS + q];
                    const float2 D = *reinterpret_cast<const float2 *>(&sw4[3 * GK_PAIRS + q]);
                    float2 pre = make_float2(Cw.z, Cw.w);
                    pre = __ffma2_rn(make_float2(A.x, A.y), o0, pre);
                    pre = __ffma2_rn(make_float2(A.z, A.w), o1, pre);
                    pre = __ffma2_rn(make_float2(B.x, B.y), o2, pre);
                    pre = __ffma2_rn(make_float2(B.z, B.w), o3, pre);
                    pre = __ffma2_rn(make_float2(Cw.x, Cw.y), o4, pre);
                    const float2 h = make_float2(fmaxf(pre.x, 0.0f), fmaxf(pre.y, 0.0f));
                    za = __ffma2_rn(D, h, za);
                }
            }
            d = za.x + za.y;
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            d += b2d;
        }
        uint32_t action = 0;
        if (AK == RL_ACTOR_CATEGORICAL_POLICY) {
            action = d < theta ? 0u : 1u;  // policies/actor.rs:42-55 (see rl_logit_threshold)
        } else if (active) {  // dqn.rs:360-379
            if (rl_gen_bool<REPLAY, RL_STREAM_ACTOR>(nz, a.eps)) action = rl_gen_range<REPLAY, RL_STREAM_ACTOR>(nz, 2u);
            else action = d > 0.0f ? 1u : 0u;
        }
        if (active) {
            if (owns[0]) a.obs[io] = obs[0];
            if (owns[1]) a.obs[io + a.E] = obs[1];
            if (owns[2]) a.obs[io + 2 * a.E] = obs[2];
            if (owns[3]) a.obs[io + 3 * a.E] = obs[3];
            if (owns[4]) a.obs[io + 4 * a.E] = obs[4];
            if (owns[5]) a.action[is] = (uint8_t)action;
        }
#pragma unroll
        for (int f = 0; f < 5; ++f) last_obs[f] = active ? obs[f] : last_obs[f];
        int sc;
        if constexpr (SPEC) {
            const int src = (threadIdx.x & 31 & ~(LANES - 1)) + (action ? LANES / 2 : 0);
            s.x = __shfl_sync(0xffffffffu, cand.x, src);
            s.xd = __shfl_sync(0xffffffffu, cand.xd, src);
            s.th = __shfl_sync(0xffffffffu, cand.th, src);
            s.thd = __shfl_sync(0xffffffffu, cand.thd, src);
            s.meta = __shfl_sync(0xffffffffu, cand.meta, src);
            sc = __shfl_sync(0xffffffffu, cand_sc, src);
        } else {
            sc = EnvT::step_fast(p, s, action);
        }
        const float r = 1.0f;  // cartpole.rs:140
        if (active) {
            if (owns[6]) a.reward[is] = r;
            if (owns[7]) a.succ[is] = (uint8_t)sc;
        }
        if (sc == RL_INTERRUPT && active) {  // rare: once per max_steps
            observe(s, obs);  // (remaining == 0 here)
            if (sub == 0) {
#pragma unroll
                for (int f = 0; f < 5; ++f)
                    if (f < F) a.next_obs[io + (uint64_t)f * a.E] = obs[f];
            }
        }
        cur_len += active ? 1u : 0u;
        if (sc != RL_CONTINUE && active) {  // steps.rs:116-124: the next call starts a new episode
            nz.set_step(t0 + i + 1);
            EnvT::reset<REPLAY>(p, s, nz);
            const double ld = (double)cur_len;
            n_eps += 1.0;
            sum_el += ld;
            sum_el2 = fma(ld, ld, sum_el2);
            cur_len = 0;
        }
        obs[0] = (float)s.x; obs[1] = (float)s.xd; obs[2] = (float)s.th; obs[3] = (float)s.thd;
        obs[4] = sc != RL_CONTINUE ? rem_full : rem_cont;
        if (active) {
            succ_prev = succ_last;
            succ_last = sc;
            i += 1;
            n -= 1;
            if (sc != RL_CONTINUE && n <= a.slack) n = 0;  // take_steps.rs:83-88
            io += FE;
            is += a.E;
        }
    }
    LaneStats st;
    st.init();
    st.v[ST_STEPS] = st.v[ST_R] = st.v[ST_R2] = (double)i;
    st.v[ST_EPS] = n_eps; st.v[ST_ER] = st.v[ST_EL] = sum_el; st.v[ST_ER2] = st.v[ST_EL2] = sum_el2;
    if (valid) {
        uint32_t len = i;
        uint32_t flags = 0;
        double eps = n_eps;
        if (i > 0 && succ_last == RL_CONTINUE) {
            len = i - 1;
            flags = 1;
            // same thread as the in-loop store of these addresses, so program order applies
            if (sub == (7 % LANES)) a.succ[(uint64_t)len * a.E + e] = RL_PAD;
            if (len > 0 && succ_prev == RL_CONTINUE) {
                flags = 3;
                if (sub == (7 % LANES)) a.succ[(uint64_t)(len - 1) * a.E + e] = RL_INTERRUPT;
                if (sub == 0) {
#pragma unroll
                    for (int f = 0; f < 5; ++f)
                        if (f < F) a.next_obs[((uint64_t)(len - 1) * F + f) * a.E + e] = last_obs[f];
                }
                eps += 1.0;
            }
        }
        if (sub == 0) {
            a.lane_len[e] = len;
            a.lane_flags[e] = (uint8_t)flags;
            nz.finish(a.noise, e);
        }
        st.v[ST_STORED_STEPS] = (double)len;
        st.v[ST_STORED_EPS] = eps;
    }
    block_reduce_stats(st, valid && sub == 0, a.partials);
}

// ------------------------------------------------------------------------------------------------
// K2w: K2c with the two halves of the step chain on different warps (few envs: E <= 64 per SM, two waves of CTAs).
//
// In K2c a warp carries 4 envs x 8 threads and issues BOTH the policy (obs -> hidden layer -> logit difference ->
// compare) and the f64 dynamics of its envs, the latter redundantly in all 8 threads of an env; with two such warps
// per scheduler the period is set by their combined instruction stream (~1940 clk per step at E = 4096).  Here a CTA
// owns 16 envs: four POLICY warps (4 envs x 8 threads each, exactly K2c's register-resident hidden layer) and one
// DYNAMICS warp whose lane (env, a) steps env `env` with action `a` speculatively -- both actions of 16 envs fill the
// warp, so nothing is computed redundantly.  The halves meet through two named barriers and a 16-row mailbox in
// shared memory:
//     dynamics: select cand[action], reset if the episode ended, write obs_{t+1} -> bar.arrive 1 ... bar.sync 2
//     policy:   bar.sync 1, read obs_t, hidden layer, d < theta ?, write action  -> bar.arrive 2
// so a step costs max(dynamics, policy) + two hand-offs instead of their sum, and every warp issues only its own half.
// The dynamics warp is the bound (its dependent f64 chain), so everything that can leave it has: the would-be reset
// states (Philox) are produced by the policy warps four steps at a time into a shared-memory ring, the policy warps
// store the observation / action columns, and the dynamics run on warp 2 rather than warp 4, which shares its
// sub-partition with warp 0 (measured placements in launch_ws).
// Same operations on the same operands as K2c<8> (summation order of the logits included): bit-identical trajectories
// and summaries (tests/test_gpu_envs.py).  Philox noise and the categorical actor only; K2c serves the rest.
// 168 registers on purpose: 16 K registers per sub-partition / 32 lanes / 3 warps = 170 -- one more register class and
// only one CTA fits an SM.
// ------------------------------------------------------------------------------------------------
constexpr int WK_ENVS = 16, WK_THREADS = 160, WK_ROW = 8;

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

constexpr int WK_RING = 8;  // steps of would-be reset states kept ahead of the dynamics warp
constexpr size_t WK_SMEM = 4 * GK_PAIRS * sizeof(float4) + (4 + GK_REM_TABLE_MAX) * sizeof(float) +
                           WK_ENVS * WK_ROW * sizeof(float) + WK_ENVS * sizeof(uint32_t) + WK_RING * WK_ENVS * 4 * sizeof(double) + 16;

__global__ void __launch_bounds__(WK_THREADS) rollout_cartpole_ws_kernel(CartPoleEnv::Params p, RolloutArgs a) {
    using EnvT = CartPoleEnv;
    constexpr int LANES = 8, PPL = GK_PAIRS / LANES;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char gk_smem[];
    float4 *sw4 = reinterpret_cast<float4 *>(gk_smem);
    float *tail = reinterpret_cast<float *>(sw4 + 4 * GK_PAIRS);
    // mailbox rows [16][8], 16-byte aligned: obs 0..4, flags (bit 0: this env takes the step, bit 1: some env of the CTA does)
    float *obs_s = tail + 4 + GK_REM_TABLE_MAX;
    uint32_t *act_s = reinterpret_cast<uint32_t *>(obs_s + WK_ENVS * WK_ROW);        // [16]
    // reset_s[slot][env] = (x, x', theta, theta') CartPole::initial_state would draw at noise step t0 + slot-step: filled by
    // the policy warps four steps at a time (they have idle issue slots; the dynamics warp does not), eight steps deep
    double *reset_s = reinterpret_cast<double *>(act_s + WK_ENVS);                   // [8][16][4], 16-byte aligned
    const bool rem_table = p.max_steps != 0 && p.max_steps < GK_REM_TABLE_MAX;
    stage_pair_weights(a.net, sw4, tail, p, rem_table ? (int)p.max_steps + 1 : 0);
    __syncthreads();
    const float *rem = tail + 2;
    auto remaining_feature = [&](uint32_t r) {
        return p.max_steps == 0 ? 0.0f : rem_table ? rem[r] : (float)__ddiv_rn((double)r, (double)p.max_steps);
    };
    // Warp roles.  Warps map to the four sub-partitions by index, so warps 0 and 4 share one; the dynamics warp is the
    // critical one and sits on warp 2 (launch_ws: measured placements; the first and the later CTAs of an SM can differ).
    const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dyn_warp = (int)blockIdx.x >= a.sm_count ? a.dyn_second : a.dyn_first;
    const bool is_dyn = hw_warp == dyn_warp;
    const int warp = hw_warp < dyn_warp ? hw_warp : hw_warp - 1;  // policy warp index 0..3 (unused by the dynamics warp)
    const uint64_t e_base = (uint64_t)blockIdx.x * WK_ENVS;
    const uint32_t t0 = a.noise.step_counter;
    const uint64_t seed = a.noise.seed;
    const int F = a.F;
    const uint64_t FE = (uint64_t)F * a.E;
    LaneStats st;
    st.init();
    bool contributes = false;

    // Would-be reset states (cartpole.rs:103-115: four uniform draws in field order = Philox blocks 0 (x, x') and 1 (theta,
    // theta') of the reset stream at noise step t0 + k).  A policy warp fills the slots of steps k0 .. k0 + 3 for its four
    // envs in one go: lane = (env of the warp, step offset, block).
    auto fill_resets = [&](uint32_t k0) {
        const int genv = lane >> 3, off = (lane >> 1) & 3, blk = lane & 1, el_w = 4 * warp + genv;
        const uint64_t eg = e_base + el_w, lg = a.lane_offset + (eg < a.E ? eg : 0);
        const uint32_t k = k0 + (uint32_t)off;
        uint32_t o[4];
        rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ENV_RESET * 64u + (uint32_t)blk, (uint32_t)seed,
                         (uint32_t)(seed >> 32), o);
        const double v0 = rl_u64_to_uniform((uint64_t)o[0] | ((uint64_t)o[1] << 32), p.reset_low, p.reset_scale);
        const double v1 = rl_u64_to_uniform((uint64_t)o[2] | ((uint64_t)o[3] << 32), p.reset_low, p.reset_scale);
        *reinterpret_cast<double2 *>(reset_s + (((size_t)(k & (WK_RING - 1)) * WK_ENVS + el_w) * 4 + 2 * blk)) = make_double2(v0, v1);
    };
    if (!is_dyn) {
        fill_resets(0);  // steps 0 .. 3 (step 0 is the initial state) and 4 .. 7
        fill_resets(4);
    }
    __syncthreads();

    if (is_dyn) {
        // ------------------------------ dynamics warp: lane = (env el, action act) ------------------------------
        const int el = lane & 15, act = lane >> 4;
        const uint64_t e = e_base + el;
        const bool valid = e < a.E;
        const uint64_t e_safe = valid ? e : 0, lane_global = a.lane_offset + e_safe;
        const float rem_full = remaining_feature(p.max_steps);
        // the reset state of noise step t0 + k, from the ring the policy warps keep eight steps ahead
        auto fresh_state = [&](uint32_t k, EnvT::State &f) {
            const double2 *slot = reinterpret_cast<const double2 *>(reset_s + ((size_t)(k & (WK_RING - 1)) * WK_ENVS + el) * 4);
            const double2 lo = slot[0], hi = slot[1];
            f.x = lo.x; f.xd = lo.y; f.th = hi.x; f.thd = hi.y;
            f.meta = 0x80000000u | p.max_steps;
        };
        EnvT::State s;
        s.x = s.xd = s.th = s.thd = 0.0;
        s.meta = 0x80000000u | p.max_steps;
        uint32_t n = (valid && a.min_steps) ? a.min_steps + a.slack : 0;
        uint32_t i = 0, cur_len = 0;
        int succ_last = RL_TERMINATE, succ_prev = RL_TERMINATE;
        double n_eps = 0.0, sum_el = 0.0, sum_el2 = 0.0;
        float cur_obs[5] = {0, 0, 0, 0, 0}, last_obs[5] = {0, 0, 0, 0, 0};
        {
            EnvT::State f;
            fresh_state(0, f);
            if (n > 0) {
                s = f;
                cur_obs[0] = (float)s.x; cur_obs[1] = (float)s.xd; cur_obs[2] = (float)s.th; cur_obs[3] = (float)s.thd;
                cur_obs[4] = rem_full;
            }
        }
        bool any = __any_sync(FULL, n > 0);
        if (act == 0) {
            obs_s[el * WK_ROW + 0] = cur_obs[0]; obs_s[el * WK_ROW + 1] = cur_obs[1]; obs_s[el * WK_ROW + 2] = cur_obs[2];
        } else {
            obs_s[el * WK_ROW + 3] = cur_obs[3]; obs_s[el * WK_ROW + 4] = cur_obs[4];
            reinterpret_cast<uint32_t *>(obs_s)[el * WK_ROW + 5] = (n > 0 ? 1u : 0u) | (any ? 2u : 0u);
        }
        __syncwarp();
        named_bar_arrive(1, WK_THREADS);
        uint32_t it = 0;  // loop counter (= step index of the envs still active)
        while (any) {
            const bool active = n > 0;
            // before the action is known: this lane's candidate step, its half of the would-be reset state, and the
            // `remaining` feature of the next observation
            EnvT::State cand = s;
            const int cand_sc = EnvT::step_fast(p, cand, (uint32_t)act);
            EnvT::State fresh;
            fresh_state(it + 1, fresh);
            const uint32_t r_now = s.meta & 0x7FFFFFFFu;
            const float rem_cont = remaining_feature(r_now > 0 ? r_now - 1 : 0);
            named_bar_sync(2, WK_THREADS);
            const uint32_t action = act_s[el];
            const int src = el + 16 * (int)action;
            EnvT::State post;
            post.x = __shfl_sync(FULL, cand.x, src);
            post.xd = __shfl_sync(FULL, cand.xd, src);
            post.th = __shfl_sync(FULL, cand.th, src);
            post.thd = __shfl_sync(FULL, cand.thd, src);
            post.meta = __shfl_sync(FULL, cand.meta, src);
            const int sc = __shfl_sync(FULL, cand_sc, src);
            const bool ended = sc != RL_CONTINUE;  // steps.rs:116-124: the next call starts a new episode
            s.x = ended ? fresh.x : post.x; s.xd = ended ? fresh.xd : post.xd;
            s.th = ended ? fresh.th : post.th; s.thd = ended ? fresh.thd : post.thd;
            s.meta = ended ? fresh.meta : post.meta;
            uint32_t n_next = n;
            if (active) {
                n_next = n - 1;
                if (ended && n_next <= a.slack) n_next = 0;  // take_steps.rs:83-88
            }
            float nobs[5];
            nobs[0] = (float)s.x; nobs[1] = (float)s.xd; nobs[2] = (float)s.th; nobs[3] = (float)s.thd;
            nobs[4] = ended ? rem_full : rem_cont;
            if (act == 0) {
                obs_s[el * WK_ROW + 0] = nobs[0]; obs_s[el * WK_ROW + 1] = nobs[1]; obs_s[el * WK_ROW + 2] = nobs[2];
            } else {
                obs_s[el * WK_ROW + 3] = nobs[3]; obs_s[el * WK_ROW + 4] = nobs[4];
            }
            const bool any_next = __any_sync(FULL, n_next > 0);
            if (act == 1) reinterpret_cast<uint32_t *>(obs_s)[el * WK_ROW + 5] = (n_next > 0 ? 1u : 0u) | (any_next ? 2u : 0u);
            __syncwarp();
            named_bar_arrive(1, WK_THREADS);
            // ---- off the chain: the rest of the step record and the statistics ----
            if (active) {
                const uint64_t is = (uint64_t)i * a.E + e_safe;
                if (act == 1) a.reward[is] = 1.0f;  // cartpole.rs:140
                if (act == 0) a.succ[is] = (uint8_t)sc;
                if (sc == RL_INTERRUPT && act == 0) {  // rare: once per max_steps; the post-step observation (remaining == 0)
                    const float io4 = remaining_feature(post.meta & 0x7FFFFFFFu);
                    const uint64_t io = (uint64_t)i * FE + e_safe;
                    a.next_obs[io] = (float)post.x;
                    a.next_obs[io + a.E] = (float)post.xd;
                    a.next_obs[io + 2 * a.E] = (float)post.th;
                    a.next_obs[io + 3 * a.E] = (float)post.thd;
                    if (F > 4) a.next_obs[io + 4 * a.E] = io4;
                }
                cur_len += 1;
                if (ended) {
                    const double ld = (double)cur_len;
                    n_eps += 1.0;
                    sum_el += ld;
                    sum_el2 = fma(ld, ld, sum_el2);
                    cur_len = 0;
                }
#pragma unroll
                for (int f = 0; f < 5; ++f) last_obs[f] = cur_obs[f];
                succ_prev = succ_last;
                succ_last = sc;
                i += 1;
            }
#pragma unroll
            for (int f = 0; f < 5; ++f) cur_obs[f] = nobs[f];
            n = n_next;
            any = any_next;
            it += 1;
        }
        st.v[ST_STEPS] = st.v[ST_R] = st.v[ST_R2] = (double)i;
        st.v[ST_EPS] = n_eps; st.v[ST_ER] = st.v[ST_EL] = sum_el; st.v[ST_ER2] = st.v[ST_EL2] = sum_el2;
        if (valid && act == 0) {
            // VecBuffer::end_experience -> finalize_last_episode (buffers/mod.rs:237-261); same thread as the in-loop
            // stores of succ, so program order applies
            uint32_t len = i, flags = 0;
            double eps = n_eps;
            if (i > 0 && succ_last == RL_CONTINUE) {
                len = i - 1;
                flags = 1;
                a.succ[(uint64_t)len * a.E + e] = RL_PAD;
                if (len > 0 && succ_prev == RL_CONTINUE) {
                    flags = 3;
                    a.succ[(uint64_t)(len - 1) * a.E + e] = RL_INTERRUPT;
#pragma unroll
                    for (int f = 0; f < 5; ++f)
                        if (f < F) a.next_obs[((uint64_t)(len - 1) * F + f) * a.E + e] = last_obs[f];
                    eps += 1.0;
                }
            }
            a.lane_len[e] = len;
            a.lane_flags[e] = (uint8_t)flags;
            st.v[ST_STORED_STEPS] = (double)len;
            st.v[ST_STORED_EPS] = eps;
            contributes = true;
        }
    } else {
        // ------------------------------ policy warps: 4 envs x 8 threads (K2c<8>) ------------------------------
        const int grp = lane >> 3, sub = lane & 7, el = 4 * warp + grp;
        const uint64_t e = e_base + el;
        const bool valid = e < a.E;
        const uint64_t e_safe = valid ? e : 0, lane_global = a.lane_offset + e_safe;
        const float b2d = tail[0];
        float4 wA[PPL], wB[PPL], wC[PPL];
        float2 wD[PPL];
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int q = sub + LANES * u;
            wA[u] = sw4[q]; wB[u] = sw4[GK_PAIRS + q]; wC[u] = sw4[2 * GK_PAIRS + q];
            wD[u] = make_float2(sw4[3 * GK_PAIRS + q].x, sw4[3 * GK_PAIRS + q].y);
        }
        // Thread `sub` stores column `sub` of the step record: observation feature sub (< F) or, sub == 5, the action.
        const bool stores_obs = valid && sub < 5 && sub < F, stores_action = valid && sub == 5;
        float *obs_ptr = a.obs + (uint64_t)(sub < 5 ? sub : 0) * a.E + e_safe;
        uint8_t *act_ptr = a.action + e_safe;
        float shared_theta = 0.0f;
        for (uint32_t i = 0;; ++i) {
            // The uniform of step i and its logit-space threshold (rl_logit_threshold), ahead of the observation.  As in
            // K2c thread `sub` draws the Philox word of step (i rounded down to 8) + sub once every 8 steps; here it also
            // turns it into the threshold, so a step costs one shuffle instead of a divide and a logarithm per thread.
            const uint32_t phase = i & (LANES - 1);
            if (phase == 0)
                shared_theta = rl_logit_threshold(rl_u32_to_f32((uint32_t)rl_philox_slot_impl(seed, lane_global, t0 + i + sub, RL_STREAM_ACTOR, 0)));
            const float theta = __shfl_sync(FULL, shared_theta, (lane & ~(LANES - 1)) + phase);
            // Reset states: the dynamics warp is at iteration i - 1 or i and reads slots up to step i + 1; steps i + 4 .. i + 7
            // reuse the slots of steps i - 4 .. i - 1, whose readers (iterations <= i - 2) are done.
            if ((i & 3u) == 0u && i > 0) fill_resets(i + 4);
            named_bar_sync(1, WK_THREADS);
            const float4 ov = *reinterpret_cast<const float4 *>(obs_s + el * WK_ROW);
            const float2 tailv = *reinterpret_cast<const float2 *>(obs_s + el * WK_ROW + 4);
            const float ob4 = tailv.x;
            const uint32_t flags = __float_as_uint(tailv.y);
            const float mine = obs_s[el * WK_ROW + (sub < 5 ? sub : 0)];  // the column this thread stores (branch-free)
            if ((flags & 2u) == 0u) break;
            const bool active = (flags & 1u) != 0u;
            const float2 o0 = make_float2(ov.x, ov.x), o1 = make_float2(ov.y, ov.y), o2 = make_float2(ov.z, ov.z);
            const float2 o3 = make_float2(ov.w, ov.w), o4 = make_float2(ob4, ob4);
            float2 pre[PPL];
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wA[u].x, wA[u].y), o0, make_float2(wC[u].z, wC[u].w));
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wA[u].z, wA[u].w), o1, pre[u]);
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wB[u].x, wB[u].y), o2, pre[u]);
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wB[u].z, wB[u].w), o3, pre[u]);
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wC[u].x, wC[u].y), o4, pre[u]);
            float2 za = make_float2(0.0f, 0.0f), zc = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int u = 0; u < PPL; ++u) {
                const float2 h = make_float2(fmaxf(pre[u].x, 0.0f), fmaxf(pre[u].y, 0.0f));
                if (u & 1) zc = __ffma2_rn(wD[u], h, zc);
                else za = __ffma2_rn(wD[u], h, za);
            }
            za = __fadd2_rn(za, zc);
            float d = za.x + za.y;
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
            d += b2d;
            const uint32_t action = d < theta ? 0u : 1u;  // policies/actor.rs:42-55 (rl_logit_threshold)
            if (sub == 0) act_s[el] = action;
            __syncwarp();
            named_bar_arrive(2, WK_THREADS);
            // ---- off the chain: the observation and the action of the step record ----
            if (active && stores_obs) *obs_ptr = mine;
            if (active && stores_action) *act_ptr = (uint8_t)action;
            obs_ptr += FE;
            act_ptr += a.E;
        }
    }
    block_reduce_stats(st, contributes, a.partials);
}

#include "rollout_ws3.cuh"
#include "rollout_ws4.cuh"
#include "rollout_ws5.cuh"
#include "rollout_ws6.cuh"

// ------------------------------------------------------------------------------------------------
// K2t: CartPole + 5->128->2 ReLU network with the hidden layer on the tensor cores (tcgen05 + TMEM), for env counts
// at which the actor GEMM is dense: a CTA owns 128 envs, one env per thread and per TMEM lane, and every step of the
// rollout is one  pre[128 envs x 128 units] = X[128 x 48] . W1e[48 x 128]  (3 x tcgen05.mma, K = 16) whose operands
// are the exact three-piece bf16 split of the f32 observations / weights (tcgen05.cuh, same K-slot pairing as the
// update passes in pass_tc.cuh: hi.hi, hi.mid, mid.hi, mid.mid, hi.lo, lo.hi => pre is f32-accurate).  W1e is built
// once and stays in shared memory for the whole rollout; per step a thread writes its env's 80-byte row of X, thread 0
// issues the MMAs, and while they run every thread stores the observation, draws its Philox word and looks up the
// step-limit feature.  The epilogue reads the env's 128 pre-activations from TMEM (tcgen05.ld, thread = env) and
// folds them into the only quantity the two-action actor needs, z_1 - z_0 = sum_j (w2_1j - w2_0j) relu(pre_j) + (b2_1
// - b2_0): 128 FMNMX + 64 FFMA2 against broadcast shared-memory weights instead of the 448 FFMA2 + 256 LDS.128 of
// the FP32-pipe kernel (K2c, LANES = 1).  Sampling, the f64 step, resets, trajectory stores and statistics are K2c's.
// ------------------------------------------------------------------------------------------------
constexpr int TK_A1 = 0, TK_B1 = 5 * tc::TC_CHUNK, TK_Z = 10 * tc::TC_CHUNK, TK_W2D = 11 * tc::TC_CHUNK;
constexpr int TK_REM = TK_W2D + GK_H * (int)sizeof(float);
constexpr int TK_BAR = TK_REM + GK_REM_TABLE_MAX * (int)sizeof(float);
constexpr int TK_SMEM = TK_BAR + 32;
constexpr int TK_CTAS_PER_SM = 4;  // 4 x 128 TMEM columns = all 512

#ifdef TK_PROFILE
#define TK_STAMP(k) do { const long long _c = clock64(); tk_ph[k] += _c - tk_last; tk_last = _c; } while (0)
#else
#define TK_STAMP(k) do { } while (0)
#endif

template <bool REPLAY, int AK, bool SPEC>
__global__ void __launch_bounds__(128, SPEC ? 2 : TK_CTAS_PER_SM) rollout_cartpole_tc_kernel(CartPoleEnv::Params p, RolloutArgs a) {
    using namespace tc;
    using EnvT = CartPoleEnv;
    constexpr int NF = 6;  // 5 features + the bias input
    extern __shared__ __align__(128) unsigned char tk_smem[];
    unsigned char *sA1 = tk_smem + TK_A1, *sB1 = tk_smem + TK_B1;
    float *w2d = reinterpret_cast<float *>(tk_smem + TK_W2D);
    float *rem = reinterpret_cast<float *>(tk_smem + TK_REM);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(tk_smem + TK_BAR + 16);
    const uint32_t bar1 = smem_u32(tk_smem + TK_BAR);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int F = a.F;
    const bool rem_table = p.max_steps != 0 && p.max_steps < GK_REM_TABLE_MAX;

    // ---- one-time setup: hidden unit `tid` -> row `tid` of the B operand; output-layer difference weights ----
    {
        const float *w1 = a.net.w1(), *b1 = a.net.b1(), *w2 = a.net.w2();
        uint32_t hi[NF], mid[NF], lo[NF], e[40];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const float w = f == 5 ? b1[tid] : f < F ? w1[tid * F + f] : 0.0f;
            split3(w, hi[f], mid[f], lo[f]);
        }
#pragma unroll
        for (int k = 0; k < 40; ++k) {
            const int g = k / NF, f = k % NF;  // piece pairing: x [hi hi mid mid hi lo] . w [hi mid hi mid lo hi]
            e[k] = k >= 6 * NF ? 0u : (g == 0 || g == 2 || g == 5) ? hi[f] : (g == 1 || g == 3) ? mid[f] : lo[f];
        }
        store_row<5>(sB1, tid, e);
        w2d[tid] = w2[GK_H + tid] - w2[tid];
        *reinterpret_cast<uint4 *>(tk_smem + TK_Z + tid * 16) = make_uint4(0u, 0u, 0u, 0u);  // K-slots 40..47 of both operands
        if (rem_table)
            for (int i = tid; i <= (int)p.max_steps; i += 128) rem[i] = (float)__ddiv_rn((double)i, (double)p.max_steps);
    }
    const float b2d = a.net.b2()[1] - a.net.b2()[0];
    if (warp == 0) {
        tmem_alloc(smem_u32(tptr), 128);
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tptr[0] + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
    constexpr uint32_t IDESC1 = make_idesc(128, 128, false, false);  // X (K-major) . W1e (K-major)
    const uint32_t aA1 = smem_u32(sA1), aB1 = smem_u32(sB1), tmem_base = tptr[0];
    uint64_t descA[3], descB[3];  // K = 16 per instruction = two 8-element chunks; the last pairs chunk 4 with the zero chunk
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        descA[k] = make_desc(aA1 + k * 2 * TC_CHUNK, k < 2 ? TC_CHUNK : TK_Z - TK_A1 - 4 * TC_CHUNK, 128);
        descB[k] = make_desc(aB1 + k * 2 * TC_CHUNK, k < 2 ? TC_CHUNK : TK_Z - TK_B1 - 4 * TC_CHUNK, 128);
    }

    auto remaining_feature = [&](uint32_t r) {
        return p.max_steps == 0 ? 0.0f : rem_table ? rem[r] : (float)__ddiv_rn((double)r, (double)p.max_steps);
    };
    const float rem_full = remaining_feature(p.max_steps);

    const uint64_t e = (uint64_t)blockIdx.x * 128 + tid;
    const bool valid = e < a.E;
    LaneNoise<REPLAY> nz;
    const uint64_t e_safe = valid ? e : 0;  // out-of-range threads shadow lane 0 without storing anything
    nz.init(a.noise, a.lane_offset + e_safe, e_safe);
    const uint32_t t0 = a.noise.step_counter;
    EnvT::State s;
    float obs[5] = {0, 0, 0, 0, 0}, last_obs[5] = {0, 0, 0, 0, 0};
    uint32_t n = (valid && a.min_steps) ? a.min_steps + a.slack : 0;
    uint32_t i = 0;
    int succ_last = RL_TERMINATE, succ_prev = RL_TERMINATE;
    s.x = s.xd = s.th = s.thd = 0.0;
    s.meta = 0x80000000u | p.max_steps;
    if (n > 0) {
        nz.set_step(t0);
        EnvT::reset<REPLAY>(p, s, nz);
        obs[0] = (float)s.x; obs[1] = (float)s.xd; obs[2] = (float)s.th; obs[3] = (float)s.thd;
        obs[4] = rem_full;
    }
    const uint64_t FE = (uint64_t)F * a.E;
    uint64_t io = e_safe, is = e_safe;
    double n_eps = 0.0, sum_el = 0.0, sum_el2 = 0.0;  // OnlineStepsSummary as raw sums (reward is the constant 1.0)
    uint32_t cur_len = 0;

    // CTA-uniform step loop: a thread whose env is done keeps stepping a dead state with every side effect masked
#ifdef TK_PROFILE
    long long tk_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tk_last = clock64();
#endif
    for (uint32_t it = 0;; ++it) {
        const bool active = n > 0;
        // ---- A operand: this env's row of X (pieces of the 5 features and of the bias input 1) ----
        {
            uint32_t hi[NF], mid[NF], lo[NF], ex[40];
#pragma unroll
            for (int f = 0; f < 5; ++f) split3(obs[f], hi[f], mid[f], lo[f]);
            hi[5] = 0x3F800000u;
            mid[5] = 0u;
            lo[5] = 0u;
#pragma unroll
            for (int k = 0; k < 40; ++k) {
                const int g = k / NF, f = k % NF;
                ex[k] = k >= 6 * NF ? 0u : (g == 0 || g == 1 || g == 4) ? hi[f] : (g == 2 || g == 3) ? mid[f] : lo[f];
            }
            store_row<5>(sA1, tid, ex);
        }
        fence_async_smem();
        fence_before();
        TK_STAMP(0);
        if (!__syncthreads_or(active)) break;  // also: every thread has read the previous step's pre-activations
        TK_STAMP(1);
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int k = 0; k < 3; ++k) umma_bf16(tmem_base, descA[k], descB[k], IDESC1, k > 0);
            umma_commit(bar1);
        }
        TK_STAMP(2);
        // ---- while the MMAs run: noise, the step-limit feature of the next observation, the observation stores ----
        if (!REPLAY) nz.set_step(t0 + i);
        const uint32_t r_now = s.meta & 0x7FFFFFFFu;
        const float rem_cont = remaining_feature(r_now > 0 ? r_now - 1 : 0);
        uint32_t w = 0;
        bool explore = false;
        uint32_t explore_action = 0;
        float theta = 0.0f;  // logit-space threshold of this step's uniform (rl_logit_threshold), computed under the MMA
        if (AK == RL_ACTOR_CATEGORICAL_POLICY) {
            if (!REPLAY || active) w = nz.template next_u32<RL_STREAM_ACTOR>();
            theta = rl_logit_threshold(rl_u32_to_f32(w));
        } else if (active) {  // dqn.rs:360-379
            explore = rl_gen_bool<REPLAY, RL_STREAM_ACTOR>(nz, a.eps);
            if (explore) explore_action = rl_gen_range<REPLAY, RL_STREAM_ACTOR>(nz, 2u);
        }
        if (active) {
            a.obs[io] = obs[0];
            a.obs[io + a.E] = obs[1];
            a.obs[io + 2 * a.E] = obs[2];
            a.obs[io + 3 * a.E] = obs[3];
            if (F > 4) a.obs[io + 4 * a.E] = obs[4];
        }
#pragma unroll
        for (int f = 0; f < 5; ++f) last_obs[f] = active ? obs[f] : last_obs[f];

        // Speculative dynamics (few CTAs: the step chain of one CTA is the bound, not throughput): both actions' f64
        // steps are evaluated while the MMAs run and the sampled action selects one -- same operations on the same
        // operands, bit-identical results.
        EnvT::State cand0 = s, cand1 = s;
        int sc0 = RL_CONTINUE, sc1 = RL_CONTINUE;
        if constexpr (SPEC) {
            sc0 = EnvT::step_fast(p, cand0, 0u);
            sc1 = EnvT::step_fast(p, cand1, 1u);
        }
        // ... and, with counter-based noise, so is the state a reset at the end of this step would produce
        EnvT::State fresh = s;
        if constexpr (SPEC && !REPLAY) {
            LaneNoise<REPLAY> nzr = nz;
            nzr.set_step(t0 + i + 1);
            EnvT::reset<REPLAY>(p, fresh, nzr);
        }

        TK_STAMP(3);
        // ---- epilogue: z_1 - z_0 from this env's 128 pre-activations ----
        if constexpr (SPEC) mbar_wait_after(bar1, it & 1u, cand0.x, cand1.x, __dadd_rn(cand0.thd, cand1.thd), __dadd_rn(fresh.thd, fresh.x), w);
        else mbar_wait(bar1, it & 1u);
        fence_after();
        TK_STAMP(4);
        float2 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.0f, 0.0f);
        constexpr int LDW = SPEC ? 64 : 32;  // columns per TMEM load: one wait per 64 when registers allow (2 CTAs per SM)
#pragma unroll
        for (int c = 0; c < GK_H / LDW; ++c) {
            uint32_t r[LDW];
            if constexpr (LDW == 64) tmem_ld64(tmem_d + c * LDW, r);
            else tmem_ld32(tmem_d + c * LDW, r);
#pragma unroll
            for (int q = 0; q < LDW / 4; ++q) {
                const float4 wq = *reinterpret_cast<const float4 *>(w2d + c * LDW + q * 4);
                const float2 h0 = make_float2(fmaxf(__uint_as_float(r[4 * q]), 0.0f), fmaxf(__uint_as_float(r[4 * q + 1]), 0.0f));
                const float2 h1 = make_float2(fmaxf(__uint_as_float(r[4 * q + 2]), 0.0f), fmaxf(__uint_as_float(r[4 * q + 3]), 0.0f));
                acc[(2 * q) & 3] = __ffma2_rn(h0, make_float2(wq.x, wq.y), acc[(2 * q) & 3]);
                acc[(2 * q + 1) & 3] = __ffma2_rn(h1, make_float2(wq.z, wq.w), acc[(2 * q + 1) & 3]);
            }
        }
        acc[0] = __fadd2_rn(acc[0], acc[2]);
        acc[1] = __fadd2_rn(acc[1], acc[3]);
        const float d = ((acc[0].x + acc[0].y) + (acc[1].x + acc[1].y)) + b2d;
        TK_STAMP(5);
        uint32_t action;
        if (AK == RL_ACTOR_CATEGORICAL_POLICY) {
            // policies/actor.rs:42-55; exp(log_softmax(z))[0] for two logits is the logistic of z_0 - z_1 (see K2c)
            action = d < theta ? 0u : 1u;
        } else {
            action = explore ? explore_action : (d > 0.0f ? 1u : 0u);
        }
        if (active) a.action[is] = (uint8_t)action;
        int sc;
        if constexpr (SPEC) {
            s = action ? cand1 : cand0;
            sc = action ? sc1 : sc0;
        } else {
            sc = EnvT::step_fast(p, s, action);
        }
        if (active) {
            a.reward[is] = 1.0f;  // cartpole.rs:140
            a.succ[is] = (uint8_t)sc;
        }
        if (sc == RL_INTERRUPT && active) {  // rare: once per max_steps (remaining == 0 here)
            a.next_obs[io] = (float)s.x;
            a.next_obs[io + a.E] = (float)s.xd;
            a.next_obs[io + 2 * a.E] = (float)s.th;
            a.next_obs[io + 3 * a.E] = (float)s.thd;
            if (F > 4) a.next_obs[io + 4 * a.E] = remaining_feature(0);
        }
        TK_STAMP(6);
        cur_len += active ? 1u : 0u;
        if (sc != RL_CONTINUE && active) {  // steps.rs:116-124: the next call starts a new episode
            if constexpr (SPEC && !REPLAY) {
                s = fresh;
            } else {
                nz.set_step(t0 + i + 1);
                EnvT::reset<REPLAY>(p, s, nz);
            }
            const double ld = (double)cur_len;
            n_eps += 1.0;
            sum_el += ld;
            sum_el2 = fma(ld, ld, sum_el2);
            cur_len = 0;
        }
        obs[0] = (float)s.x; obs[1] = (float)s.xd; obs[2] = (float)s.th; obs[3] = (float)s.thd;
        obs[4] = sc != RL_CONTINUE ? rem_full : rem_cont;
        if (active) {
            succ_prev = succ_last;
            succ_last = sc;
            i += 1;
            n -= 1;
            if (sc != RL_CONTINUE && n <= a.slack) n = 0;  // take_steps.rs:83-88
            io += FE;
            is += a.E;
        }
        TK_STAMP(7);
    }
#ifdef TK_PROFILE
    if (blockIdx.x == 0 && (tid == 0 || tid == 77))
        printf("K2t tid %d steps %u clk/step: xbuild %lld sync %lld mma_issue %lld overlap %lld mbar_wait %lld epilogue %lld sample+step %lld tail %lld\n",
               tid, i, tk_ph[0] / (i ? i : 1), tk_ph[1] / (i ? i : 1), tk_ph[2] / (i ? i : 1), tk_ph[3] / (i ? i : 1), tk_ph[4] / (i ? i : 1),
               tk_ph[5] / (i ? i : 1), tk_ph[6] / (i ? i : 1), tk_ph[7] / (i ? i : 1));
#endif
    if (warp == 0) {
        fence_after();
        tmem_dealloc(tmem_base, 128);
    }
    LaneStats st;
    st.init();
    st.v[ST_STEPS] = st.v[ST_R] = st.v[ST_R2] = (double)i;
    st.v[ST_EPS] = n_eps; st.v[ST_ER] = st.v[ST_EL] = sum_el; st.v[ST_ER2] = st.v[ST_EL2] = sum_el2;
    if (valid) {
        uint32_t len = i;
        uint32_t flags = 0;
        double eps = n_eps;
        if (i > 0 && succ_last == RL_CONTINUE) {  // buffers/mod.rs:237-261: the dangling last step is dropped
            len = i - 1;
            flags = 1;
            a.succ[(uint64_t)len * a.E + e] = RL_PAD;
            if (len > 0 && succ_prev == RL_CONTINUE) {
                flags = 3;
                a.succ[(uint64_t)(len - 1) * a.E + e] = RL_INTERRUPT;
#pragma unroll
                for (int f = 0; f < 5; ++f)
                    if (f < F) a.next_obs[((uint64_t)(len - 1) * F + f) * a.E + e] = last_obs[f];
                eps += 1.0;
            }
        }
        a.lane_len[e] = len;
        a.lane_flags[e] = (uint8_t)flags;
        nz.finish(a.noise, e);
        st.v[ST_STORED_STEPS] = (double)len;
        st.v[ST_STORED_EPS] = eps;
    }
    block_reduce_stats(st, valid, a.partials);
}

// Sum the per-block partials in a fixed order (deterministic) and publish the summary: one warp per statistic,
// lane l adds rows l, l + 32, ... (independent coalesced-by-row loads), then a shuffle tree combines the lanes.
__global__ void __launch_bounds__(32 * ST_COUNT) rollout_finalize_kernel(const double *__restrict__ partials, int nblocks,
                                                                        double *__restrict__ out,
                                                                        double *__restrict__ traj_counts,
                                                                        double *__restrict__ host_out = nullptr) {
    const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (i >= ST_COUNT) return;
    double s = 0.0;
#pragma unroll 4
    for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * ST_COUNT + i];
    s = warp_sum(s);
    if (lane == 0) {
        out[i] = s;
        // the caller's summary: written straight into page-locked host memory (no separate device-to-host copy)
        if (host_out) host_out[i] = s;
        if (i == ST_STORED_STEPS) traj_counts[0] = s;
        if (i == ST_STORED_EPS) traj_counts[1] = s;
    }
}

template <class EnvT>
size_t rollout_smem_bytes(const rl_mlp *net) {
    if (!net) return 16;
    constexpr bool PACK8 = EnvT::MAXF <= 5 && EnvT::MAXA <= 2;
    return (PACK8 && net->n_hidden == 1) ? ((size_t)net->hidden * 8 + 4) * sizeof(float) : net->n_params * sizeof(float);
}

template <class EnvT>
rl_status launch_rollout(rl_ctx *ctx, const typename EnvT::Params &p, RolloutArgs &a, const rl_mlp *net, bool replay,
                         int *nblocks_out) {
    // Network actors on few lanes: one warp per lane (K2g).  K2a's serial evaluation wins back once its E / 32 warps fill the
    // GPU (scripts/time_rollout_shapes.py); RL_ROLLOUT_WARP=0 | 1 forces the choice (measurements).
    const bool uses_net = net && (a.actor_kind == RL_ACTOR_CATEGORICAL_POLICY || (a.actor_kind == RL_ACTOR_EPS_GREEDY_Q && a.eps < 1.0));
    if (uses_net) {
        const DeepLayout d = rl_mlp_layout(net->in_dim, net->n_hidden, net->hid, net->out_dim);
        // eight threads per env (four envs per warp) for the smaller modules once that still gives every scheduler a few warps
        // (E = 4096: 5-128 tanh 1.31 -> 0.94 ms per period, 5-256 ReLU 1.51 -> 1.22; 5-64-64 tanh 2.07 -> 3.06: its serial part per
        // thread grows fourfold), a whole warp per env otherwise
        const int GL = (a.E >= 2368 && net->n_params <= 3000) ? 8 : 32;
        const size_t wsmem = rollout_warp_smem_bytes(d, WG_THREADS / 32, GL);
        static const char *force = getenv("RL_ROLLOUT_WARP");
        // Cost model fitted to scripts/time_rollout_shapes.py (profiles/r2_summary.md section 7), in clocks per step.  K2a is
        // latency-bound -- ~3000 for env + noise + actor + record, plus the module evaluated serially: 80 per unit in its packed
        // one-hidden-layer form (160 with tanh / sigmoid), 26 per parameter otherwise -- until its E / 32 warps fill the GPU
        // (~40 K envs).  K2g issues for its resident warps: per wave of 148 x 28 warps ~8000 + 1.4 per parameter + 600 per hidden
        // layer with a warp per env; with four envs per warp the fixed part is shared and the module's part grows fourfold.
        constexpr bool PACK8 = EnvT::MAXF <= 5 && EnvT::MAXA <= 2;
        const bool smooth = !(net->act == RL_ACT_RELU || net->act == RL_ACT_IDENTITY);
        const double ka = 3000.0 + ((PACK8 && net->n_hidden == 1) ? (smooth ? 160.0 : 80.0) * net->hidden : 26.0 * (double)net->n_params);
        const double kg = 8000.0 + (1.4 * (double)net->n_params + 600.0 * net->n_hidden) * (32 / GL);
        const uint64_t envs_per_wave = 148ull * 28 * (32 / GL);
        // (below a full wave K2g is latency-bound too: ~4.2 clk per instruction of one env's step)
        int units = 0;
        for (int l = 0; l < net->n_hidden; ++l) units += net->hid[l];
        const double lat_g = 4.2 * (700.0 + (3.0 * (double)net->n_params + (smooth ? 40.0 : 0.0) * units) / GL);
        const double thr_g = kg * (double)a.E / (double)envs_per_wave, fill_a = a.E > 40000 ? (double)a.E / 40000.0 : 1.0;
        const bool pick = force ? force[0] == '1' : (a.E <= (uint64_t)WG_MAX_ENVS && (lat_g > thr_g ? lat_g : thr_g) < ka * fill_a);
        if (pick && wsmem <= 200 * 1024) {
            const unsigned wgrid = rl_grid_for(a.E, WG_THREADS / GL);
            double *wpartials;
            RL_TRY(rl_ctx_scratch(ctx, ((size_t)wgrid + 1) * ST_COUNT * sizeof(double), (void **)&wpartials));
            a.partials = wpartials + ST_COUNT;
            *nblocks_out = (int)wgrid;
#define RL_WARP_LAUNCH(RP, G)                                                                                                          \
    do {                                                                                                                               \
        RL_CUDA(ctx, cudaFuncSetAttribute(rollout_warp_kernel<EnvT, RP, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem)); \
        RL_LAUNCH(ctx, (rollout_warp_kernel<EnvT, RP, G>), wgrid, WG_THREADS, wsmem, p, a);                                            \
    } while (0)
            if (replay) {
                if (GL == 8) RL_WARP_LAUNCH(true, 8);
                else RL_WARP_LAUNCH(true, 32);
            } else {
                if (GL == 8) RL_WARP_LAUNCH(false, 8);
                else RL_WARP_LAUNCH(false, 32);
            }
#undef RL_WARP_LAUNCH
            return RL_OK;
        }
    }
    const unsigned block = 128, grid = rl_grid_for(a.E, block);
    const size_t smem = rollout_smem_bytes<EnvT>(net);
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    a.partials = partials + ST_COUNT;
    *nblocks_out = (int)grid;
    if (replay) {
        RL_CUDA(ctx, cudaFuncSetAttribute(rollout_kernel<EnvT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RL_LAUNCH(ctx, (rollout_kernel<EnvT, true>), grid, block, smem, p, a);
    } else {
        RL_CUDA(ctx, cudaFuncSetAttribute(rollout_kernel<EnvT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RL_LAUNCH(ctx, (rollout_kernel<EnvT, false>), grid, block, smem, p, a);
    }
    return RL_OK;
}

template <int LANES>
rl_status launch_coop(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, bool replay, int *nblocks_out) {
    const unsigned block = 128, grid = rl_grid_for(a.E * LANES, block);
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    a.partials = partials + ST_COUNT;
    *nblocks_out = (int)grid;
    if (replay) {
        RL_LAUNCH(ctx, (rollout_cartpole_coop_kernel<LANES, true>), grid, block, 0, p, a);
    } else {
        RL_LAUNCH(ctx, (rollout_cartpole_coop_kernel<LANES, false>), grid, block, 0, p, a);
    }
    return RL_OK;
}


template <int LANES, int AK>
rl_status launch_group_ak(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, bool replay, int *nblocks_out) {
    // one warp per CTA while CTAs are scarce, so that the warps spread evenly over the SM sub-partitions
    const unsigned block = a.E * LANES < (uint64_t)ctx->sm_count * 4 * 128 ? 32 : 128;
    const unsigned grid = rl_grid_for(a.E * LANES, block);
    const size_t smem = 4 * GK_PAIRS * sizeof(float4) + (2 + GK_REM_TABLE_MAX) * sizeof(float);
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    a.partials = partials + ST_COUNT;
    *nblocks_out = (int)grid;
    if (replay) {
        RL_LAUNCH(ctx, (rollout_cartpole_group_kernel<LANES, true, AK>), grid, block, smem, p, a);
    } else {
        RL_LAUNCH(ctx, (rollout_cartpole_group_kernel<LANES, false, AK>), grid, block, smem, p, a);
    }
    return RL_OK;
}

template <bool REPLAY, int AK, bool SPEC>
rl_status launch_tc_variant(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, unsigned grid, size_t smem) {
    RL_CUDA(ctx, cudaFuncSetAttribute(rollout_cartpole_tc_kernel<REPLAY, AK, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RL_LAUNCH(ctx, (rollout_cartpole_tc_kernel<REPLAY, AK, SPEC>), grid, 128, smem, p, a);
    return RL_OK;
}

template <int AK>
rl_status launch_tc_ak(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, bool replay, int *nblocks_out) {
    const unsigned block = 128, grid = rl_grid_for(a.E, block);
    // the request is padded so that at most TK_CTAS_PER_SM CTAs (128 TMEM columns each) are resident per SM
    const size_t smem = TK_SMEM > 56 * 1024 ? TK_SMEM : 56 * 1024;
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    a.partials = partials + ST_COUNT;
    *nblocks_out = (int)grid;
    static const char *spec_env = getenv("RL_TC_SPEC");  // 0 / 1 overrides the choice (measurements)
    const bool spec = spec_env ? spec_env[0] == '1' : grid <= 2u * (unsigned)ctx->sm_count;
    if (replay) return spec ? launch_tc_variant<true, AK, true>(ctx, p, a, grid, smem) : launch_tc_variant<true, AK, false>(ctx, p, a, grid, smem);
    return spec ? launch_tc_variant<false, AK, true>(ctx, p, a, grid, smem) : launch_tc_variant<false, AK, false>(ctx, p, a, grid, smem);
}

rl_status launch_tc(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, bool replay, int *nblocks_out) {
    if (a.actor_kind == RL_ACTOR_CATEGORICAL_POLICY)
        return launch_tc_ak<RL_ACTOR_CATEGORICAL_POLICY>(ctx, p, a, replay, nblocks_out);
    return launch_tc_ak<RL_ACTOR_EPS_GREEDY_Q>(ctx, p, a, replay, nblocks_out);
}

rl_status launch_ws(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, int *nblocks_out) {
    const unsigned grid = (unsigned)((a.E + WK_ENVS - 1) / WK_ENVS);
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    a.partials = partials + ST_COUNT;
    *nblocks_out = (int)grid;
    a.sm_count = ctx->sm_count;
    // The dynamics warp is warp 2 of its CTA: warps 0 and 4 of a 5-warp CTA share a sub-partition, and of the 25
    // (first CTA, second CTA of an SM) placements measured at E = 4096 (profiles/r1_summary.md section 13) the ones with the
    // second CTA's dynamics on warp 1 or 2 are fastest (0.156 ms; 0.162 with warp 4, 0.167 with warp 3).
    // RL_WS_DYN="a,b" overrides (measurements).
    a.dyn_first = 2;
    a.dyn_second = 2;
    if (const char *ov = getenv("RL_WS_DYN")) {
        int x = 0, y = 0;
        if (sscanf(ov, "%d,%d", &x, &y) == 2 && x >= 0 && x <= 4 && y >= 0 && y <= 4) { a.dyn_first = x; a.dyn_second = y; }
    }
    RL_LAUNCH(ctx, rollout_cartpole_ws_kernel, grid, WK_THREADS, WK_SMEM, p, a);
    return RL_OK;
}

rl_status launch_ws3(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, int *nblocks_out) {
    const unsigned grid = (unsigned)((a.E + YK_ENVS - 1) / YK_ENVS);
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    a.partials = partials + ST_COUNT;
    *nblocks_out = (int)grid;
    a.sm_count = ctx->sm_count;
    // Warps 4..7 sit on sub-partitions 0..3 next to policy warps 0..3.  RL_WS3_ROLES="d1,h1,a1,d2,h2,a2" overrides the
    // (first CTA, later CTAs of an SM) placements of the dynamics / head / aux warps (measurements).
    a.dyn_first = 4; a.head_first = 5; a.aux_first = 6; a.dyn_second = 6; a.head_second = 7; a.aux_second = 4;
    if (const char *ov = getenv("RL_WS3_ROLES")) {
        int v[6];
        if (sscanf(ov, "%d,%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6) {
            bool ok = true;
            for (int k = 0; k < 6; ++k) ok = ok && v[k] >= 4 && v[k] <= 7;
            ok = ok && v[0] != v[1] && v[0] != v[2] && v[1] != v[2] && v[3] != v[4] && v[3] != v[5] && v[4] != v[5];
            if (ok) { a.dyn_first = v[0]; a.head_first = v[1]; a.aux_first = v[2]; a.dyn_second = v[3]; a.head_second = v[4]; a.aux_second = v[5]; }
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        RL_CUDA(ctx, cudaFuncSetAttribute(rollout_cartpole_ws3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(YkShared)));
        attr_set = true;
    }
    RL_LAUNCH(ctx, rollout_cartpole_ws3_kernel, grid, YK_THREADS, sizeof(YkShared), p, a);
    return RL_OK;
}

rl_status launch_ws4(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, int *nblocks_out) {
    const unsigned grid = (unsigned)((a.E + VK_ENVS - 1) / VK_ENVS);
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    a.partials = partials + ST_COUNT;
    *nblocks_out = (int)grid;
    a.sm_count = ctx->sm_count;
    // Warps map to the four sub-partitions by index (warps 0 / 4 and 1 / 5 share one).  RL_WS4_ROLES="d1,a1,d2,a2" overrides
    // the (first CTA, later CTAs of an SM) placements of the dynamics / aux warps (measurements).
    a.dyn_first = 2; a.aux_first = 5; a.dyn_second = 2; a.aux_second = 5;
    if (const char *ov = getenv("RL_WS4_ROLES")) {
        int d1, a1, d2, a2;
        if (sscanf(ov, "%d,%d,%d,%d", &d1, &a1, &d2, &a2) == 4 && d1 >= 0 && d1 <= 5 && a1 >= 0 && a1 <= 5 && d1 != a1 &&
            d2 >= 0 && d2 <= 5 && a2 >= 0 && a2 <= 5 && d2 != a2) {
            a.dyn_first = d1; a.aux_first = a1; a.dyn_second = d2; a.aux_second = a2;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        RL_CUDA(ctx, cudaFuncSetAttribute(rollout_cartpole_ws4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VkShared)));
        attr_set = true;
    }
    RL_LAUNCH(ctx, rollout_cartpole_ws4_kernel, grid, VK_THREADS, sizeof(VkShared), p, a);
    return RL_OK;
}

rl_status launch_ws5(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, int *nblocks_out) {
    const unsigned grid = (unsigned)((a.E + VK_ENVS - 1) / VK_ENVS);
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    a.partials = partials + ST_COUNT;
    *nblocks_out = (int)grid;
    a.sm_count = ctx->sm_count;
    // Warps map to the four sub-partitions by index (warps 0 / 4 and 1 / 5 share one).  RL_WS4_ROLES="d1,a1,d2,a2" overrides
    // the (first CTA, later CTAs of an SM) placements of the dynamics / aux warps (measurements).
    a.dyn_first = 2; a.aux_first = 5; a.dyn_second = 2; a.aux_second = 5;
    if (const char *ov = getenv("RL_WS4_ROLES")) {
        int d1, a1, d2, a2;
        if (sscanf(ov, "%d,%d,%d,%d", &d1, &a1, &d2, &a2) == 4 && d1 >= 0 && d1 <= 5 && a1 >= 0 && a1 <= 5 && d1 != a1 &&
            d2 >= 0 && d2 <= 5 && a2 >= 0 && a2 <= 5 && d2 != a2) {
            a.dyn_first = d1; a.aux_first = a1; a.dyn_second = d2; a.aux_second = a2;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        RL_CUDA(ctx, cudaFuncSetAttribute(rollout_cartpole_ws5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZkShared)));
        attr_set = true;
    }
    RL_LAUNCH(ctx, rollout_cartpole_ws5_kernel, grid, VK_THREADS, sizeof(ZkShared), p, a);
    return RL_OK;
}

rl_status launch_ws6(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, int *nblocks_out) {
    // Twelve warps; warp w runs on sub-partition w % 4.  The dynamics warp (3) has its scheduler to itself (7 and 11 idle),
    // the aux warp (10) shares one with two policy warps.  With seven policy warps (28 envs per CTA) warp 9 idles too.
    // 28 envs per CTA when that still fits one CTA per SM (E = 4096: 147 CTAs instead of 128: less work per SM).
    // RL_WS6_TABLE="r0,...,r11" (policy index 0..7, 8 = dynamics, 9 = aux, 15 = idle) overrides (measurements).
    int table[QK_THREADS / 32] = {0, 1, 2, QK_ROLE_DYN, 3, 4, 5, QK_ROLE_IDLE, 6, 7, QK_ROLE_AUX, QK_ROLE_IDLE};
    int npw = QK_POLICY_WARPS;
    if ((a.E + 27) / 28 <= (uint64_t)ctx->sm_count && (a.E + 31) / 32 < (a.E + 27) / 28) {
        npw = 7;
        table[9] = QK_ROLE_IDLE;
    }
    if (const char *ov = getenv("RL_WS6_TABLE")) {
        int v[12], seen = 0, count = 0, dyn = 0, aux = 0;
        if (sscanf(ov, "%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8], &v[9],
                   &v[10], &v[11]) == 12) {
            bool ok = true;
            for (int i = 0; i < 12; ++i) {
                if (v[i] >= 0 && v[i] < 8) { ok = ok && !((seen >> v[i]) & 1); seen |= 1 << v[i]; ++count; }
                else if (v[i] == QK_ROLE_DYN) ++dyn;
                else if (v[i] == QK_ROLE_AUX) ++aux;
                else ok = ok && v[i] == QK_ROLE_IDLE;
            }
            ok = ok && dyn == 1 && aux == 1 && count >= 1 && seen == (1 << count) - 1;
            if (ok) {
                for (int i = 0; i < 12; ++i) table[i] = v[i];
                npw = count;
            }
        }
    }
    a.policy_warps = npw;
    a.role_table = 0;
    for (int i = 0; i < QK_THREADS / 32; ++i) a.role_table |= (uint64_t)table[i] << (4 * i);
    const unsigned envs = 4u * (unsigned)npw, grid = (unsigned)((a.E + envs - 1) / envs);
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    a.partials = partials + ST_COUNT;
    *nblocks_out = (int)grid;
    a.sm_count = ctx->sm_count;
    static bool attr_set = false;
    if (!attr_set) {
        RL_CUDA(ctx, cudaFuncSetAttribute(rollout_cartpole_ws6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QkShared)));
        attr_set = true;
    }
    RL_LAUNCH(ctx, rollout_cartpole_ws6_kernel, grid, QK_THREADS, sizeof(QkShared), p, a);
    return RL_OK;
}

template <int LANES>
rl_status launch_group(rl_ctx *ctx, const CartPoleEnv::Params &p, RolloutArgs &a, bool replay, int *nblocks_out) {
    if (a.actor_kind == RL_ACTOR_CATEGORICAL_POLICY)
        return launch_group_ak<LANES, RL_ACTOR_CATEGORICAL_POLICY>(ctx, p, a, replay, nblocks_out);
    return launch_group_ak<LANES, RL_ACTOR_EPS_GREEDY_Q>(ctx, p, a, replay, nblocks_out);
}

}  // namespace

extern "C" {

rl_status rl_traj_create(rl_env *env, uint64_t step_capacity, rl_traj **out) {
    if (!env || !out) return rl_fail(env ? env->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_traj_create: NULL argument");
    rl_ctx *ctx = env->ctx;
    RL_REQUIRE(ctx, step_capacity > 0 && step_capacity < (1ull << 31), "rl_traj_create: step_capacity out of range");
    RL_CUDA(ctx, cudaSetDevice(ctx->device));
    rl_traj *t = new (std::nothrow) rl_traj();
    if (!t) return rl_fail(ctx, RL_ERR_OOM, "rl_traj_create: host allocation failed");
    t->ctx = ctx; t->env = env; t->E = env->E; t->T = step_capacity; t->F = (uint64_t)env->structure.num_features;
    const size_t TE = (size_t)t->T * t->E;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    };
    alloc((void **)&t->obs, TE * t->F * sizeof(float));
    alloc((void **)&t->next_obs, TE * t->F * sizeof(float));
    alloc((void **)&t->reward, TE * sizeof(float));
    alloc((void **)&t->action, TE);
    alloc((void **)&t->succ, TE);
    alloc((void **)&t->lane_len, t->E * sizeof(uint32_t));
    alloc((void **)&t->lane_flags, t->E);
    alloc((void **)&t->counts_dev, 8 * sizeof(double));
    if (e != cudaSuccess) {
        rl_traj_destroy(t);
        return rl_fail(ctx, e == cudaErrorMemoryAllocation ? RL_ERR_OOM : RL_ERR_CUDA, "rl_traj_create: %s",
                       cudaGetErrorString(e));
    }
    cudaMemsetAsync(t->succ, RL_PAD, TE, ctx->stream);
    cudaMemsetAsync(t->lane_len, 0, t->E * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(t->lane_flags, 0, t->E, ctx->stream);
    cudaMemsetAsync(t->counts_dev, 0, 8 * sizeof(double), ctx->stream);
    *out = t;
    return RL_OK;
}

rl_status rl_traj_destroy(rl_traj *t) {
    if (!t) return RL_OK;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->obs); cudaFree(t->next_obs); cudaFree(t->reward); cudaFree(t->action); cudaFree(t->succ);
    cudaFree(t->lane_len); cudaFree(t->lane_flags); cudaFree(t->counts_dev);
    delete t;
    return RL_OK;
}

rl_status rl_traj_view_of(rl_traj *t, rl_traj_view *out) {
    if (!t || !out) return rl_fail(t ? t->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_traj_view_of: NULL argument");
    rl_ctx *ctx = t->ctx;
    double counts[2];
    RL_CUDA(ctx, cudaMemcpyAsync(counts, t->counts_dev, sizeof counts, cudaMemcpyDeviceToHost, ctx->stream));
    RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    t->num_steps = (uint64_t)counts[0];
    t->num_episodes = (uint64_t)counts[1];
    out->num_lanes = t->E; out->step_capacity = t->T; out->num_features = t->F;
    out->obs = t->obs; out->action = t->action; out->reward = t->reward; out->succ = t->succ;
    out->next_obs = t->next_obs; out->lane_len = t->lane_len; out->num_steps = t->num_steps;
    return RL_OK;
}

}  // extern "C"

namespace {
// lane_len and counts for caller-loaded trajectories
__global__ void traj_index_kernel(const uint8_t *__restrict__ succ, uint64_t T, uint64_t E, uint32_t *lane_len,
                                  double *partials) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    LaneStats st;
    st.init();
    if (e < E) {
        uint32_t len = 0, eps = 0;
        for (uint64_t t = 0; t < T; ++t) {
            const uint8_t sc = succ[t * E + e];
            if (sc == RL_PAD) break;
            len += 1;
            if (sc != RL_CONTINUE) eps += 1;
        }
        lane_len[e] = len;
        st.v[ST_STORED_STEPS] = (double)len;
        st.v[ST_STORED_EPS] = (double)eps;
    }
    block_reduce_stats(st, e < E, partials);
}
}  // namespace

extern "C" {

rl_status rl_traj_load(rl_traj *t, uint64_t steps, const float *obs_dev, const uint8_t *action_dev,
                       const float *reward_dev, const uint8_t *succ_dev, const float *next_obs_dev) {
    if (!t || !obs_dev || !action_dev || !reward_dev || !succ_dev)
        return rl_fail(t ? t->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_traj_load: NULL argument");
    rl_ctx *ctx = t->ctx;
    RL_REQUIRE(ctx, steps <= t->T, "rl_traj_load: steps exceed capacity");
    const size_t n = (size_t)steps * t->E;
    const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
    RL_CUDA(ctx, cudaMemsetAsync(t->succ, RL_PAD, (size_t)t->T * t->E, ctx->stream));
    RL_CUDA(ctx, cudaMemcpyAsync(t->obs, obs_dev, n * t->F * sizeof(float), k, ctx->stream));
    RL_CUDA(ctx, cudaMemcpyAsync(t->action, action_dev, n, k, ctx->stream));
    RL_CUDA(ctx, cudaMemcpyAsync(t->reward, reward_dev, n * sizeof(float), k, ctx->stream));
    RL_CUDA(ctx, cudaMemcpyAsync(t->succ, succ_dev, n, k, ctx->stream));
    if (next_obs_dev) RL_CUDA(ctx, cudaMemcpyAsync(t->next_obs, next_obs_dev, n * t->F * sizeof(float), k, ctx->stream));
    const unsigned block = 128, grid = rl_grid_for(t->E, block);
    double *partials;
    RL_TRY(rl_ctx_scratch(ctx, ((size_t)grid + 1) * ST_COUNT * sizeof(double), (void **)&partials));
    RL_CUDA(ctx, cudaMemsetAsync(t->lane_flags, 0, t->E, ctx->stream));  // loaded histories are already finalised
    RL_LAUNCH(ctx, traj_index_kernel, grid, block, 0, t->succ, steps, t->E, t->lane_len, partials + ST_COUNT);
    RL_LAUNCH(ctx, rollout_finalize_kernel, 1, 32 * ST_COUNT, 0, partials + ST_COUNT, (int)grid, partials, t->counts_dev);
    t->used_T = steps;
    return RL_OK;
}

rl_status rl_rollout(rl_env *env, const rl_actor_cfg *actor, rl_bound bound, rl_traj *traj, rl_steps_summary *summary) {
    if (!env || !actor || !traj) return rl_fail(env ? env->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_rollout: NULL argument");
    rl_ctx *ctx = env->ctx;
    RL_REQUIRE(ctx, traj->env == env, "rl_rollout: trajectory belongs to another env");
    const uint64_t cap = bound.min_steps ? bound.min_steps + bound.slack_steps : 0;
    RL_REQUIRE(ctx, cap <= traj->T, "rl_rollout: min_steps + slack_steps exceeds the trajectory capacity");
    // The Philox step index is 32 bits wide: past 2^32 steps per lane the (seed, lane, step) counters would recur and
    // every reset state and action uniform would be replayed.  Fail instead of wrapping; rl_env_set_noise_philox
    // with a fresh seed starts a new stream.
    RL_REQUIRE(ctx, (uint64_t)env->noise.step_counter + cap + 1 <= 0xFFFFFFFFull,
               "rl_rollout: the 32-bit Philox step counter of this env would wrap; reseed with rl_env_set_noise_philox");
    const rl_env_structure &es = env->structure;
    rl_mlp *net = actor->net;
    const bool seq_policy = actor->kind == RL_ACTOR_CATEGORICAL_POLICY && actor->seq_net != nullptr;
    const bool needs_net = !seq_policy && (actor->kind == RL_ACTOR_CATEGORICAL_POLICY || actor->kind == RL_ACTOR_EPS_GREEDY_Q);
    if (needs_net) {
        RL_REQUIRE(ctx, net != nullptr, "rl_rollout: actor needs a network");
        RL_REQUIRE(ctx, net->in_dim == es.num_features && net->out_dim == es.num_actions,
                   "rl_rollout: network dimensions do not match the environment");
    } else {
        net = nullptr;
    }
    if (actor->kind == RL_ACTOR_REPLAY_ACTIONS) RL_REQUIRE(ctx, actor->actions_dev, "rl_rollout: actions_dev is NULL");
    if (actor->kind == RL_ACTOR_TABULAR_EPS_GREEDY) {
        RL_REQUIRE(ctx, actor->table, "rl_rollout: table is NULL");
        RL_REQUIRE(ctx, (actor->table->R == env->E || actor->table->R == 1) && actor->table->S == es.num_observations &&
                            actor->table->A == es.num_actions,
                   "rl_rollout: table shape does not match the environment (one replica per lane, or one shared table)");
    }
    if (actor->kind == RL_ACTOR_UCB1) {
        RL_REQUIRE(ctx, actor->ucb, "rl_rollout: ucb is NULL");
        RL_REQUIRE(ctx, (actor->ucb->R == env->E || actor->ucb->R == 1) && actor->ucb->S == es.num_observations &&
                            actor->ucb->A == es.num_actions,
                   "rl_rollout: UCB1 tables do not match the environment (one replica per lane, or one shared set)");
    }
    RolloutArgs a{};
    a.E = env->E; a.lane_offset = env->lane_offset; a.Tcap = traj->T;
    a.noise = env->noise;
    a.min_steps = (uint32_t)bound.min_steps; a.slack = (uint32_t)bound.slack_steps;
    a.obs = traj->obs; a.reward = traj->reward; a.next_obs = traj->next_obs; a.action = traj->action; a.succ = traj->succ;
    a.lane_len = traj->lane_len;
    a.lane_flags = traj->lane_flags;
    a.actor_kind = actor->kind;
    a.net = rl_mlp_view(net);
    a.actions = actor->actions_dev;
    a.eps = actor->exploration_rate;
    a.training = actor->training;
    a.qtable = actor->table ? actor->table->q : nullptr;
    a.qtable_lane_stride = (actor->table && actor->table->R == env->E) ? (uint64_t)actor->table->S * actor->table->A : 0;
    if (actor->kind == RL_ACTOR_UCB1) {
        a.ucb_mean = actor->ucb->mean; a.ucb_count = actor->ucb->count; a.ucb_visits = actor->ucb->visits;
        a.ucb_rate = actor->ucb->rate;
        a.qtable_lane_stride = actor->ucb->R == env->E ? (uint64_t)actor->ucb->S * actor->ucb->A : 0;
    }
    a.S = es.num_observations; a.A = es.num_actions; a.F = es.num_features;

    RL_CUDA(ctx, cudaMemsetAsync(traj->succ, RL_PAD, (size_t)traj->T * traj->E, ctx->stream));
    const bool replay = env->noise.mode == RL_NOISE_REPLAY;
    int nblocks = 0;
    double *totals = nullptr, *summary_host = nullptr;
    if (seq_policy) {
        RL_TRY(rl_rollout_seq(env, actor->seq_net, bound, traj, &totals));
    } else {
    switch (env->kind) {
    case RL_ENV_CARTPOLE: {
        int lanes = actor->lanes_per_env;
        // K2c serves the two network actors on the reference's default module (MlpConfig: one hidden layer of
        // 128, ReLU); anything else takes the generic thread-per-env kernel K2a.
        const bool group_ok = net && net->n_hidden == 1 && net->hidden == GK_H && net->act == RL_ACT_RELU && es.num_actions == 2 &&
                              (es.num_features == 5 || es.num_features == 4) && env->cartpole.max_angle <= 0.5;
        static const bool legacy = getenv("RL_ROLLOUT_LEGACY") != nullptr;
        if (!group_ok || legacy) {
            const bool coop_ok = net && net->n_hidden == 1 && net->hidden == 128 && es.num_actions == 2;
            if (lanes == 0) lanes = 1;
            if (lanes > 1 && !coop_ok)
                return rl_fail(ctx, RL_ERR_UNSUPPORTED, "rl_rollout: lanes_per_env > 1 needs a 128-unit MLP on CartPole");
            switch (lanes) {
            case 1: RL_TRY((launch_rollout<CartPoleEnv>(ctx, env->cartpole, a, net, replay, &nblocks))); break;
            case 8: RL_TRY((launch_coop<8>(ctx, env->cartpole, a, replay, &nblocks))); break;
            case 16: RL_TRY((launch_coop<16>(ctx, env->cartpole, a, replay, &nblocks))); break;
            case 32: RL_TRY((launch_coop<32>(ctx, env->cartpole, a, replay, &nblocks))); break;
            default: return rl_fail(ctx, RL_ERR_UNSUPPORTED, "rl_rollout: lanes_per_env must be 0, 1, 8, 16 or 32");
            }
            break;
        }
        if (lanes == 0) {
            // auto, from the B200 sweeps in profiles/r1_summary.md (sections 0 and 10): with few envs the step chain of a
            // single warp is the bound, so the hidden layer is split over 8 threads (weights in registers) -- as long as
            // those warps are resident at once (224 registers: 8 warps of 4 envs per SM, E <= 32 envs per SM = 4736).
            // Beyond one wave the tensor-core kernel wins at every size measured (6 K envs: 0.31 ms per 256-step period
            // vs 0.39 for LANES = 4; 1 M envs: 24 G env-steps/s vs 15.5 G for LANES = 1).
            lanes = env->E <= (uint64_t)ctx->sm_count * 32 ? 8 : RL_LANES_TENSOR_CORE;
            // ... and within that one wave the warp-specialised kernel (K2w: bit-identical to K2c<8>, policy and dynamics
            // on different warps) is 30 % faster (E = 4096: 0.178 vs 0.254 ms per 256-step period); it serves the
            // categorical actor on Philox noise.  RL_ROLLOUT_WS=0 keeps K2c (measurements).
            // It also stays ahead of K2t through a second wave of CTAs (E <= 64 per SM = 9472: 0.27 .. 0.30 ms against K2t's
            // 0.32); from the third wave on K2t wins.
            static const char *ws_env = getenv("RL_ROLLOUT_WS");
            if (env->E <= (uint64_t)ctx->sm_count * 64 && !replay && a.actor_kind == RL_ACTOR_CATEGORICAL_POLICY &&
                !(ws_env && ws_env[0] == '0'))
                lanes = RL_LANES_WARP_SPECIALIZED;
        }
        switch (lanes) {
        case 1: RL_TRY((launch_group<1>(ctx, env->cartpole, a, replay, &nblocks))); break;
        case 2: RL_TRY((launch_group<2>(ctx, env->cartpole, a, replay, &nblocks))); break;
        case 4: RL_TRY((launch_group<4>(ctx, env->cartpole, a, replay, &nblocks))); break;
        case 8: RL_TRY((launch_group<8>(ctx, env->cartpole, a, replay, &nblocks))); break;
        case 16: RL_TRY((launch_group<16>(ctx, env->cartpole, a, replay, &nblocks))); break;
        case 32: RL_TRY((launch_group<32>(ctx, env->cartpole, a, replay, &nblocks))); break;
        case RL_LANES_TENSOR_CORE: RL_TRY(launch_tc(ctx, env->cartpole, a, replay, &nblocks)); break;
        case RL_LANES_WARP_SPECIALIZED:
            if (replay || a.actor_kind != RL_ACTOR_CATEGORICAL_POLICY)
                return rl_fail(ctx, RL_ERR_UNSUPPORTED, "rl_rollout: RL_LANES_WARP_SPECIALIZED serves the categorical actor on Philox noise");
            {
                // K2z while its 16-env CTAs fit one per SM (E = 1024: 0.115 .. 0.118 ms per period against K2v's 0.123 and
                // K2w's 0.131), K2q through two waves of its 32-env CTAs (E = 4096: 0.127 ms against K2w's 0.156; E = 4800 ..
                // 9472: 0.239 .. 0.242 ms against K2w's 0.268 .. 0.301 and K2t's 0.322), K2w beyond (explicit requests only:
                // `lanes_per_env = 0` switches to K2t there).  RL_WS_VARIANT = 1 | 3 | 4 | 5 | 6 forces K2w / K2y / K2v / K2z /
                // K2q (measurements: profiles/r2_summary.md).
                static const char *ws_variant = getenv("RL_WS_VARIANT");
                const char v = ws_variant ? ws_variant[0]
                               : a.E <= (uint64_t)VK_ENVS * ctx->sm_count ? '5'
                               : a.E <= 2ull * QK_ENVS * ctx->sm_count ? '6' : '1';
                if (v == '6') RL_TRY(launch_ws6(ctx, env->cartpole, a, &nblocks));
                else if (v == '5') RL_TRY(launch_ws5(ctx, env->cartpole, a, &nblocks));
                else if (v == '3') RL_TRY(launch_ws3(ctx, env->cartpole, a, &nblocks));
                else if (v == '4') RL_TRY(launch_ws4(ctx, env->cartpole, a, &nblocks));
                else RL_TRY(launch_ws(ctx, env->cartpole, a, &nblocks));
            }
            break;
        default:
            return rl_fail(ctx, RL_ERR_UNSUPPORTED,
                           "rl_rollout: lanes_per_env must be 0, 1, 2, 4, 8, 16, 32, RL_LANES_TENSOR_CORE or RL_LANES_WARP_SPECIALIZED");
        }
        break;
    }
    case RL_ENV_CHAIN: RL_TRY((launch_rollout<ChainEnv>(ctx, env->chain, a, net, replay, &nblocks))); break;
    case RL_ENV_MEMORY_GAME: RL_TRY((launch_rollout<MemoryEnv>(ctx, env->memory, a, net, replay, &nblocks))); break;
    case RL_ENV_BANDIT_META: RL_TRY((launch_rollout<BanditMetaEnv>(ctx, env->bandit, a, net, replay, &nblocks))); break;
    case RL_ENV_PARTITION_GAME: RL_TRY((launch_rollout<PartitionEnv>(ctx, env->partition, a, net, replay, &nblocks))); break;
    }
    totals = a.partials - ST_COUNT;
    if (summary) RL_TRY(rl_ctx_pinned(ctx, ST_COUNT * sizeof(double), (void **)&summary_host));
    RL_LAUNCH(ctx, rollout_finalize_kernel, 1, 32 * ST_COUNT, 0, a.partials, nblocks, totals, traj->counts_dev, summary_host);
    }
    env->noise.step_counter += (uint32_t)cap + 1;  // fresh noise for the next period
    traj->used_T = cap;
    if (summary) {
        double *host = summary_host;
        if (!host) {  // sequence policies: their finalize kernel leaves the totals on the device
            RL_TRY(rl_ctx_pinned(ctx, ST_COUNT * sizeof(double), (void **)&host));
            RL_CUDA(ctx, cudaMemcpyAsync(host, totals, ST_COUNT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        auto mv = [&](int n_i, int s_i, int s2_i) {
            rl_mean_var m{};
            m.count = (uint64_t)host[n_i];
            if (m.count) {
                m.mean = host[s_i] / host[n_i];
                m.squared_residual_sum = host[s2_i] - host[s_i] * m.mean;
                if (m.squared_residual_sum < 0.0) m.squared_residual_sum = 0.0;
            }
            return m;
        };
        summary->step_reward = mv(ST_STEPS, ST_R, ST_R2);
        summary->episode_reward = mv(ST_EPS, ST_ER, ST_ER2);
        summary->episode_length = mv(ST_EPS, ST_EL, ST_EL2);
        summary->num_stored_steps = (uint64_t)host[ST_STORED_STEPS];
        summary->num_stored_episodes = (uint64_t)host[ST_STORED_EPS];
        traj->num_steps = summary->num_stored_steps;
        traj->num_episodes = summary->num_stored_episodes;
    }
    return RL_OK;
}

}  // extern "C"
