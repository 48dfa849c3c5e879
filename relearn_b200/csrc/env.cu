// env.cu -- environment handles and the unfused per-step kernels (one thread per lane, SoA state).
//
// K1  env_step_kernel<CartPoleEnv>: HBM-bound, 98 algorithmic bytes per env-step
//     (read 4 x f64 state + u32 packed counter/flag + u8 action = 37 B; write state 32 + packed 4 +
//     reward 4 + succ 1 + obs 5 x f32 = 61 B).  next_obs is written only on Interrupt.
#include "handles.cuh"

#include <cmath>

namespace {

template <class EnvT, bool REPLAY>
__global__ void __launch_bounds__(256) env_reset_kernel(typename EnvT::Params p, EnvStatePtrs st, NoiseSource nsrc,
                                                        uint64_t lane_offset, float *__restrict__ obs) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.E) return;
    LaneNoise<REPLAY> nz;
    nz.init(nsrc, lane_offset + e, e);
    typename EnvT::State s;
    EnvT::load(st, e, s);
    EnvT::template reset<REPLAY>(p, s, nz);
    float o[EnvT::MAXF];
    EnvT::observe(p, s, o);
    const int F = EnvT::num_features(p);
#pragma unroll
    for (int f = 0; f < EnvT::MAXF; ++f)
        if (f < F) obs[(uint64_t)f * st.E + e] = o[f];
    EnvT::store(st, e, s);
    if constexpr (std::is_same<EnvT, BanditMetaEnv>::value) BanditMetaEnv::store_means(st, e, (int)p.num_arms, s);
    nz.finish(nsrc, e);
}

template <class EnvT, bool REPLAY>
__global__ void __launch_bounds__(256)
    env_step_kernel(typename EnvT::Params p, EnvStatePtrs st, NoiseSource nsrc, uint64_t lane_offset,
                    const uint8_t *__restrict__ actions, float *__restrict__ obs, float *__restrict__ reward,
                    uint8_t *__restrict__ succ, float *__restrict__ next_obs) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.E) return;
    LaneNoise<REPLAY> nz;
    nz.init(nsrc, lane_offset + e, e);
    typename EnvT::State s;
    EnvT::load(st, e, s);
    if constexpr (std::is_same<EnvT, BanditMetaEnv>::value) {
        if (!((s.w >> 26) & 1u)) BanditMetaEnv::load_means(st, e, (int)p.num_arms, s);
    }
    const uint32_t action = actions[e];
    float r;
    int sc;
    if constexpr (std::is_same<EnvT, CartPoleEnv>::value) {
        // the branch-free step (polynomial sin/cos, inline Newton division) whenever its angle bound holds
        if (p.max_angle <= 0.5) {
            sc = CartPoleEnv::step_fast(p, s, action);
            r = 1.0f;
        } else {
            sc = EnvT::template step<REPLAY>(p, s, action, nz, r);
        }
    } else {
        sc = EnvT::template step<REPLAY>(p, s, action, nz, r);
    }
    const int F = EnvT::num_features(p);
    float o[EnvT::MAXF];
    if (sc == RL_INTERRUPT) {  // steps.rs:155-157: Interrupt carries observe(next_state)
        EnvT::observe(p, s, o);
#pragma unroll
        for (int f = 0; f < EnvT::MAXF; ++f)
            if (f < F) next_obs[(uint64_t)f * st.E + e] = o[f];
    }
    if (sc != RL_CONTINUE) {  // steps.rs:116-124: the next call starts a new episode
        nz.set_step(nsrc.step_counter + 1);
        EnvT::template reset<REPLAY>(p, s, nz);
        if constexpr (std::is_same<EnvT, BanditMetaEnv>::value) BanditMetaEnv::store_means(st, e, (int)p.num_arms, s);
    }
    EnvT::observe(p, s, o);
#pragma unroll
    for (int f = 0; f < EnvT::MAXF; ++f)
        if (f < F) obs[(uint64_t)f * st.E + e] = o[f];
    EnvT::store(st, e, s);
    reward[e] = r;
    succ[e] = (uint8_t)sc;
    nz.finish(nsrc, e);
}

template <class EnvT>
rl_status launch_reset(rl_env *env, const typename EnvT::Params &p) {
    rl_ctx *ctx = env->ctx;
    const unsigned block = 256, grid = rl_grid_for(env->E, block);
    if (env->noise.mode == RL_NOISE_REPLAY) {
        RL_LAUNCH(ctx, (env_reset_kernel<EnvT, true>), grid, block, 0, p, env->state, env->noise, env->lane_offset,
                  env->obs);
    } else {
        RL_LAUNCH(ctx, (env_reset_kernel<EnvT, false>), grid, block, 0, p, env->state, env->noise, env->lane_offset,
                  env->obs);
    }
    return RL_OK;
}

template <class EnvT>
rl_status launch_step(rl_env *env, const typename EnvT::Params &p, const uint8_t *actions) {
    rl_ctx *ctx = env->ctx;
    const unsigned block = 256, grid = rl_grid_for(env->E, block);
    if (env->noise.mode == RL_NOISE_REPLAY) {
        RL_LAUNCH(ctx, (env_step_kernel<EnvT, true>), grid, block, 0, p, env->state, env->noise, env->lane_offset,
                  actions, env->obs, env->reward, env->succ, env->next_obs);
    } else {
        RL_LAUNCH(ctx, (env_step_kernel<EnvT, false>), grid, block, 0, p, env->state, env->noise, env->lane_offset,
                  actions, env->obs, env->reward, env->succ, env->next_obs);
    }
    return RL_OK;
}

// UniformFloat<f64>::new_inclusive (rand 0.8.5 uniform.rs), host side
void uniform_inclusive(double low, double high, double *out_low, double *out_scale) {
    const double max_rand = 1.0 - 2.220446049250313e-16;
    double scale = (high - low) / max_rand;
    while (scale * max_rand + low > high) scale = std::nextafter(scale, -INFINITY);
    *out_low = low;
    *out_scale = scale;
}

}  // namespace

extern "C" {

void rl_cartpole_cfg_default(rl_cartpole_cfg *c, uint64_t max_steps_per_episode) {
    // PhysicalConstants / EnvironmentParams defaults, cartpole.rs:178-216
    c->gravity = 9.8; c->mass_cart = 1.0; c->mass_pole = 0.1; c->length_half_pole = 0.5;
    c->friction_cart = 0.01; c->friction_pole = 0.01; c->time_step = 0.02;
    c->action_force = 10.0; c->max_pos = 2.4;
    c->max_angle = 12.0 * (3.14159265358979323846264338327950288 / 180.0);
    c->discount_factor = 0.99;
    c->max_steps_per_episode = max_steps_per_episode;
    c->step_limit_visible = 1;
}

void rl_chain_cfg_default(rl_chain_cfg *c) {
    c->size = 5;  // chain.rs:38-45
    c->discount_factor = 0.95;
}

rl_status rl_env_create(rl_ctx *ctx, rl_env_kind kind, const void *cfg, uint64_t num_envs, uint64_t lane_offset,
                        uint64_t seed, rl_env **out) {
    RL_REQUIRE(ctx, ctx && cfg && out, "rl_env_create: NULL argument");
    RL_REQUIRE(ctx, num_envs > 0 && num_envs < (1ull << 31), "rl_env_create: num_envs out of range");
    RL_CUDA(ctx, cudaSetDevice(ctx->device));
    rl_env *env = new (std::nothrow) rl_env();
    if (!env) return rl_fail(ctx, RL_ERR_OOM, "rl_env_create: host allocation failed");
    env->ctx = ctx;
    env->kind = kind;
    env->E = num_envs;
    env->lane_offset = lane_offset;
    env->noise.mode = RL_NOISE_PHILOX;
    env->noise.seed = seed;
    env->state.E = num_envs;
    rl_env_structure &st = env->structure;
    size_t f64_planes = 0, mean_planes = 0;
    switch (kind) {
    case RL_ENV_CARTPOLE: {
        const rl_cartpole_cfg *c = (const rl_cartpole_cfg *)cfg;
        if (c->max_steps_per_episode >= (1ull << 31)) {
            delete env;
            return rl_fail(ctx, RL_ERR_UNSUPPORTED, "cartpole: max_steps_per_episode must be < 2^31");
        }
        CartPoleEnv::Params &p = env->cartpole;
        p.gravity = c->gravity; p.mass_cart = c->mass_cart; p.mass_pole = c->mass_pole;
        p.length_half_pole = c->length_half_pole; p.friction_cart = c->friction_cart;
        p.friction_pole = c->friction_pole; p.time_step = c->time_step; p.action_force = c->action_force;
        p.max_pos = c->max_pos; p.max_angle = c->max_angle;
        const double total_mass = c->mass_cart + c->mass_pole;  // cartpole.rs:238-251
        p.total_weight = c->gravity * total_mass;
        p.inv_total_mass = 1.0 / total_mass;
        p.mass_length_pole = c->mass_pole * c->length_half_pole;
        uniform_inclusive(-0.05, 0.05, &p.reset_low, &p.reset_scale);  // cartpole.rs:105
        p.max_steps = (uint32_t)c->max_steps_per_episode;
        p.visible = c->step_limit_visible != 0 ? 1u : 0u;
        st.num_features = CartPoleEnv::num_features(p);
        st.num_actions = 2;
        st.num_observations = 0;
        st.reward_lo = 0.0; st.reward_hi = 1.0;  // cartpole.rs:88-90
        st.discount_factor = c->discount_factor;
        f64_planes = 4;
        break;
    }
    case RL_ENV_CHAIN: {
        const rl_chain_cfg *c = (const rl_chain_cfg *)cfg;
        if (c->size < 1 || c->size > (uint64_t)ChainEnv::MAXF) {
            delete env;
            return rl_fail(ctx, RL_ERR_UNSUPPORTED, "chain: size must be in [1, %d]", ChainEnv::MAXF);
        }
        env->chain.size = (uint32_t)c->size;
        st.num_features = (int)c->size; st.num_actions = 2; st.num_observations = (int)c->size;
        st.reward_lo = 0.0; st.reward_hi = 10.0;  // chain.rs:60-62
        st.discount_factor = c->discount_factor;
        break;
    }
    case RL_ENV_MEMORY_GAME: {
        const rl_memory_cfg *c = (const rl_memory_cfg *)cfg;
        if (c->num_actions < 1 || c->num_actions + c->history_len > (uint64_t)MemoryEnv::MAXF) {
            delete env;
            return rl_fail(ctx, RL_ERR_UNSUPPORTED, "memory game: num_actions + history_len must be <= %d",
                           MemoryEnv::MAXF);
        }
        env->memory.num_actions = (uint32_t)c->num_actions;
        env->memory.history_len = (uint32_t)c->history_len;
        st.num_features = (int)(c->num_actions + c->history_len);
        st.num_actions = (int)c->num_actions;
        st.num_observations = st.num_features;
        st.reward_lo = -1.0; st.reward_hi = 1.0;  // memory.rs:69-71
        st.discount_factor = 1.0;
        break;
    }
    case RL_ENV_BANDIT_META: {
        const rl_bandit_meta_cfg *c = (const rl_bandit_meta_cfg *)cfg;
        if (c->num_arms < 1 || c->num_arms > (uint64_t)BanditMetaEnv::MAX_ARMS || c->episodes_per_trial < 1 ||
            c->episodes_per_trial > 65535 || c->distribution > 1) {
            delete env;
            return rl_fail(ctx, RL_ERR_UNSUPPORTED, "bandit meta: num_arms in [1,%d], episodes_per_trial in [1,65535]",
                           BanditMetaEnv::MAX_ARMS);
        }
        env->bandit.num_arms = (uint32_t)c->num_arms;
        env->bandit.episodes_per_trial = (uint32_t)c->episodes_per_trial;
        env->bandit.one_hot = (uint32_t)c->distribution;
        uniform_inclusive(0.0, 1.0, &env->bandit.mean_low, &env->bandit.mean_scale);  // bandits.rs:100
        st.num_features = (int)c->num_arms + 4;  // meta.rs:357-363
        st.num_actions = (int)c->num_arms;
        st.num_observations = 0;
        st.reward_lo = 0.0; st.reward_hi = 1.0;
        st.discount_factor = 1.0;  // bandits.rs:164-166
        mean_planes = c->num_arms;
        break;
    }
    case RL_ENV_PARTITION_GAME: {
        st.num_features = PartitionEnv::MAXF; st.num_actions = 2; st.num_observations = 0;
        st.reward_lo = -1.0; st.reward_hi = 1.0;  // partition.rs feedback_space
        st.discount_factor = 0.999;               // partition.rs discount_factor
        break;
    }
    default:
        delete env;
        return rl_fail(ctx, RL_ERR_INVALID_ARG, "rl_env_create: unknown env kind %d", (int)kind);
    }
    const uint64_t E = num_envs;
    const size_t F = (size_t)st.num_features;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMalloc(p, bytes);
    };
    alloc((void **)&env->state.f64, f64_planes * E * sizeof(double));
    alloc((void **)&env->state.u32, E * sizeof(uint32_t));
    alloc((void **)&env->state.means, mean_planes * E * sizeof(double));
    alloc((void **)&env->obs, F * E * sizeof(float));
    alloc((void **)&env->next_obs, F * E * sizeof(float));
    alloc((void **)&env->reward, E * sizeof(float));
    alloc((void **)&env->succ, E);
    alloc((void **)&env->noise.env_cursor, E * sizeof(uint32_t));
    alloc((void **)&env->noise.actor_cursor, E * sizeof(uint32_t));
    if (e != cudaSuccess) {
        rl_env_destroy(env);
        return rl_fail(ctx, e == cudaErrorMemoryAllocation ? RL_ERR_OOM : RL_ERR_CUDA, "rl_env_create: %s",
                       cudaGetErrorString(e));
    }
    cudaMemsetAsync(env->state.u32, 0, E * sizeof(uint32_t), ctx->stream);
    if (env->state.f64) cudaMemsetAsync(env->state.f64, 0, f64_planes * E * sizeof(double), ctx->stream);
    if (env->state.means) cudaMemsetAsync(env->state.means, 0, mean_planes * E * sizeof(double), ctx->stream);
    cudaMemsetAsync(env->noise.env_cursor, 0, E * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(env->noise.actor_cursor, 0, E * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(env->next_obs, 0, F * E * sizeof(float), ctx->stream);
    *out = env;
    return RL_OK;
}

rl_status rl_env_destroy(rl_env *env) {
    if (!env) return RL_OK;
    cudaSetDevice(env->ctx->device);
    cudaStreamSynchronize(env->ctx->stream);
    cudaFree(env->state.f64); cudaFree(env->state.u32); cudaFree(env->state.means);
    cudaFree(env->obs); cudaFree(env->next_obs); cudaFree(env->reward); cudaFree(env->succ);
    cudaFree(env->noise.env_cursor); cudaFree(env->noise.actor_cursor);
    delete env;
    return RL_OK;
}

rl_status rl_env_structure_of(rl_env *env, rl_env_structure *out) {
    if (!env || !out) return rl_fail(env ? env->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_env_structure_of: NULL argument");
    *out = env->structure;
    return RL_OK;
}

rl_status rl_env_set_noise_replay(rl_env *env, const uint32_t *env_words_dev, const uint32_t *actor_words_dev,
                                  uint64_t words_per_lane) {
    if (!env) return rl_fail(nullptr, RL_ERR_INVALID_ARG, "rl_env_set_noise_replay: env is NULL");
    rl_ctx *ctx = env->ctx;
    RL_REQUIRE(ctx, words_per_lane < (1ull << 32), "words_per_lane too large");
    env->noise.mode = RL_NOISE_REPLAY;
    env->noise.env_words = env_words_dev;
    env->noise.actor_words = actor_words_dev;
    env->noise.words_per_lane = words_per_lane;
    RL_CUDA(ctx, cudaMemsetAsync(env->noise.env_cursor, 0, env->E * sizeof(uint32_t), ctx->stream));
    RL_CUDA(ctx, cudaMemsetAsync(env->noise.actor_cursor, 0, env->E * sizeof(uint32_t), ctx->stream));
    return RL_OK;
}

rl_status rl_env_set_noise_philox(rl_env *env, uint64_t seed, uint32_t step_counter) {
    if (!env) return rl_fail(nullptr, RL_ERR_INVALID_ARG, "rl_env_set_noise_philox: env is NULL");
    env->noise.mode = RL_NOISE_PHILOX;
    env->noise.seed = seed;
    env->noise.step_counter = step_counter;
    return RL_OK;
}

rl_status rl_env_reset_all(rl_env *env) {
    if (!env) return rl_fail(nullptr, RL_ERR_INVALID_ARG, "rl_env_reset_all: env is NULL");
    switch (env->kind) {
    case RL_ENV_CARTPOLE: return launch_reset<CartPoleEnv>(env, env->cartpole);
    case RL_ENV_CHAIN: return launch_reset<ChainEnv>(env, env->chain);
    case RL_ENV_MEMORY_GAME: return launch_reset<MemoryEnv>(env, env->memory);
    case RL_ENV_BANDIT_META: return launch_reset<BanditMetaEnv>(env, env->bandit);
    case RL_ENV_PARTITION_GAME: return launch_reset<PartitionEnv>(env, env->partition);
    }
    return rl_fail(env->ctx, RL_ERR_INVALID_ARG, "bad env kind");
}

rl_status rl_env_step(rl_env *env, const uint8_t *actions_dev, rl_step_out *out) {
    if (!env || !actions_dev) return rl_fail(env ? env->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_env_step: NULL argument");
    rl_status s = RL_ERR_INVALID_ARG;
    switch (env->kind) {
    case RL_ENV_CARTPOLE: s = launch_step<CartPoleEnv>(env, env->cartpole, actions_dev); break;
    case RL_ENV_CHAIN: s = launch_step<ChainEnv>(env, env->chain, actions_dev); break;
    case RL_ENV_MEMORY_GAME: s = launch_step<MemoryEnv>(env, env->memory, actions_dev); break;
    case RL_ENV_BANDIT_META: s = launch_step<BanditMetaEnv>(env, env->bandit, actions_dev); break;
    case RL_ENV_PARTITION_GAME: s = launch_step<PartitionEnv>(env, env->partition, actions_dev); break;
    }
    if (s != RL_OK) return s;
    env->noise.step_counter += 1;
    if (out) {
        out->obs = env->obs; out->reward = env->reward; out->succ = env->succ; out->next_obs = env->next_obs;
    }
    return RL_OK;
}

rl_status rl_env_observation(rl_env *env, const float **obs_dev) {
    if (!env || !obs_dev) return rl_fail(env ? env->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_env_observation: NULL argument");
    *obs_dev = env->obs;
    return RL_OK;
}

rl_status rl_env_get_state(rl_env *env, double *f64_host, uint32_t *u32_host) {
    if (!env) return rl_fail(nullptr, RL_ERR_INVALID_ARG, "rl_env_get_state: env is NULL");
    rl_ctx *ctx = env->ctx;
    RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (f64_host && env->state.f64)
        RL_CUDA(ctx, cudaMemcpy(f64_host, env->state.f64, 4 * env->E * sizeof(double), cudaMemcpyDeviceToHost));
    if (f64_host && env->state.means)
        RL_CUDA(ctx, cudaMemcpy(f64_host, env->state.means, env->bandit.num_arms * env->E * sizeof(double),
                                cudaMemcpyDeviceToHost));
    if (u32_host) RL_CUDA(ctx, cudaMemcpy(u32_host, env->state.u32, env->E * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RL_OK;
}

rl_status rl_env_set_state(rl_env *env, const double *f64_host, const uint32_t *u32_host) {
    if (!env) return rl_fail(nullptr, RL_ERR_INVALID_ARG, "rl_env_set_state: env is NULL");
    rl_ctx *ctx = env->ctx;
    RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (f64_host && env->state.f64)
        RL_CUDA(ctx, cudaMemcpy(env->state.f64, f64_host, 4 * env->E * sizeof(double), cudaMemcpyHostToDevice));
    if (f64_host && env->state.means)
        RL_CUDA(ctx, cudaMemcpy(env->state.means, f64_host, env->bandit.num_arms * env->E * sizeof(double),
                                cudaMemcpyHostToDevice));
    if (u32_host) RL_CUDA(ctx, cudaMemcpy(env->state.u32, u32_host, env->E * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return RL_OK;
}

}  // extern "C"
