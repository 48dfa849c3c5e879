// tabq.cu -- tabular Q-learning (src/agents/tabular.rs:84-232).
//
// The reference update is a strictly sequential fold over the buffer whose bootstrap term reads the
// live table (tabular.rs:160-178), so a single shared table cannot be updated in parallel and stay
// bit-exact.  The data-parallel unit here is the replica: replica r owns table r and folds lane r of
// the trajectory in order -- atomic-free, f64 values and u64 counts bit-identical to the reference
// fold of that lane.  K7: latency-bound (one dependent read-modify-write per step per replica).
#include "handles.cuh"

namespace {

__global__ void __launch_bounds__(128)
    tabq_update_kernel(double *__restrict__ q, unsigned long long *__restrict__ counts, uint64_t R, int S, int A,
                       double discount, const float *__restrict__ obs, const float *__restrict__ next_obs,
                       const uint8_t *__restrict__ action, const float *__restrict__ reward,
                       const uint8_t *__restrict__ succ, uint64_t T, uint64_t E, int F) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= R) return;
    double *tq = q + tid * (uint64_t)S * A;
    unsigned long long *tc = counts + tid * (uint64_t)S * A;
    // one table per lane: replica tid folds lane tid; one shared table (R == 1): thread 0 folds every lane in order
    const uint64_t lane_begin = R == E ? tid : 0, lane_end = R == E ? tid + 1 : E;
    for (uint64_t r = lane_begin; r < lane_end; ++r) {
    // finite observation spaces are stored one-hot (index.rs:97-115); recover the index
    auto obs_index = [&](const float *planes, uint64_t t) {
        int idx = 0;
        for (int f = 0; f < F; ++f)
            if (planes[(t * F + f) * E + r] != 0.0f) idx = f;
        return idx;
    };
    int cur = T > 0 ? obs_index(obs, 0) : 0;
    for (uint64_t t = 0; t < T; ++t) {
        const uint8_t sc = succ[t * E + r];
        if (sc == RL_PAD) break;
        // fold_transient (simulation/mod.rs:287-313): Continue borrows the next step's observation
        int nxt = 0;
        bool has_next = false;
        if (sc == RL_CONTINUE) {
            if (t + 1 >= T || succ[(t + 1) * E + r] == RL_PAD) break;  // trailing Continue is skipped
            nxt = obs_index(obs, t + 1);
            has_next = true;
        } else if (sc == RL_INTERRUPT) {
            nxt = obs_index(next_obs, t);
            has_next = true;
        }
        // step_update (tabular.rs:159-179)
        double discounted_next = 0.0;
        if (has_next) {
            double m = tq[nxt * A];
            for (int k = 1; k < A; ++k) m = fmax(m, tq[nxt * A + k]);
            discounted_next = __dmul_rn(m, discount);
        }
        const int idx = cur * A + action[t * E + r];
        const unsigned long long c = tc[idx] + 1ull;
        tc[idx] = c;
        const double value = __dadd_rn((double)reward[t * E + r], discounted_next);
        const double weight = __ddiv_rn(1.0, (double)c);
        double qv = __dmul_rn(tq[idx], __dsub_rn(1.0, weight));
        qv = __dadd_rn(qv, __dmul_rn(weight, value));
        tq[idx] = qv;
        if (sc == RL_CONTINUE) cur = nxt;
        else if (t + 1 < T && succ[(t + 1) * E + r] != RL_PAD) cur = obs_index(obs, t + 1);
    }
    }
}

}  // namespace

extern "C" {

rl_status rl_tabq_create(rl_ctx *ctx, uint64_t num_replicas, int32_t num_observations, int32_t num_actions,
                         double discount_factor, rl_tabq **out) {
    RL_REQUIRE(ctx, ctx && out, "rl_tabq_create: NULL argument");
    RL_REQUIRE(ctx, num_replicas > 0 && num_observations > 0 && num_actions > 0, "rl_tabq_create: empty table");
    RL_CUDA(ctx, cudaSetDevice(ctx->device));
    rl_tabq *t = new (std::nothrow) rl_tabq();
    if (!t) return rl_fail(ctx, RL_ERR_OOM, "rl_tabq_create: host allocation failed");
    t->ctx = ctx; t->R = num_replicas; t->S = num_observations; t->A = num_actions; t->discount = discount_factor;
    const size_t n = (size_t)num_replicas * num_observations * num_actions;
    cudaError_t e = cudaMalloc((void **)&t->q, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&t->counts, n * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        rl_tabq_destroy(t);
        return rl_fail(ctx, RL_ERR_OOM, "rl_tabq_create: %s", cudaGetErrorString(e));
    }
    // build_agent ignores initial_action_count/value and starts from zeros (tabular.rs:61-78,97-111)
    cudaMemsetAsync(t->q, 0, n * sizeof(double), ctx->stream);
    cudaMemsetAsync(t->counts, 0, n * sizeof(unsigned long long), ctx->stream);
    *out = t;
    return RL_OK;
}

rl_status rl_tabq_destroy(rl_tabq *t) {
    if (!t) return RL_OK;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->q); cudaFree(t->counts);
    delete t;
    return RL_OK;
}

rl_status rl_tabq_update(rl_tabq *t, rl_traj *traj) {
    if (!t || !traj) return rl_fail(t ? t->ctx : nullptr, RL_ERR_INVALID_ARG, "rl_tabq_update: NULL argument");
    rl_ctx *ctx = t->ctx;
    RL_REQUIRE(ctx, traj->E == t->R || t->R == 1, "rl_tabq_update: one replica per lane, or one shared table (num_replicas = 1)");
    RL_REQUIRE(ctx, (int)traj->F == t->S, "rl_tabq_update: observation space size mismatch");
    const uint64_t T = traj->used_T ? traj->used_T : traj->T;
    const unsigned block = 128, grid = rl_grid_for(t->R, block);
    RL_LAUNCH(ctx, tabq_update_kernel, grid, block, 0, t->q, t->counts, t->R, t->S, t->A, t->discount, traj->obs,
              traj->next_obs, traj->action, traj->reward, traj->succ, T, traj->E, (int)traj->F);
    return RL_OK;
}

rl_status rl_tabq_get_table(rl_tabq *t, double *q_host, uint64_t *counts_host) {
    if (!t) return rl_fail(nullptr, RL_ERR_INVALID_ARG, "rl_tabq_get_table: NULL argument");
    rl_ctx *ctx = t->ctx;
    const size_t n = (size_t)t->R * t->S * t->A;
    RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (q_host) RL_CUDA(ctx, cudaMemcpy(q_host, t->q, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (counts_host) RL_CUDA(ctx, cudaMemcpy(counts_host, t->counts, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return RL_OK;
}

rl_status rl_tabq_set_table(rl_tabq *t, const double *q_host, const uint64_t *counts_host) {
    if (!t) return rl_fail(nullptr, RL_ERR_INVALID_ARG, "rl_tabq_set_table: NULL argument");
    rl_ctx *ctx = t->ctx;
    const size_t n = (size_t)t->R * t->S * t->A;
    RL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (q_host) RL_CUDA(ctx, cudaMemcpy(t->q, q_host, n * sizeof(double), cudaMemcpyHostToDevice));
    if (counts_host) RL_CUDA(ctx, cudaMemcpy(t->counts, counts_host, n * sizeof(uint64_t), cudaMemcpyHostToDevice));
    return RL_OK;
}

}  // extern "C"
