// stubs.cu -- entry points declared in relearn_b200.h whose kernels are not built yet.
// They fail loudly (RL_ERR_UNSUPPORTED); nothing falls back to the CPU.
#include "handles.cuh"

extern "C" {

void rl_trpo_cfg_default(rl_trpo_cfg *c) {
    // trpo.rs:29-38, conjugate_gradient.rs:55-64
    c->max_policy_step_kl = 0.01; c->cg_iterations = 10; c->max_backtracks = 15; c->backtrack_ratio = 0.8;
    c->hpv_reg_coeff = 1e-5; c->accept_violation = 0;
}
void rl_adam_cfg_default(rl_adam_cfg *c) {
    // coptimizer.rs:136-168; eps is libtorch's default
    c->learning_rate = 1e-3; c->beta1 = 0.9; c->beta2 = 0.999; c->weight_decay = 0.0; c->eps = 1e-8;
}
rl_status rl_trpo_update(rl_traj *traj, const float *, rl_mlp *, const rl_trpo_cfg *, rl_trpo_stats *) {
    return rl_fail(traj ? traj->ctx : nullptr, RL_ERR_UNSUPPORTED, "rl_trpo_update: not built yet");
}
rl_status rl_adam_create(rl_mlp *mlp, const rl_adam_cfg *, rl_adam **) {
    return rl_fail(mlp ? mlp->ctx : nullptr, RL_ERR_UNSUPPORTED, "rl_adam_create: not built yet");
}
rl_status rl_adam_destroy(rl_adam *) { return RL_OK; }
rl_status rl_value_update(rl_traj *traj, const float *, rl_mlp *, rl_adam *, int32_t, rl_opt_stats *) {
    return rl_fail(traj ? traj->ctx : nullptr, RL_ERR_UNSUPPORTED, "rl_value_update: not built yet");
}
rl_status rl_replay_create(rl_env *env, uint64_t, rl_replay **) {
    return rl_fail(env ? env->ctx : nullptr, RL_ERR_UNSUPPORTED, "rl_replay_create: not built yet");
}
rl_status rl_replay_destroy(rl_replay *) { return RL_OK; }
rl_status rl_replay_append(rl_replay *, rl_traj *traj) {
    return rl_fail(traj ? traj->ctx : nullptr, RL_ERR_UNSUPPORTED, "rl_replay_append: not built yet");
}
rl_status rl_replay_stats_of(rl_replay *, rl_replay_stats *) {
    return rl_fail(nullptr, RL_ERR_UNSUPPORTED, "rl_replay_stats_of: not built yet");
}
rl_status rl_dqn_update(rl_replay *, rl_mlp *q, rl_adam *, const rl_dqn_cfg *, rl_opt_stats *) {
    return rl_fail(q ? q->ctx : nullptr, RL_ERR_UNSUPPORTED, "rl_dqn_update: not built yet");
}
double rl_exploration_rate(double start, double end, uint64_t period, uint64_t global_steps, int32_t training) {
    // ExplorationRateSchedule::exploration_rate (schedules.rs:35-45)
    if (!training) return 0.0;
    if (period == 0) return end;
    double frac = (double)global_steps / (double)period;
    if (frac > 1.0) frac = 1.0;
    return frac * (end - start) + start;
}

}  // extern "C"
