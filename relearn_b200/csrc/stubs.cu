// stubs.cu -- entry points declared in relearn_b200.h whose kernels are not built yet.
// They fail loudly (RL_ERR_UNSUPPORTED); nothing falls back to the CPU.
#include "handles.cuh"

extern "C" {

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
