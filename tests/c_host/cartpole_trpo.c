/*
 * cartpole_trpo.c -- the `cartpole-trpo` example (examples/cartpole-trpo.rs:14-66: CartPole wrapped in
 * VisibleStepLimit(500), ActorCriticAgent with a TRPO policy and a ValuesOpt critic, train_parallel periods)
 * driven from plain C99 through include/relearn_b200.h alone -- what a Rust `-sys` crate would do over the same
 * symbols.  Test infrastructure: tests/test_c_host.py compiles it with gcc, runs it, and compares
 * every printed number bit for bit with the Python host mirror driving the same library.
 *
 *   cartpole_trpo <weights.bin> <num_envs> <steps_per_period> <periods> <seed>
 *
 * weights.bin: f32 policy parameters (5-128-2) followed by f32 critic parameters (5-128-1), Module::variables()
 * order.  Floating-point results are printed as C99 hex floats (%a) so that the comparison is exact.
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "relearn_b200.h"

static rl_ctx *g_ctx = NULL;

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        rl_status s_ = (call);                                                                        \
        if (s_ != RL_OK && !(s_ >= RL_STEP_NAN_LOSS && s_ <= RL_STEP_CONSTRAINT_VIOLATED)) {          \
            fprintf(stderr, "%s failed: %s (%s)\n", #call, rl_status_str(s_), rl_last_error(g_ctx)); \
            return 10 + (int)s_;                                                                      \
        }                                                                                             \
    } while (0)

static uint64_t fnv1a(const void *data, size_t n) {
    const unsigned char *p = (const unsigned char *)data;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

int main(int argc, char **argv) {
    if (argc != 6) {
        fprintf(stderr, "usage: %s weights.bin num_envs steps_per_period periods seed\n", argv[0]);
        return 2;
    }
    const uint64_t num_envs = strtoull(argv[2], NULL, 10), steps = strtoull(argv[3], NULL, 10);
    const int periods = atoi(argv[4]);
    const uint64_t seed = strtoull(argv[5], NULL, 10);

    printf("version %" PRIu32 "\n", rl_version());
    CHECK(rl_ctx_create(0, NULL, &g_ctx));

    /* CartPole::default().wrap(VisibleStepLimit::new(500))  (examples/cartpole-trpo.rs:23) */
    rl_cartpole_cfg env_cfg;
    rl_cartpole_cfg_default(&env_cfg, 500);
    rl_env *env = NULL;
    CHECK(rl_env_create(g_ctx, RL_ENV_CARTPOLE, &env_cfg, num_envs, 0, seed, &env));
    rl_env_structure st;
    CHECK(rl_env_structure_of(env, &st));
    printf("structure %d %d %a\n", (int)st.num_features, (int)st.num_actions, st.discount_factor);

    /* MlpConfig::default(): one hidden layer of 128, ReLU (mlp.rs:25-34) */
    const int32_t hidden[1] = {128};
    rl_mlp *policy = NULL, *critic = NULL;
    CHECK(rl_mlp_create(g_ctx, st.num_features, hidden, 1, st.num_actions, RL_ACT_RELU, &policy));
    CHECK(rl_mlp_create(g_ctx, st.num_features, hidden, 1, 1, RL_ACT_RELU, &critic));
    uint64_t np = 0, nc = 0;
    CHECK(rl_mlp_num_params(policy, &np));
    CHECK(rl_mlp_num_params(critic, &nc));
    float *w = (float *)malloc((np + nc) * sizeof(float));
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(w, sizeof(float), np + nc, f) != np + nc) {
        fprintf(stderr, "cannot read %" PRIu64 " floats from %s\n", np + nc, argv[1]);
        return 3;
    }
    fclose(f);
    CHECK(rl_mlp_set_weights(policy, w, np));
    CHECK(rl_mlp_set_weights(critic, w + np, nc));

    rl_adam_cfg adam_cfg;
    rl_adam_cfg_default(&adam_cfg);
    rl_adam *adam = NULL;
    CHECK(rl_adam_create(critic, &adam_cfg, &adam));
    rl_trpo_cfg trpo_cfg;
    rl_trpo_cfg_default(&trpo_cfg);

    rl_traj *traj = NULL;
    CHECK(rl_traj_create(env, steps, &traj));
    void *adv = NULL, *rtg = NULL;
    CHECK(rl_malloc(g_ctx, steps * num_envs * sizeof(float), &adv));
    CHECK(rl_malloc(g_ctx, steps * num_envs * sizeof(float), &rtg));

    /* ValuesOptConfig: gamma = min(0.99, env gamma) as f32 (opt.rs:73), lambda 0.95 (critics/mod.rs:78) */
    const float gamma = (float)(st.discount_factor < 0.99 ? st.discount_factor : 0.99), lambda = 0.95f;

    for (int p = 0; p < periods; ++p) {
        /* train_parallel (train.rs:98-158): collect, then batch_update (actor_critic.rs:176-211) */
        rl_actor_cfg actor;
        memset(&actor, 0, sizeof(actor));
        actor.kind = RL_ACTOR_CATEGORICAL_POLICY;
        actor.net = policy;
        actor.training = 1;
        rl_bound bound = {steps, 0};
        rl_steps_summary summ;
        CHECK(rl_rollout(env, &actor, bound, traj, &summ));
        printf("period %d steps %" PRIu64 " episodes %" PRIu64 " step_reward %a %a %" PRIu64 " episode_length %a %a %" PRIu64 "\n",
               p, summ.num_stored_steps, summ.num_stored_episodes, summ.step_reward.mean,
               summ.step_reward.squared_residual_sum, summ.step_reward.count, summ.episode_length.mean,
               summ.episode_length.squared_residual_sum, summ.episode_length.count);

        CHECK(rl_gae(traj, critic, gamma, lambda, (float *)adv, (float *)rtg));
        rl_trpo_stats ts;
        rl_status status = rl_trpo_update(traj, (const float *)adv, policy, &trpo_cfg, &ts);
        CHECK(status);
        printf("trpo %d status %d entropy %a step_size %a loss_initial %a loss_final %a kl %a step_scale %a backtracks %" PRId64
               " cg %" PRId64 " n %" PRIu64 "\n",
               p, (int)status, ts.entropy, ts.step_size, ts.loss_initial, ts.loss_final, ts.constraint_val_final,
               ts.step_scale, ts.num_backtracks, ts.cg_iterations, ts.num_steps);

        CHECK(rl_gae(traj, NULL, gamma, lambda, NULL, (float *)rtg));
        rl_opt_stats os;
        CHECK(rl_value_update(traj, (const float *)rtg, critic, adam, 80, &os));
        printf("critic %d loss_first %a loss_last %a n %" PRIu64 " opt_steps %" PRIu64 "\n", p, os.loss_first, os.loss_last,
               os.num_steps, os.opt_steps);
    }

    CHECK(rl_mlp_get_weights(policy, w, np));
    CHECK(rl_mlp_get_weights(critic, w + np, nc));
    printf("weights policy %016" PRIx64 " critic %016" PRIx64 "\n", fnv1a(w, np * sizeof(float)),
           fnv1a(w + np, nc * sizeof(float)));
    printf("launches %" PRIu64 "\n", rl_ctx_launch_count(g_ctx));

    free(w);
    CHECK(rl_free(g_ctx, adv));
    CHECK(rl_free(g_ctx, rtg));
    CHECK(rl_traj_destroy(traj));
    CHECK(rl_adam_destroy(adam));
    CHECK(rl_mlp_destroy(policy));
    CHECK(rl_mlp_destroy(critic));
    CHECK(rl_env_destroy(env));
    CHECK(rl_ctx_destroy(g_ctx));
    return 0;
}
