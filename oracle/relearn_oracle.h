/*
 * relearn_oracle.h -- CPU restatement of relearn's rollout/update hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under relearn_b200/ (the product) may
 * include, link or dlopen this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Every function cites the reference file:line (relative to the upstream
 * edlanglois/relearn tree) whose behaviour it restates.  The reference is
 * Rust and cannot be compiled in this image (no rustc/cargo), so this is a
 * restatement, pinned against the known-answer vectors in the reference's
 * own #[test]s (see tests/test_oracle_golden.py).
 *
 * PARITY UNPINNED for the RNG boundary: rand 0.8.5 / rand_chacha 0.3.1 are
 * third-party crates that are not vendored in the reference and no reference
 * test asserts a concrete random value.  The u32/u64 -> sample conversion
 * rules below (ro_gen_f32 ... ro_bernoulli) restate the published rand 0.8.5
 * algorithms; parity is therefore anchored on *replayed word streams*, not on
 * ChaCha seeds.
 *
 * Build with -ffp-contract=off (rustc never contracts a*b+c into an fma).
 */
#ifndef RELEARN_ORACLE_H
#define RELEARN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ */
/* Successor codes  (src/envs/mod.rs:257-269)                          */
/* ------------------------------------------------------------------ */
enum { RO_CONTINUE = 0, RO_TERMINATE = 1, RO_INTERRUPT = 2 };

/* ------------------------------------------------------------------ */
/* Noise source.                                                       */
/*  SCRIPT : a per-lane stream of u32 words consumed sequentially, the  */
/*           way rand_core's BlockRng hands out ChaCha words            */
/*           (next_u64 = two consecutive u32, low word first).          */
/*  PHILOX : counter-based; word(e, t, stream, draw) = Philox4x32-10    */
/*           with key = seed, counter = (e_lo, e_hi, t, stream*64+d/2), */
/*           pair d%2 of the 128-bit output.  Every draw (u32 or u64)   */
/*           takes one 64-bit slot.  Same rule as the CUDA kernels.     */
/* ------------------------------------------------------------------ */
enum { RO_STREAM_ENV_STEP = 0, RO_STREAM_ENV_RESET = 1, RO_STREAM_ACTOR = 2 };
enum { RO_RNG_SCRIPT = 0, RO_RNG_PHILOX = 1 };

typedef struct ro_rng {
    int mode;
    /* SCRIPT */
    const uint32_t *words;
    size_t n_words;
    size_t cursor;
    int exhausted; /* set when a draw ran past n_words (returns 0 words) */
    /* PHILOX */
    uint64_t seed;
    uint64_t lane;      /* global env index */
    uint32_t t;         /* global step counter */
    uint32_t draw[3];   /* per-stream draw index inside (lane, t) */
} ro_rng;

void ro_rng_script(ro_rng *r, const uint32_t *words, size_t n_words);
void ro_rng_philox(ro_rng *r, uint64_t seed, uint64_t lane, uint32_t t);
void ro_rng_set_step(ro_rng *r, uint32_t t); /* PHILOX: move to step t, reset draw indices */
uint32_t ro_next_u32(ro_rng *r, int stream);
uint64_t ro_next_u64(ro_rng *r, int stream);

void ro_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint64_t ro_philox_slot(uint64_t seed, uint64_t lane, uint32_t t, int stream, uint32_t draw);

/* rand 0.8.5 conversions (third-party; see header note) */
float ro_gen_f32(ro_rng *r, int stream);                 /* (u32>>8) * 2^-24 */
double ro_gen_f64(ro_rng *r, int stream);                /* (u64>>11) * 2^-53 */
uint64_t ro_gen_range(ro_rng *r, int stream, uint64_t n);/* sample_single, 0..n */
uint64_t ro_uniform_int(ro_rng *r, int stream, uint64_t n); /* Uniform::new(0,n) */
int ro_gen_bool(ro_rng *r, int stream, double p);        /* Bernoulli */
typedef struct { double low, scale; } ro_uniform_f64;
ro_uniform_f64 ro_uniform_inclusive(double low, double high);
double ro_uniform_sample(const ro_uniform_f64 *u, ro_rng *r, int stream);
double ro_u64_to_uniform(const ro_uniform_f64 *u, uint64_t word);
float ro_u32_to_f32(uint32_t w);
double ro_u64_to_f64(uint64_t w);

/* ------------------------------------------------------------------ */
/* Environments                                                        */
/* ------------------------------------------------------------------ */
enum { RO_ENV_CARTPOLE = 0, RO_ENV_CHAIN = 1, RO_ENV_MEMORY = 2, RO_ENV_BANDIT_META = 3, RO_ENV_PARTITION = 4 };
#define RO_PARTITION_FEATURES 10 /* partition.rs: NUM_FEATURES */
enum { RO_BANDIT_UNIFORM_BERNOULLI = 0, RO_BANDIT_ROUND_ROBIN_DETERMINISTIC = 1, RO_BANDIT_ONE_HOT = 2 };
#define RO_MAX_ARMS 32
#define RO_MAX_FEATURES 64

typedef struct ro_env_cfg {
    int kind;
    /* CartPole: PhysicalConstants + EnvironmentParams (src/envs/cartpole.rs:157-216) */
    double gravity, mass_cart, mass_pole, length_half_pole, friction_cart, friction_pole, time_step;
    double action_force, max_pos, max_angle, discount_factor;
    /* VisibleStepLimit / LatentStepLimit wrapper; 0 = unwrapped (src/envs/wrappers/step_limit.rs) */
    uint64_t max_steps_per_episode;
    int step_limit_visible; /* 1 = VisibleStepLimit (adds `remaining` feature), 0 = Latent */
    /* Chain (src/envs/chain.rs:21-45) */
    uint64_t chain_size;
    /* MemoryGame (src/envs/memory.rs:24-55) */
    uint64_t num_actions, history_len;
    /* MetaEnv<bandits> + TrialEpisodeLimit (src/envs/meta.rs:49-203,541-617) */
    uint64_t num_arms, episodes_per_trial;
    int bandit_dist;
} ro_env_cfg;

void ro_cfg_cartpole_default(ro_env_cfg *c, uint64_t step_limit);
void ro_cfg_chain_default(ro_env_cfg *c);
void ro_cfg_memory(ro_env_cfg *c, uint64_t num_actions, uint64_t history_len);
void ro_cfg_bandit_meta(ro_env_cfg *c, uint64_t num_arms, uint64_t episodes_per_trial, int dist);
void ro_cfg_partition(ro_env_cfg *c);

typedef struct ro_state {
    /* cartpole */
    double x, xd, th, thd;
    int flag;
    uint64_t steps_remaining;
    /* chain / memory; PartitionGame: s = element bits, s_init = supervisor axis, has_prev / prev_action / inner_done =
     * feedback present / its element bits / its label */
    uint64_t s, s_init;
    /* meta bandit */
    double means[RO_MAX_ARMS];
    int inner_done, has_prev;
    uint64_t prev_action;
    double prev_reward;
    uint64_t remaining_episodes;
} ro_state;

typedef struct ro_env {
    ro_env_cfg cfg;
    /* InternalPhysicalConstants (cartpole.rs:238-251) */
    double total_weight, inv_total_mass, mass_length_pole;
    ro_uniform_f64 reset_dist, mean_dist;
    uint64_t rr_good_arm; /* RoundRobinDeterministicBandits cell (envs/testing.rs:108-161) */
} ro_env;

void ro_env_init(ro_env *env, const ro_env_cfg *cfg);
int ro_env_num_features(const ro_env *env);
int ro_env_num_actions(const ro_env *env);
double ro_env_discount(const ro_env *env);
void ro_env_reward_range(const ro_env *env, double *lo, double *hi);
int ro_env_num_observations(const ro_env *env); /* finite obs spaces (chain, memory); 0 otherwise */

void ro_env_initial_state(ro_env *env, ro_state *s, ro_rng *rng);
/* feature row of observe(state) (f32, zero-filled first) */
void ro_env_observe(const ro_env *env, const ro_state *s, float *out);
/* index form of observe(state) for finite observation spaces */
uint64_t ro_env_observe_index(const ro_env *env, const ro_state *s);
/* returns successor code; *s becomes the successor state for CONTINUE/INTERRUPT */
int ro_env_step(ro_env *env, ro_state *s, uint64_t action, ro_rng *rng, double *reward);

/* cartpole.rs:306-387 */
void ro_cartpole_next_state(const ro_env *env, const ro_state *in, double force, ro_state *out);

/* ------------------------------------------------------------------ */
/* MLP (one hidden layer), Categorical                                 */
/*   flat parameter order = Module::variables(): W1[H,F] row-major, b1, */
/*   W2[A,H], b2  (mlp.rs:126-128, linear.rs:108-110)                   */
/* ------------------------------------------------------------------ */
enum { RO_ACT_IDENTITY = 0, RO_ACT_RELU = 1, RO_ACT_SIGMOID = 2, RO_ACT_TANH = 3 };
typedef struct ro_mlp { int in, hidden, out, act; const float *params; } ro_mlp;
size_t ro_mlp_num_params(int in, int hidden, int out);
void ro_mlp_forward(const ro_mlp *m, const float *x, float *out, float *hidden_scratch);
void ro_log_softmax(const float *z, int n, float *out);
/* inverse-CDF categorical sample from logits with a uniform u in [0,1) */
int ro_categorical_sample(const float *logits, int n, float u);

/* ------------------------------------------------------------------ */
/* Actors                                                              */
/* ------------------------------------------------------------------ */
enum {
    RO_ACTOR_REPLAY = 0,         /* scripted actions */
    RO_ACTOR_RANDOM = 1,         /* action_space.sample(rng) = gen_range(0..A) (index.rs:66) */
    RO_ACTOR_POLICY = 2,         /* PolicyActor::act (policies/actor.rs:42-55) */
    RO_ACTOR_EPS_GREEDY_Q = 3,   /* DqnActor::act (dqn.rs:360-379) */
    RO_ACTOR_TABULAR = 4,        /* tabular.rs:222-232 */
    RO_ACTOR_CALLBACK = 5        /* act_fn(act_ud, obs, F): an externally evaluated policy (oracle/aten_actor.cpp) */
};
typedef struct ro_actor {
    int kind;
    const uint8_t *actions; size_t n_actions; size_t cursor; /* REPLAY */
    ro_mlp mlp;                                              /* POLICY / EPS_GREEDY_Q */
    double exploration_rate;                                 /* EPS_GREEDY_Q / TABULAR */
    int training;                                            /* TABULAR: ActorMode::Training */
    const double *q_table; int n_obs, n_act;                 /* TABULAR (row-major [S,A]) */
    uint64_t (*act_fn)(void *ud, const float *obs, int n_features); void *act_ud; /* CALLBACK */
} ro_actor;

/* ------------------------------------------------------------------ */
/* OnlineMeanVariance / StepsSummary (utils/stats.rs:121-209,          */
/* simulation/summary.rs:198-216)                                      */
/* ------------------------------------------------------------------ */
typedef struct ro_omv { double mean, m2; uint64_t count; } ro_omv;
void ro_omv_push(ro_omv *s, double v);
ro_omv ro_omv_add(ro_omv a, ro_omv b);
typedef struct ro_summary {
    ro_omv step_reward, episode_reward, episode_length;
    uint64_t cur_len; double cur_reward;
} ro_summary;
void ro_summary_push(ro_summary *s, double reward, int succ);
void ro_summary_merge(ro_summary *into, const ro_summary *other);

/* ------------------------------------------------------------------ */
/* Rollout of one lane = Steps::step loop + TakeAlignedSteps +          */
/* VecBuffer::write_experience (steps.rs:113-168, take_steps.rs:18-89,  */
/* vec.rs:113-141, buffers/mod.rs:237-261)                              */
/* Output arrays are lane-local, row-major: obs[cap][F], next_obs[cap][F]*/
/* (valid where succ==INTERRUPT).  Returns the number of stored steps   */
/* after finalize_last_episode.                                         */
/* ------------------------------------------------------------------ */
typedef struct ro_lane_out {
    float *obs; uint8_t *action; float *reward; uint8_t *succ; float *next_obs;
    size_t cap;
    size_t n_taken;   /* steps produced by the iterator (before finalize) */
} ro_lane_out;
size_t ro_rollout_lane(ro_env *env, ro_actor *actor, size_t min_steps, size_t slack_steps,
                       ro_rng *rng_env, ro_rng *rng_actor, uint32_t t0,
                       ro_lane_out *out, ro_summary *summary);

/* TakeAlignedSteps over an array of successor codes: number taken (take_steps.rs:77-89) */
size_t ro_take_aligned_steps(const uint8_t *succ, size_t n_avail, size_t min_steps, size_t slack);
/* finalize_last_episode on (succ[], n): returns new n; sets *new_episode (buffers/mod.rs:237-261) */
size_t ro_finalize_last_episode(uint8_t *succ, size_t n, int *new_episode);
/* HistoryDataBound helpers (buffers/mod.rs:57-86) */
size_t ro_default_slack(size_t min_steps);
size_t ro_div_ceil(size_t a, size_t b);

/* ReplayBuffer bookkeeping (buffers/replay.rs:11-126): lengths-only model */
typedef struct ro_replay {
    size_t capacity;
    uint8_t *succ;            /* ring content, oldest first, compacted */
    size_t n;                 /* stored steps */
    uint64_t *episode_ends;   /* absolute end indices (total_step_count based) */
    size_t n_eps, eps_cap;
    uint64_t index_offset, total_step_count;
} ro_replay;
int ro_replay_init(ro_replay *rb, size_t capacity);
void ro_replay_free(ro_replay *rb);
/* returns 0 ok, -1 Full */
int ro_replay_write_step(ro_replay *rb, uint8_t succ);
void ro_replay_end_experience(ro_replay *rb);

/* ------------------------------------------------------------------ */
/* Scans (torch/packed.rs:280-342, critics/mod.rs:101-199)             */
/* ------------------------------------------------------------------ */
/* Packed (time-major interleaved, longest first) in place, with batch sizes */
void ro_discounted_cumsum_packed_f64(double *x, size_t n, const size_t *batch_sizes, size_t n_batches, double d);
void ro_discounted_cumsum_packed_f32(float *x, size_t n, const size_t *batch_sizes, size_t n_batches, float d);
/* One lane of a [T] trajectory with successor codes: y_t = x_t + d*y_{t+1}, restart after episode ends */
void ro_discounted_cumsum_lane_f32(const float *x, const uint8_t *succ, size_t n, float d, float *y);
/* GAE on one lane. v[t] = V(o_t); v_next_intr[t] = V(next obs) where succ[t]==INTERRUPT.
   delta_t = r + gamma*v_ext[t+1] - v[t]  evaluated as (r + gamma*next) - cur in f32. */
void ro_gae_lane_f32(const float *reward, const float *v, const float *v_next_intr, const uint8_t *succ,
                     size_t n, float gamma, float lambda, float *adv);
void ro_td_lane_f32(const float *reward, const float *v, const float *v_next_intr, const uint8_t *succ,
                    size_t n, float gamma, float *delta);

/* ------------------------------------------------------------------ */
/* Tabular Q (agents/tabular.rs:159-179)                               */
/* ------------------------------------------------------------------ */
typedef struct ro_tabq { int n_obs, n_act; double discount; double *q; uint64_t *counts; } ro_tabq;
void ro_tabq_step_update(ro_tabq *t, uint64_t obs, uint64_t action, double reward, int has_next, uint64_t next_obs);
/* fold over one buffer of steps in order (fold_transient, simulation/mod.rs:287-313):
   obs[i], action[i], reward[i], succ[i], next_obs[i] (valid where succ==INTERRUPT) */
void ro_tabq_update_buffer(ro_tabq *t, const uint32_t *obs, const uint8_t *action, const float *reward,
                           const uint8_t *succ, const uint32_t *next_obs, size_t n);
int ro_argmax_f64(const double *row, int n);

/* multi-threaded lane rollouts for the CPU baseline (OpenMP); returns total stored steps */
uint64_t ro_rollout_lanes_philox(const ro_env_cfg *cfg, const ro_mlp *policy, uint64_t n_lanes, uint64_t lane0,
                                 size_t min_steps, size_t slack, uint64_t seed, uint32_t t0,
                                 int n_threads, ro_summary *summary);

#ifdef __cplusplus
}
#endif
#endif
