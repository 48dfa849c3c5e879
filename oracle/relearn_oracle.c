/*
 * relearn_oracle.c -- CPU restatement of relearn's rollout/update hot path.
 * TEST INFRASTRUCTURE ONLY (see relearn_oracle.h).  Compile with
 *   gcc -O2 -ffp-contract=off -pthread -shared -fPIC
 * Citations are file:line in the upstream edlanglois/relearn tree.
 */
#include "relearn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ================================================================== */
/* Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as    */
/* easy as 1, 2, 3"; public algorithm, constants from the paper).       */
/* ================================================================== */
void ro_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint64_t ro_philox_slot(uint64_t seed, uint64_t lane, uint32_t t, int stream, uint32_t draw) {
    uint32_t ctr[4] = {(uint32_t)lane, (uint32_t)(lane >> 32), t, (uint32_t)stream * 64u + (draw >> 1)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    ro_philox4x32_10(ctr, key, o);
    int p = (int)(draw & 1u) * 2;
    return (uint64_t)o[p] | ((uint64_t)o[p + 1] << 32);
}

void ro_rng_script(ro_rng *r, const uint32_t *words, size_t n_words) {
    memset(r, 0, sizeof(*r));
    r->mode = RO_RNG_SCRIPT;
    r->words = words;
    r->n_words = n_words;
}

void ro_rng_philox(ro_rng *r, uint64_t seed, uint64_t lane, uint32_t t) {
    memset(r, 0, sizeof(*r));
    r->mode = RO_RNG_PHILOX;
    r->seed = seed;
    r->lane = lane;
    r->t = t;
}

void ro_rng_set_step(ro_rng *r, uint32_t t) {
    if (r->mode == RO_RNG_PHILOX) {
        r->t = t;
        r->draw[0] = r->draw[1] = r->draw[2] = 0;
    }
}

/* BlockRng::next_u32 / next_u64 (rand_core 0.6.3 block.rs): consecutive
 * u32 words of the block stream; u64 = lo | hi << 32. */
uint32_t ro_next_u32(ro_rng *r, int stream) {
    if (r->mode == RO_RNG_SCRIPT) {
        if (r->cursor >= r->n_words) { r->exhausted = 1; return 0; }
        return r->words[r->cursor++];
    }
    return (uint32_t)ro_philox_slot(r->seed, r->lane, r->t, stream, r->draw[stream]++);
}

uint64_t ro_next_u64(ro_rng *r, int stream) {
    if (r->mode == RO_RNG_SCRIPT) {
        uint64_t lo = ro_next_u32(r, stream);
        uint64_t hi = ro_next_u32(r, stream);
        return lo | (hi << 32);
    }
    return ro_philox_slot(r->seed, r->lane, r->t, stream, r->draw[stream]++);
}

/* ------------------------------------------------------------------ */
/* rand 0.8.5 sampling rules (restated; third-party, parity unpinned)  */
/* ------------------------------------------------------------------ */
float ro_u32_to_f32(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
double ro_u64_to_f64(uint64_t w) { return (double)(w >> 11) * (1.0 / 9007199254740992.0); }

/* Standard f32: 24 bits of a u32 (rand distributions/float.rs) -- call site chain.rs:91 */
float ro_gen_f32(ro_rng *r, int stream) { return ro_u32_to_f32(ro_next_u32(r, stream)); }
/* Standard f64: 53 bits of a u64 -- call site tabular.rs:223 */
double ro_gen_f64(ro_rng *r, int stream) { return ro_u64_to_f64(ro_next_u64(r, stream)); }

static inline uint64_t mulhi64(uint64_t a, uint64_t b, uint64_t *lo) {
    __uint128_t p = (__uint128_t)a * b;
    *lo = (uint64_t)p;
    return (uint64_t)(p >> 64);
}

/* UniformInt::<usize>::sample_single (rand uniform.rs): widening multiply
 * with zone (range << lz) - 1 -- call sites memory.rs:87, tabular.rs:225, index.rs:66 */
uint64_t ro_gen_range(ro_rng *r, int stream, uint64_t n) {
    uint64_t zone = (n << __builtin_clzll(n)) - 1;
    for (;;) {
        uint64_t v = ro_next_u64(r, stream), lo;
        uint64_t hi = mulhi64(v, n, &lo);
        if (lo <= zone) return hi;
        if (r->mode == RO_RNG_SCRIPT && r->exhausted) return 0;
    }
}

/* Uniform::new(0, n).sample (UniformInt::new + sample) -- call site dqn.rs:283 */
uint64_t ro_uniform_int(ro_rng *r, int stream, uint64_t n) {
    uint64_t ints_to_reject = (UINT64_MAX - n + 1) % n;
    uint64_t zone = UINT64_MAX - ints_to_reject;
    for (;;) {
        uint64_t v = ro_next_u64(r, stream), lo;
        uint64_t hi = mulhi64(v, n, &lo);
        if (lo <= zone) return hi;
        if (r->mode == RO_RNG_SCRIPT && r->exhausted) return 0;
    }
}

/* rand::distributions::Bernoulli: p_int = (p * 2^64) as u64, ALWAYS_TRUE for p == 1
 * (no draw consumed) -- call sites dqn.rs:366, utils/distributions.rs:113-120 */
int ro_gen_bool(ro_rng *r, int stream, double p) {
    if (p == 1.0) return 1;
    double scaled = p * 18446744073709551616.0;
    uint64_t p_int = scaled >= 18446744073709551616.0 ? UINT64_MAX : (uint64_t)scaled;
    uint64_t v = ro_next_u64(r, stream);
    return v < p_int;
}

/* UniformFloat<f64>::new_inclusive (rand uniform.rs) -- call sites cartpole.rs:105, bandits.rs:100 */
ro_uniform_f64 ro_uniform_inclusive(double low, double high) {
    const double max_rand = 1.0 - 2.220446049250313e-16; /* 1 - 2^-52 */
    double scale = (high - low) / max_rand;
    while (scale * max_rand + low > high) {
        scale = nextafter(scale, -INFINITY);
    }
    ro_uniform_f64 u = {low, scale};
    return u;
}

double ro_u64_to_uniform(const ro_uniform_f64 *u, uint64_t word) {
    uint64_t bits = (word >> 12) | 0x3FF0000000000000ull; /* [1, 2) */
    double v12;
    memcpy(&v12, &bits, sizeof v12);
    double v01 = v12 - 1.0;
    return v01 * u->scale + u->low;
}

double ro_uniform_sample(const ro_uniform_f64 *u, ro_rng *r, int stream) {
    return ro_u64_to_uniform(u, ro_next_u64(r, stream));
}

/* ================================================================== */
/* Environments                                                        */
/* ================================================================== */
void ro_cfg_cartpole_default(ro_env_cfg *c, uint64_t step_limit) {
    memset(c, 0, sizeof(*c));
    c->kind = RO_ENV_CARTPOLE;
    /* cartpole.rs:178-216 */
    c->gravity = 9.8; c->mass_cart = 1.0; c->mass_pole = 0.1; c->length_half_pole = 0.5;
    c->friction_cart = 0.01; c->friction_pole = 0.01; c->time_step = 0.02;
    c->action_force = 10.0; c->max_pos = 2.4;
    c->max_angle = 12.0 * (3.14159265358979323846264338327950288 / 180.0); /* f64::to_radians */
    c->discount_factor = 0.99;
    c->max_steps_per_episode = step_limit;
    c->step_limit_visible = 1;
}

void ro_cfg_chain_default(ro_env_cfg *c) {
    memset(c, 0, sizeof(*c));
    c->kind = RO_ENV_CHAIN;
    c->chain_size = 5;          /* chain.rs:38-45 */
    c->discount_factor = 0.95;
}

void ro_cfg_memory(ro_env_cfg *c, uint64_t num_actions, uint64_t history_len) {
    memset(c, 0, sizeof(*c));
    c->kind = RO_ENV_MEMORY;
    c->num_actions = num_actions;
    c->history_len = history_len;
    c->discount_factor = 1.0;   /* memory.rs:73-75 */
}

void ro_cfg_bandit_meta(ro_env_cfg *c, uint64_t num_arms, uint64_t episodes_per_trial, int dist) {
    memset(c, 0, sizeof(*c));
    c->kind = RO_ENV_BANDIT_META;
    c->num_arms = num_arms;
    c->episodes_per_trial = episodes_per_trial;
    c->bandit_dist = dist;
    c->discount_factor = 1.0;   /* bandits.rs:53-55 */
}

void ro_cfg_partition(ro_env_cfg *c) {
    memset(c, 0, sizeof(*c));
    c->kind = RO_ENV_PARTITION;
    c->discount_factor = 0.999; /* partition.rs: EnvStructure::discount_factor */
}

void ro_env_init(ro_env *env, const ro_env_cfg *cfg) {
    memset(env, 0, sizeof(*env));
    env->cfg = *cfg;
    if (cfg->kind == RO_ENV_CARTPOLE) {
        /* cartpole.rs:238-251 */
        double total_mass = cfg->mass_cart + cfg->mass_pole;
        env->total_weight = cfg->gravity * total_mass;
        env->inv_total_mass = 1.0 / total_mass;
        env->mass_length_pole = cfg->mass_pole * cfg->length_half_pole;
        env->reset_dist = ro_uniform_inclusive(-0.05, 0.05); /* cartpole.rs:105 */
    }
    if (cfg->kind == RO_ENV_BANDIT_META) {
        env->mean_dist = ro_uniform_inclusive(0.0, 1.0);     /* bandits.rs:100 */
    }
}

static int has_step_limit(const ro_env *env) { return env->cfg.max_steps_per_episode > 0; }

int ro_env_num_features(const ro_env *env) {
    const ro_env_cfg *c = &env->cfg;
    int f = 0;
    switch (c->kind) {
    case RO_ENV_CARTPOLE: f = 4; break;                       /* cartpole.rs:273-284 */
    case RO_ENV_CHAIN: f = (int)c->chain_size; break;         /* index.rs:97-115 one-hot */
    case RO_ENV_MEMORY: f = (int)(c->num_actions + c->history_len); break;
    case RO_ENV_BANDIT_META: return (int)c->num_arms + 4;     /* meta.rs:357-363 */
    /* TupleSpace2<PowerSpace<Boolean, 10>, OptionSpace<TupleSpace2<PowerSpace<Boolean, 10>, IndexedTypeSpace<Classification>>>> */
    case RO_ENV_PARTITION: return 2 * RO_PARTITION_FEATURES + 3;
    }
    if (has_step_limit(env) && c->step_limit_visible) f += 1; /* step_limit.rs:133-138 */
    return f;
}

int ro_env_num_actions(const ro_env *env) {
    switch (env->cfg.kind) {
    case RO_ENV_CARTPOLE: return 2;
    case RO_ENV_CHAIN: return 2;
    case RO_ENV_MEMORY: return (int)env->cfg.num_actions;
    case RO_ENV_BANDIT_META: return (int)env->cfg.num_arms;
    case RO_ENV_PARTITION: return 2;                          /* Action::{ClassifyLeft, ClassifyRight} */
    }
    return 0;
}

int ro_env_num_observations(const ro_env *env) {
    switch (env->cfg.kind) {
    case RO_ENV_CHAIN: return (int)env->cfg.chain_size;
    case RO_ENV_MEMORY: return (int)(env->cfg.num_actions + env->cfg.history_len);
    }
    return 0;
}

double ro_env_discount(const ro_env *env) { return env->cfg.discount_factor; }

void ro_env_reward_range(const ro_env *env, double *lo, double *hi) {
    switch (env->cfg.kind) {
    case RO_ENV_CARTPOLE: *lo = 0.0; *hi = 1.0; break;        /* cartpole.rs:88-90 */
    case RO_ENV_CHAIN: *lo = 0.0; *hi = 10.0; break;          /* chain.rs:60-62 */
    case RO_ENV_MEMORY: *lo = -1.0; *hi = 1.0; break;         /* memory.rs:69-71 */
    case RO_ENV_PARTITION: *lo = -1.0; *hi = 1.0; break;      /* partition.rs feedback_space */
    default: *lo = 0.0; *hi = 1.0; break;                     /* bandits.rs:160-162 */
    }
}

/* cartpole.rs:398-431 */
static double cp_angular_acceleration(const ro_env *e, double thd, double force, double mu,
                                      double w2, double s, double c) {
    const ro_env_cfg *k = &e->cfg;
    double alpha = (-force - e->mass_length_pole * w2 * (s + mu * c)) * e->inv_total_mass;
    double beta = k->friction_pole * thd / e->mass_length_pole;
    double numerator = k->gravity * s + c * (alpha + k->gravity * mu) - beta;
    double denominator =
        k->length_half_pole * (4.0 / 3.0 - k->mass_pole * c * e->inv_total_mass * (c - mu));
    return numerator / denominator;
}

/* cartpole.rs:436-446 */
static double cp_normal_force(const ro_env *e, double acc, double w2, double s, double c) {
    return e->total_weight - e->mass_length_pole * (acc * s + w2 * c);
}

/* cartpole.rs:306-387 */
void ro_cartpole_next_state(const ro_env *e, const ro_state *in, double force, ro_state *out) {
    const ro_env_cfg *k = &e->cfg;
    double mu = in->flag ? k->friction_cart : -k->friction_cart;
    double s = sin(in->th), c = cos(in->th);  /* f64::sin_cos */
    double w2 = in->thd * in->thd;
    double acc = cp_angular_acceleration(e, in->thd, force, mu, w2, s, c);
    double nf = cp_normal_force(e, acc, w2, s, c);
    int positive = !signbit(nf * in->xd);     /* is_sign_positive: sign bit clear */
    if (positive != in->flag) {
        mu = -mu;
        acc = cp_angular_acceleration(e, in->thd, force, mu, w2, s, c);
        nf = cp_normal_force(e, acc, w2, s, c);
    }
    double force_pole = e->mass_length_pole * (w2 * s + acc * c);
    double force_friction = -mu * nf;
    double net = force + force_pole + force_friction;
    double xacc = net * e->inv_total_mass;
    double xd = in->xd + k->time_step * xacc;
    double x = in->x + k->time_step * xd;
    double thd = in->thd + k->time_step * acc;
    double th = in->th + k->time_step * in->thd;
    *out = *in;
    out->x = x; out->xd = xd; out->th = th; out->thd = thd; out->flag = positive;
}

/* rng.gen::<[bool; 10]>(): elements in index order, each `(next_u32() as i32) < 0` (rand 0.8.5 Standard for bool) */
static uint64_t partition_element(ro_rng *rng, int stream) {
    uint64_t bits = 0;
    for (int i = 0; i < RO_PARTITION_FEATURES; ++i)
        if (ro_next_u32(rng, stream) >> 31) bits |= 1ull << i;
    return bits;
}

void ro_env_initial_state(ro_env *env, ro_state *s, ro_rng *rng) {
    const ro_env_cfg *c = &env->cfg;
    memset(s, 0, sizeof(*s));
    switch (c->kind) {
    case RO_ENV_CARTPOLE:                      /* cartpole.rs:103-115 */
        s->x = ro_uniform_sample(&env->reset_dist, rng, RO_STREAM_ENV_RESET);
        s->xd = ro_uniform_sample(&env->reset_dist, rng, RO_STREAM_ENV_RESET);
        s->th = ro_uniform_sample(&env->reset_dist, rng, RO_STREAM_ENV_RESET);
        s->thd = ro_uniform_sample(&env->reset_dist, rng, RO_STREAM_ENV_RESET);
        s->flag = 1;
        break;
    case RO_ENV_CHAIN:                         /* chain.rs:75-77 */
        s->s = 0;
        break;
    case RO_ENV_MEMORY:                        /* memory.rs:86-89 */
        s->s = ro_gen_range(rng, RO_STREAM_ENV_RESET, c->num_actions);
        s->s_init = s->s;
        break;
    case RO_ENV_BANDIT_META:                   /* meta.rs:141-150, bandits.rs:98-105, meta.rs:582-592 */
        if (c->bandit_dist == RO_BANDIT_UNIFORM_BERNOULLI) {
            for (uint64_t i = 0; i < c->num_arms; ++i)
                s->means[i] = ro_uniform_sample(&env->mean_dist, rng, RO_STREAM_ENV_RESET);
        } else if (c->bandit_dist == RO_BANDIT_ONE_HOT) {  /* OneHotBandits::sample_environment, bandits.rs:236-241 */
            for (uint64_t i = 0; i < c->num_arms; ++i) s->means[i] = 0.0;
            s->means[ro_gen_range(rng, RO_STREAM_ENV_RESET, c->num_arms)] = 1.0;
        } else {                               /* envs/testing.rs:147-160 */
            for (uint64_t i = 0; i < c->num_arms; ++i) s->means[i] = 0.0;
            s->means[env->rr_good_arm] = 1.0;
            env->rr_good_arm = (env->rr_good_arm + 1) % c->num_arms;
        }
        s->inner_done = 0;
        s->has_prev = 0;
        s->remaining_episodes = c->episodes_per_trial;
        break;
    case RO_ENV_PARTITION:                     /* partition.rs initial_state: supervisor axis, first element, no feedback */
        s->s_init = ro_gen_range(rng, RO_STREAM_ENV_RESET, RO_PARTITION_FEATURES);
        s->s = partition_element(rng, RO_STREAM_ENV_RESET);
        s->has_prev = 0;
        break;
    }
    s->steps_remaining = c->max_steps_per_episode; /* step_limit.rs:187-192 */
}

uint64_t ro_env_observe_index(const ro_env *env, const ro_state *s) {
    (void)env;
    return s->s; /* chain.rs:79-81, memory.rs:91-94 */
}

void ro_env_observe(const ro_env *env, const ro_state *s, float *out) {
    const ro_env_cfg *c = &env->cfg;
    int f = ro_env_num_features(env);
    for (int i = 0; i < f; ++i) out[i] = 0.0f;
    int pos = 0;
    switch (c->kind) {
    case RO_ENV_CARTPOLE:                      /* interval.rs:101-117: [x as f32] per field, derive order */
        out[0] = (float)s->x; out[1] = (float)s->xd; out[2] = (float)s->th; out[3] = (float)s->thd;
        pos = 4;
        break;
    case RO_ENV_CHAIN:
    case RO_ENV_MEMORY:                        /* index.rs:97-115 one-hot */
        out[s->s] = 1.0f;
        pos = ro_env_num_observations(env);
        break;
    case RO_ENV_BANDIT_META: {                 /* meta.rs:152-163,357-363; option.rs:88-116 */
        int k = (int)c->num_arms;
        out[0] = s->inner_done ? 1.0f : 0.0f;  /* Option<()>: [is_none] */
        if (!s->has_prev) {
            out[1] = 1.0f;                     /* None: [1, 0...] */
        } else {
            out[1] = 0.0f;
            out[2 + s->prev_action] = 1.0f;    /* one-hot action */
            out[2 + k] = (float)s->prev_reward;/* Interval<Reward> */
        }
        out[3 + k] = s->inner_done ? 1.0f : 0.0f; /* boolean.rs:125-139 */
        return;
    }
    case RO_ENV_PARTITION: {                   /* (element, feedback): power.rs:106-115, option.rs:88-116, one-hot label */
        const int n = RO_PARTITION_FEATURES;
        for (int i = 0; i < n; ++i) out[i] = (s->s >> i) & 1 ? 1.0f : 0.0f;
        if (!s->has_prev) {
            out[n] = 1.0f;
        } else {
            for (int i = 0; i < n; ++i) out[n + 1 + i] = (s->prev_action >> i) & 1 ? 1.0f : 0.0f;
            out[2 * n + 1 + (s->inner_done ? 1 : 0)] = 1.0f;   /* Classification::{Left, Right} */
        }
        return;
    }
    }
    if (has_step_limit(env) && c->step_limit_visible) {
        /* step_limit.rs:194-200 */
        double remaining = (double)s->steps_remaining / (double)c->max_steps_per_episode;
        out[pos] = (float)remaining;
    }
}

int ro_env_step(ro_env *env, ro_state *s, uint64_t action, ro_rng *rng, double *reward) {
    const ro_env_cfg *c = &env->cfg;
    int succ = RO_CONTINUE;
    switch (c->kind) {
    case RO_ENV_CARTPOLE: {                    /* cartpole.rs:128-153 */
        double force = action == 0 ? -c->action_force : c->action_force;
        ro_state next;
        ro_cartpole_next_state(env, s, force, &next);
        *reward = 1.0;
        int terminal = fabs(next.x) > c->max_pos || fabs(next.th) > c->max_angle;
        if (terminal) return RO_TERMINATE;
        *s = next;
        break;
    }
    case RO_ENV_CHAIN: {                       /* chain.rs:83-105 */
        uint64_t a = action;
        if (ro_gen_f32(rng, RO_STREAM_ENV_STEP) < 0.2f) a = 1 - a;
        if (a == 0) { s->s = 0; *reward = 2.0; }
        else if (s->s == c->chain_size - 1) { *reward = 10.0; }
        else { s->s += 1; *reward = 0.0; }
        break;
    }
    case RO_ENV_MEMORY: {                      /* memory.rs:96-115 */
        if (s->s == c->num_actions + c->history_len - 1) {
            *reward = action == s->s_init ? 1.0 : -1.0;
            return RO_TERMINATE;
        }
        s->s = s->s < c->num_actions ? c->num_actions : s->s + 1;
        *reward = 0.0;
        break;
    }
    case RO_ENV_PARTITION: {                   /* partition.rs step: never ends */
        const int label = (int)((s->s >> s->s_init) & 1);      /* Supervisor::AxisAligned(axis).classify */
        *reward = (uint64_t)label == action ? 1.0 : -1.0;
        s->prev_action = s->s;
        s->inner_done = label;
        s->has_prev = 1;
        s->s = partition_element(rng, RO_STREAM_ENV_STEP);
        break;
    }
    case RO_ENV_BANDIT_META: {                 /* meta.rs:165-202 + TrialEpisodeLimit meta.rs:594-616 */
        if (!s->inner_done) {
            double r;
            if (c->bandit_dist == RO_BANDIT_UNIFORM_BERNOULLI) {
                /* bandits.rs:75 -> utils/distributions.rs:113-120 */
                r = ro_gen_bool(rng, RO_STREAM_ENV_STEP, s->means[action]) ? 1.0 : 0.0;
            } else {
                r = s->means[action];
            }
            s->inner_done = 1;                 /* Bandit::step always Terminate (bandits.rs:76) */
            s->has_prev = 1;
            s->prev_action = action;
            s->prev_reward = r;
            *reward = r;
        } else {
            s->inner_done = 0;                 /* new inner episode; Bandit::initial_state draws nothing */
            s->has_prev = 0;
            *reward = 0.0;                     /* neutral_outer (meta.rs:274-276) */
        }
        if (s->inner_done) s->remaining_episodes -= 1;
        if (s->remaining_episodes == 0) return RO_INTERRUPT;
        return RO_CONTINUE;
    }
    }
    /* step limit wrapper: decrement on Continue, interrupt at 0 (step_limit.rs:202-223) */
    if (has_step_limit(env) && succ == RO_CONTINUE) {
        s->steps_remaining -= 1;
        if (s->steps_remaining == 0) succ = RO_INTERRUPT;
    }
    return succ;
}

/* ================================================================== */
/* MLP / categorical                                                   */
/* ================================================================== */
size_t ro_mlp_num_params(int in, int hidden, int out) {
    return (size_t)hidden * in + hidden + (size_t)out * hidden + out;
}

static float apply_act(int act, float v) {
    switch (act) {
    case RO_ACT_RELU: return v > 0.0f ? v : 0.0f;
    case RO_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    case RO_ACT_TANH: return tanhf(v);
    default: return v;
    }
}

/* mlp.rs:139-151: for each hidden layer linear -> activation; output linear -> identity.
 * linear.rs:118-123: x . W^T + b with W[out, in]. */
void ro_mlp_forward(const ro_mlp *m, const float *x, float *out, float *h) {
    const float *w1 = m->params, *b1 = w1 + (size_t)m->hidden * m->in;
    const float *w2 = b1 + m->hidden, *b2 = w2 + (size_t)m->out * m->hidden;
    for (int j = 0; j < m->hidden; ++j) {
        float acc = b1[j];
        for (int i = 0; i < m->in; ++i) acc += w1[(size_t)j * m->in + i] * x[i];
        h[j] = apply_act(m->act, acc);
    }
    for (int k = 0; k < m->out; ++k) {
        float acc = b2[k];
        for (int j = 0; j < m->hidden; ++j) acc += w2[(size_t)k * m->hidden + j] * h[j];
        out[k] = acc;
    }
}

/* categorical.rs:29-33 */
void ro_log_softmax(const float *z, int n, float *out) {
    float m = z[0];
    for (int i = 1; i < n; ++i) m = z[i] > m ? z[i] : m;
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) sum += expf(z[i] - m);
    float lse = m + logf(sum);
    for (int i = 0; i < n; ++i) out[i] = z[i] - lse;
}

/* categorical.rs:52-54 samples from exp(log_probs) with libtorch's multinomial, whose
 * uniform comes from libtorch's global generator (not seedable from relearn).  Restated
 * as inverse-CDF over p = exp(log_softmax(z)) in index order with a supplied uniform. */
int ro_categorical_sample(const float *logits, int n, float u) {
    float lp[RO_MAX_FEATURES];
    ro_log_softmax(logits, n, lp);
    float c = 0.0f;
    for (int i = 0; i < n; ++i) {
        c += expf(lp[i]);
        if (u < c) return i;
    }
    return n - 1;
}

int ro_argmax_f64(const double *row, int n) {
    int best = 0; /* first maximal index (ndarray-stats argmax; from memory, unpinned) */
    for (int i = 1; i < n; ++i) if (row[i] > row[best]) best = i;
    return best;
}

static int argmax_f32(const float *row, int n) {
    int best = 0; /* torch argmax returns the first maximal index on CPU */
    for (int i = 1; i < n; ++i) if (row[i] > row[best]) best = i;
    return best;
}

static uint64_t actor_act(ro_actor *a, const ro_env *env, const ro_state *st, const float *obs, ro_rng *rng) {
    int n_act = ro_env_num_actions(env);
    float z[RO_MAX_FEATURES], h[1024];
    switch (a->kind) {
    case RO_ACTOR_REPLAY:
        return a->cursor < a->n_actions ? a->actions[a->cursor++] : 0;
    case RO_ACTOR_RANDOM:                      /* index.rs:66 / indexed_type.rs:146 */
        return ro_gen_range(rng, RO_STREAM_ACTOR, (uint64_t)n_act);
    case RO_ACTOR_POLICY:                      /* policies/actor.rs:42-55 */
        ro_mlp_forward(&a->mlp, obs, z, h);
        return (uint64_t)ro_categorical_sample(z, n_act, ro_gen_f32(rng, RO_STREAM_ACTOR));
    case RO_ACTOR_EPS_GREEDY_Q:                /* dqn.rs:360-379 */
        if (ro_gen_bool(rng, RO_STREAM_ACTOR, a->exploration_rate))
            return ro_gen_range(rng, RO_STREAM_ACTOR, (uint64_t)n_act);
        ro_mlp_forward(&a->mlp, obs, z, h);
        return (uint64_t)argmax_f32(z, n_act);
    case RO_ACTOR_TABULAR:                     /* tabular.rs:222-232 */
        if (a->training && ro_gen_f64(rng, RO_STREAM_ACTOR) < a->exploration_rate)
            return ro_gen_range(rng, RO_STREAM_ACTOR, (uint64_t)a->n_act);
        return (uint64_t)ro_argmax_f64(a->q_table + ro_env_observe_index(env, st) * a->n_act, a->n_act);
    case RO_ACTOR_CALLBACK:
        return a->act_fn(a->act_ud, obs, ro_env_num_features(env));
    }
    return 0;
}

/* ================================================================== */
/* Summary statistics                                                  */
/* ================================================================== */
void ro_omv_push(ro_omv *s, double v) {        /* stats.rs:122-128 */
    double pre = v - s->mean;
    s->count += 1;
    s->mean = s->mean + pre / (double)s->count;
    double post = v - s->mean;
    s->m2 = s->m2 + pre * post;
}

ro_omv ro_omv_add(ro_omv a, ro_omv b) {        /* stats.rs:184-209 (Chan et al.) */
    ro_omv o;
    double ac = (double)a.count, bc = (double)b.count;
    o.count = a.count + b.count;
    double tc = (double)o.count;
    if (o.count == 0) { o.mean = 0.0; o.m2 = 0.0; return o; } /* 0/0 guard; reference yields NaN mean */
    o.mean = (a.mean * ac + b.mean * bc) / tc;
    double delta = a.mean - b.mean;
    o.m2 = a.m2 + b.m2 + delta * delta * ac * bc / tc;
    return o;
}

void ro_summary_push(ro_summary *s, double reward, int succ) { /* summary.rs:198-216 */
    ro_omv_push(&s->step_reward, reward);
    s->cur_len += 1;
    s->cur_reward += reward;
    if (succ != RO_CONTINUE) {
        ro_omv_push(&s->episode_reward, s->cur_reward);
        s->cur_reward = 0.0;
        ro_omv_push(&s->episode_length, (double)s->cur_len);
        s->cur_len = 0;
    }
}

void ro_summary_merge(ro_summary *into, const ro_summary *o) { /* summary.rs:96-117 */
    into->step_reward = ro_omv_add(into->step_reward, o->step_reward);
    into->episode_reward = ro_omv_add(into->episode_reward, o->episode_reward);
    into->episode_length = ro_omv_add(into->episode_length, o->episode_length);
}

/* ================================================================== */
/* Step iterators and buffers                                          */
/* ================================================================== */
size_t ro_default_slack(size_t min_steps) {    /* buffers/mod.rs:57-63 */
    size_t s = min_steps / 100;
    return s < 5 ? 5 : (s > 1000 ? 1000 : s);
}

size_t ro_div_ceil(size_t a, size_t b) { return a / b + (a % b ? 1 : 0); } /* buffers/mod.rs:117-124 */

size_t ro_take_aligned_steps(const uint8_t *succ, size_t n_avail, size_t min_steps, size_t slack) {
    /* take_steps.rs:18-31,77-89 */
    size_t n = min_steps == 0 ? 0 : min_steps + slack, taken = 0;
    while (n > 0 && taken < n_avail) {
        int done = succ[taken] != RO_CONTINUE;
        taken += 1;
        n -= 1;
        if (done && n <= slack) n = 0;
    }
    return taken;
}

size_t ro_finalize_last_episode(uint8_t *succ, size_t n, int *new_episode) {
    /* buffers/mod.rs:237-261 */
    *new_episode = 0;
    if (n == 0 || succ[n - 1] != RO_CONTINUE) return n;
    n -= 1; /* pop; its observation becomes the Interrupt payload of the new last step */
    if (n > 0 && succ[n - 1] == RO_CONTINUE) {
        succ[n - 1] = RO_INTERRUPT;
        *new_episode = 1;
    }
    return n;
}

size_t ro_rollout_lane(ro_env *env, ro_actor *actor, size_t min_steps, size_t slack_steps,
                       ro_rng *rng_env, ro_rng *rng_actor, uint32_t t0,
                       ro_lane_out *out, ro_summary *summary) {
    int F = ro_env_num_features(env);
    size_t n = min_steps == 0 ? 0 : min_steps + slack_steps; /* take_steps.rs:20-31 */
    size_t i = 0;
    int have_state = 0;
    ro_state st;
    float obs[RO_MAX_FEATURES];
    while (n > 0 && i < out->cap) {
        ro_rng_set_step(rng_env, t0 + (uint32_t)i);
        if (rng_actor != rng_env) ro_rng_set_step(rng_actor, t0 + (uint32_t)i);
        /* steps.rs:116-124 */
        if (!have_state) {
            ro_env_initial_state(env, &st, rng_env);
            ro_env_observe(env, &st, obs);
            have_state = 1;
        }
        /* steps.rs:126-128 */
        uint64_t action = actor_act(actor, env, &st, obs, rng_actor);
        memcpy(out->obs + i * F, obs, sizeof(float) * F);
        /* steps.rs:130-135 */
        double reward = 0.0;
        int succ = ro_env_step(env, &st, action, rng_env, &reward);
        /* steps.rs:139-159 */
        if (succ == RO_CONTINUE) {
            ro_env_observe(env, &st, obs);
        } else if (succ == RO_INTERRUPT) {
            ro_env_observe(env, &st, out->next_obs + i * F);
            have_state = 0;
        } else {
            have_state = 0;
        }
        out->action[i] = (uint8_t)action;
        out->reward[i] = (float)reward;        /* features.rs:202 f64 -> f32 */
        out->succ[i] = (uint8_t)succ;
        if (summary) ro_summary_push(summary, reward, succ); /* train.rs:132-140 */
        i += 1;
        /* take_steps.rs:83-88 */
        n -= 1;
        if (succ != RO_CONTINUE && n <= slack_steps) n = 0;
    }
    out->n_taken = i;
    /* vec.rs:137-141 -> buffers/mod.rs:237-261 */
    if (i > 0 && out->succ[i - 1] == RO_CONTINUE) {
        i -= 1; /* popped step: its observation is the Interrupt payload */
        if (i > 0 && out->succ[i - 1] == RO_CONTINUE) {
            out->succ[i - 1] = RO_INTERRUPT;
            memcpy(out->next_obs + (i - 1) * F, out->obs + i * F, sizeof(float) * F);
        }
    }
    return i;
}

/* ---------------- ReplayBuffer (lengths-only model) ---------------- */
int ro_replay_init(ro_replay *rb, size_t capacity) {
    memset(rb, 0, sizeof(*rb));
    rb->capacity = capacity;
    rb->succ = (uint8_t *)malloc(capacity ? capacity : 1);
    rb->eps_cap = 16;
    rb->episode_ends = (uint64_t *)malloc(sizeof(uint64_t) * rb->eps_cap);
    return rb->succ && rb->episode_ends ? 0 : -1;
}

void ro_replay_free(ro_replay *rb) { free(rb->succ); free(rb->episode_ends); memset(rb, 0, sizeof(*rb)); }

static void replay_push_end(ro_replay *rb, uint64_t end) {
    if (rb->n_eps == rb->eps_cap) {
        rb->eps_cap *= 2;
        rb->episode_ends = (uint64_t *)realloc(rb->episode_ends, sizeof(uint64_t) * rb->eps_cap);
    }
    rb->episode_ends[rb->n_eps++] = end;
}

int ro_replay_write_step(ro_replay *rb, uint8_t succ) { /* replay.rs:89-113 */
    if (rb->n == rb->capacity) {
        if (rb->n_eps == 0) return -1; /* WriteExperienceError::Full */
        uint64_t ep_end = rb->episode_ends[0];
        memmove(rb->episode_ends, rb->episode_ends + 1, sizeof(uint64_t) * (rb->n_eps - 1));
        rb->n_eps -= 1;
        size_t ep_len = (size_t)(ep_end - rb->index_offset);
        memmove(rb->succ, rb->succ + ep_len, rb->n - ep_len);
        rb->n -= ep_len;
        rb->index_offset = ep_end;
    }
    rb->succ[rb->n++] = succ;
    rb->total_step_count += 1;
    if (succ != RO_CONTINUE) replay_push_end(rb, rb->total_step_count);
    return 0;
}

void ro_replay_end_experience(ro_replay *rb) { /* replay.rs:115-125 */
    int new_ep = 0;
    size_t before = rb->n;
    rb->n = ro_finalize_last_episode(rb->succ, rb->n, &new_ep);
    if (new_ep) {
        rb->total_step_count -= 1;
        replay_push_end(rb, rb->total_step_count);
    }
    (void)before;
}

/* ================================================================== */
/* Scans                                                               */
/* ================================================================== */
/* packed.rs:312-342: walk batches from the end; a[offset + i] += b[i] * discount */
#define RO_CUMSUM_PACKED(NAME, T)                                                              \
    void NAME(T *x, size_t n, const size_t *batch_sizes, size_t n_batches, T d) {              \
        size_t offset = n, prev_start = n, prev_size = 0;                                      \
        for (size_t bi = n_batches; bi-- > 0;) {                                               \
            size_t bs = batch_sizes[bi];                                                       \
            offset -= bs;                                                                      \
            for (size_t i = 0; i < prev_size; ++i) {                                           \
                T prod = x[prev_start + i] * d;                                                \
                x[offset + i] += prod;                                                         \
            }                                                                                  \
            prev_start = offset;                                                               \
            prev_size = bs;                                                                    \
        }                                                                                      \
    }
RO_CUMSUM_PACKED(ro_discounted_cumsum_packed_f64, double)
RO_CUMSUM_PACKED(ro_discounted_cumsum_packed_f32, float)

void ro_discounted_cumsum_lane_f32(const float *x, const uint8_t *succ, size_t n, float d, float *y) {
    float carry = 0.0f;
    for (size_t i = n; i-- > 0;) {
        if (succ[i] != RO_CONTINUE) carry = 0.0f; /* last step of its episode */
        float prod = carry * d;                   /* packed.rs:336  *a += *b * discount */
        carry = x[i] + prod;
        y[i] = carry;
    }
}

/* critics/mod.rs:158-174: rewards + discount * next - cur, f32, left to right */
void ro_td_lane_f32(const float *reward, const float *v, const float *v_next_intr, const uint8_t *succ,
                    size_t n, float gamma, float *delta) {
    for (size_t i = 0; i < n; ++i) {
        float next;
        if (succ[i] == RO_CONTINUE) next = i + 1 < n ? v[i + 1] : 0.0f;
        else if (succ[i] == RO_INTERRUPT) next = v_next_intr[i]; /* features.rs:139-185 */
        else next = 0.0f;                                        /* masked_fill_(is_invalid, 0) */
        float gn = gamma * next;
        float s = reward[i] + gn;
        delta[i] = s - v[i];
    }
}

/* critics/mod.rs:190-199 */
void ro_gae_lane_f32(const float *reward, const float *v, const float *v_next_intr, const uint8_t *succ,
                     size_t n, float gamma, float lambda, float *adv) {
    float *delta = (float *)malloc(sizeof(float) * (n ? n : 1));
    ro_td_lane_f32(reward, v, v_next_intr, succ, n, gamma, delta);
    float d = lambda * gamma; /* f32 product, critics/mod.rs:198 */
    ro_discounted_cumsum_lane_f32(delta, succ, n, d, adv);
    free(delta);
}

/* ================================================================== */
/* Tabular Q                                                           */
/* ================================================================== */
void ro_tabq_step_update(ro_tabq *t, uint64_t obs, uint64_t action, double reward, int has_next, uint64_t next_obs) {
    /* tabular.rs:159-179 */
    double discounted_next = 0.0;
    if (has_next) {
        const double *row = t->q + next_obs * t->n_act;
        double m = row[0];
        for (int i = 1; i < t->n_act; ++i) if (row[i] > m) m = row[i];
        discounted_next = m * t->discount;
    }
    size_t idx = obs * t->n_act + action;
    t->counts[idx] += 1;
    double value = reward + discounted_next;
    double weight = 1.0 / (double)t->counts[idx];
    t->q[idx] *= 1.0 - weight;
    t->q[idx] += weight * value;
}

void ro_tabq_update_buffer(ro_tabq *t, const uint32_t *obs, const uint8_t *action, const float *reward,
                           const uint8_t *succ, const uint32_t *next_obs, size_t n) {
    /* simulation/mod.rs:287-313 fold_transient: Continue borrows the following step's observation;
       a trailing Continue with no successor is skipped. */
    for (size_t i = 0; i < n; ++i) {
        if (succ[i] == RO_CONTINUE) {
            if (i + 1 >= n) break;
            ro_tabq_step_update(t, obs[i], action[i], (double)reward[i], 1, obs[i + 1]);
        } else if (succ[i] == RO_INTERRUPT) {
            ro_tabq_step_update(t, obs[i], action[i], (double)reward[i], 1, next_obs[i]);
        } else {
            ro_tabq_step_update(t, obs[i], action[i], (double)reward[i], 0, 0);
        }
    }
}

/* ================================================================== */
/* Multi-threaded rollouts: the reference-equivalent CPU baseline.     */
/* One lane = one train_parallel worker period (train.rs:124-158);     */
/* lanes are spread over n_threads OS threads like the reference's     */
/* crossbeam scope.                                                    */
/* ================================================================== */
typedef struct lanes_job {
    const ro_env_cfg *cfg; const ro_mlp *policy;
    uint64_t lane_begin, lane_end, lane0;
    size_t min_steps, slack; uint64_t seed; uint32_t t0;
    uint64_t total; ro_summary summary;
} lanes_job;

static void *lanes_worker(void *arg) {
    lanes_job *job = (lanes_job *)arg;
    ro_env env;
    ro_env_init(&env, job->cfg);
    int F = ro_env_num_features(&env);
    size_t cap = job->min_steps + job->slack;
    ro_lane_out out;
    out.cap = cap;
    out.obs = (float *)malloc(sizeof(float) * cap * F);
    out.next_obs = (float *)malloc(sizeof(float) * cap * F);
    out.action = (uint8_t *)malloc(cap);
    out.reward = (float *)malloc(sizeof(float) * cap);
    out.succ = (uint8_t *)malloc(cap);
    memset(&job->summary, 0, sizeof(job->summary));
    job->total = 0;
    for (uint64_t l = job->lane_begin; l < job->lane_end; ++l) {
        ro_actor actor;
        memset(&actor, 0, sizeof(actor));
        actor.kind = job->policy ? RO_ACTOR_POLICY : RO_ACTOR_RANDOM;
        if (job->policy) actor.mlp = *job->policy;
        ro_rng rng;
        ro_rng_philox(&rng, job->seed, job->lane0 + l, job->t0);
        ro_summary lane_summary;
        memset(&lane_summary, 0, sizeof(lane_summary));
        job->total += ro_rollout_lane(&env, &actor, job->min_steps, job->slack, &rng, &rng, job->t0, &out,
                                      &lane_summary);
        ro_summary_merge(&job->summary, &lane_summary);
    }
    free(out.obs); free(out.next_obs); free(out.action); free(out.reward); free(out.succ);
    return NULL;
}

uint64_t ro_rollout_lanes_philox(const ro_env_cfg *cfg, const ro_mlp *policy, uint64_t n_lanes, uint64_t lane0,
                                 size_t min_steps, size_t slack, uint64_t seed, uint32_t t0,
                                 int n_threads, ro_summary *summary) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_lanes && n_lanes > 0) n_threads = (int)n_lanes;
    lanes_job *jobs = (lanes_job *)calloc((size_t)n_threads, sizeof(lanes_job));
    pthread_t *threads = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 0; i < n_threads; ++i) {
        jobs[i].cfg = cfg; jobs[i].policy = policy; jobs[i].lane0 = lane0;
        jobs[i].lane_begin = n_lanes * (uint64_t)i / (uint64_t)n_threads;
        jobs[i].lane_end = n_lanes * (uint64_t)(i + 1) / (uint64_t)n_threads;
        jobs[i].min_steps = min_steps; jobs[i].slack = slack; jobs[i].seed = seed; jobs[i].t0 = t0;
        pthread_create(&threads[i], NULL, lanes_worker, &jobs[i]);
    }
    uint64_t total = 0;
    ro_summary merged;
    memset(&merged, 0, sizeof(merged));
    for (int i = 0; i < n_threads; ++i) {
        pthread_join(threads[i], NULL);
        total += jobs[i].total;
        ro_summary_merge(&merged, &jobs[i].summary); /* train.rs:153-156 */
    }
    free(jobs); free(threads);
    if (summary) *summary = merged;
    return total;
}
