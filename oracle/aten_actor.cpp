// aten_actor.cpp -- the reference's CPU rollout with PolicyActor::act as batch-1 ATen calls (TEST / BENCH INFRASTRUCTURE,
// like the rest of oracle/: only tests/, smoke() and bench.py's CPU legs may load it).
//
// What the reference does per env-step (src/torch/agents/policies/actor.rs:42-55 -> modules/ff/mlp.rs forward ->
// distributions/categorical.rs:29-33,52-54) through `tch`, which forwards one-to-one to ATen:
//     features [1, F] -> linear -> relu -> linear -> log_softmax(-1) -> exp -> multinomial(1, true) -> item
// under a no_grad guard, one env per OS thread (train.rs:124-158).  This harness runs exactly those ATen calls
// against the torch wheel's libtorch_cpu.so, around the oracle's own env loop (ro_rollout_lane, relearn_oracle.c), one
// lane at a time per thread (batch-1 ops never leave the calling thread).  It exists to TIME the path BASELINE.md section 2 names; its
// actions come from torch's own generator (as the reference's do), so they are not comparable with the Philox-noise
// oracle lane by lane -- only the step counts and the summary statistics are.
#include <ATen/ATen.h>
#include <ATen/Parallel.h>
#include <torch/csrc/autograd/grad_mode.h>

#include <cstring>
#include <thread>
#include <vector>

extern "C" {
#include "relearn_oracle.h"
}

namespace {

struct AtenPolicy {
    at::Tensor w1, b1, w2, b2;
};

uint64_t aten_act(void *ud, const float *obs, int F) {
    const AtenPolicy &p = *static_cast<const AtenPolicy *>(ud);
    at::Tensor x = at::from_blob(const_cast<float *>(obs), {1, F}, at::kFloat);
    at::Tensor z = at::linear(at::relu(at::linear(x, p.w1, p.b1)), p.w2, p.b2);
    at::Tensor a = at::log_softmax(z, -1).exp().multinomial(1, true).squeeze(-1);
    return (uint64_t)a.item<int64_t>();
}

struct Job {
    const ro_env_cfg *cfg;
    const float *params;
    int in, hidden, out;
    uint64_t lane_begin, lane_end, lane0;
    size_t min_steps, slack;
    uint64_t seed;
    uint32_t t0;
    uint64_t total;
    ro_summary summary;
};

void worker(Job *job) {
    // (no at::set_num_threads here: it would change the process-wide intra-op setting the update baseline relies on; ops
    //  on [1, 128] tensors are far below ATen's parallel grain size and run on the calling thread)
    torch::autograd::AutoGradMode no_grad(false);
    AtenPolicy pol;
    // flat layout of ro_mlp (relearn_oracle.h): w1 [hidden][in], b1 [hidden], w2 [out][hidden], b2 [out]
    const float *q = job->params;
    pol.w1 = at::from_blob(const_cast<float *>(q), {job->hidden, job->in}, at::kFloat).clone();
    q += (size_t)job->hidden * job->in;
    pol.b1 = at::from_blob(const_cast<float *>(q), {job->hidden}, at::kFloat).clone();
    q += job->hidden;
    pol.w2 = at::from_blob(const_cast<float *>(q), {job->out, job->hidden}, at::kFloat).clone();
    q += (size_t)job->out * job->hidden;
    pol.b2 = at::from_blob(const_cast<float *>(q), {job->out}, at::kFloat).clone();
    ro_env env;
    ro_env_init(&env, job->cfg);
    const int F = ro_env_num_features(&env);
    const size_t cap = job->min_steps + job->slack;
    std::vector<float> obs(cap * F), next_obs(cap * F), reward(cap);
    std::vector<uint8_t> action(cap), succ(cap);
    ro_lane_out out;
    out.cap = cap; out.obs = obs.data(); out.next_obs = next_obs.data(); out.action = action.data();
    out.reward = reward.data(); out.succ = succ.data();
    std::memset(&job->summary, 0, sizeof(job->summary));
    job->total = 0;
    for (uint64_t l = job->lane_begin; l < job->lane_end; ++l) {
        ro_actor actor;
        std::memset(&actor, 0, sizeof(actor));
        actor.kind = RO_ACTOR_CALLBACK;
        actor.act_fn = aten_act;
        actor.act_ud = &pol;
        ro_rng rng;
        ro_rng_philox(&rng, job->seed, job->lane0 + l, job->t0);
        ro_summary lane_summary;
        std::memset(&lane_summary, 0, sizeof(lane_summary));
        job->total += ro_rollout_lane(&env, &actor, job->min_steps, job->slack, &rng, &rng, job->t0, &out, &lane_summary);
        ro_summary_merge(&job->summary, &lane_summary);
    }
}

}  // namespace

extern "C" uint64_t ro_rollout_lanes_aten(const ro_env_cfg *cfg, const float *params, int in, int hidden, int out,
                                          uint64_t n_lanes, uint64_t lane0, size_t min_steps, size_t slack, uint64_t seed,
                                          uint32_t t0, int n_threads, ro_summary *summary) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_lanes && n_lanes > 0) n_threads = (int)n_lanes;
    std::vector<Job> jobs((size_t)n_threads);
    std::vector<std::thread> threads;
    for (int i = 0; i < n_threads; ++i) {
        Job &j = jobs[(size_t)i];
        j.cfg = cfg; j.params = params; j.in = in; j.hidden = hidden; j.out = out; j.lane0 = lane0;
        j.lane_begin = n_lanes * (uint64_t)i / (uint64_t)n_threads;
        j.lane_end = n_lanes * (uint64_t)(i + 1) / (uint64_t)n_threads;
        j.min_steps = min_steps; j.slack = slack; j.seed = seed; j.t0 = t0;
        threads.emplace_back(worker, &j);
    }
    uint64_t total = 0;
    ro_summary merged;
    std::memset(&merged, 0, sizeof(merged));
    for (int i = 0; i < n_threads; ++i) {
        threads[(size_t)i].join();
        total += jobs[(size_t)i].total;
        ro_summary_merge(&merged, &jobs[(size_t)i].summary);
    }
    if (summary) *summary = merged;
    return total;
}
