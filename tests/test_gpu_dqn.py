"""GPU parity: device ReplayBuffer (replay.rs), DQN minibatch sampling and update (dqn.rs) vs the oracle.

Integer bookkeeping (stored steps, evictions, episode boundaries, total_step_count), the sampled episode
indices and the reward-to-go targets are bit-exact; OneStepTd targets and the Adam-updated parameters are
compared at the stated f32 tolerances.
"""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from oracle import tensor_oracle as TO
import relearn_b200 as R
from relearn_b200 import _lib as L
from tests import parity as P

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CONT, TERM, INTR = L.RL_CONTINUE, L.RL_TERMINATE, L.RL_INTERRUPT


def _raw_codes(ref, e, max_steps):
    """The successor stream of lane e as write_step saw it, before finalize_last_episode."""
    n, n_taken = int(ref["lane_len"][e]), int(ref["n_taken"][e])
    codes = [int(c) for c in ref["succ"][:n, e]]
    if n_taken > n:  # the dangling step was dropped
        if n > 0 and codes[-1] == INTR:
            natural = max_steps and round(float(ref["obs"][n - 1, e, 4]) * max_steps) == 1
            if not natural:
                codes[-1] = CONT  # converted by finalize_last_episode (buffers/mod.rs:237-261)
        codes.append(CONT)
    return codes


class LaneModel:
    """Oracle ReplayBuffer bookkeeping + a host copy of the step records it should hold."""

    def __init__(self, capacity):
        self.rb = O.Replay()
        assert O.lib().ro_replay_init(C.byref(self.rb), capacity) == 0
        self.records = []  # (obs, action, reward, next_obs) of the stored steps, oldest first

    def write(self, codes, host, e, n):
        for c in codes:
            assert O.lib().ro_replay_write_step(C.byref(self.rb), c) == 0
        O.lib().ro_replay_end_experience(C.byref(self.rb))
        for t in range(n):
            self.records.append((host["obs"][t, e].copy(), int(host["action"][t, e]), float(host["reward"][t, e]),
                                 host["next_obs"][t, e].copy()))
        self.records = self.records[len(self.records) - self.rb.n:]

    def stored_codes(self):
        return np.array([self.rb.succ[i] for i in range(self.rb.n)], np.uint8)

    def episode_lens(self):
        ends = [self.rb.episode_ends[i] - self.rb.index_offset for i in range(self.rb.n_eps)]
        return np.diff([0] + ends).astype(np.int64)


def _fill(ctx, cfg, max_steps, E, C_cap, bound, periods, seed, actor_kind=L.RL_ACTOR_RANDOM, check=True):
    env = R.build_env(ctx, cfg, E, seed=seed)
    env.set_noise_philox(seed, 0)
    rb = R.ReplayBuffer(env, C_cap)
    traj = R.Trajectory(env, bound.min_steps + bound.slack_steps)
    models = [LaneModel(C_cap) for _ in range(E)]
    t0 = 0
    for _ in range(periods):
        R.rollout(env, R.ActorSpec(kind=actor_kind), bound, traj)
        rb.write_experience(traj)
        host = traj.to_host()
        ref = P.oracle_rollout(cfg, E, bound.min_steps, bound.slack_steps, actor_kind=O.ACTOR_RANDOM, philox_seed=seed, t0=t0)
        P.compare_traj(host, ref, obs_rtol=1e-6, obs_atol=1e-7, what="dqn fill")
        for e in range(E):
            models[e].write(_raw_codes(ref, e, max_steps), host, e, int(ref["lane_len"][e]))
        t0 += bound.min_steps + bound.slack_steps + 1
    return env, rb, models


@pytest.mark.parametrize("max_steps,C_cap,bound", [(9, 37, (12, 4)), (500, 90, (25, 5)), (6, 13, (12, 1))])
def test_replay_buffer_matches_oracle(ctx, max_steps, C_cap, bound):
    """write_experience for several periods with evictions: every lane holds exactly the steps, episode
    boundaries and counters of the reference ReplayBuffer.

    slack_steps >= 1 in every case (DataCollectionSchedule always uses with_default_slack, >= 5): with zero
    slack a period can end one step after an episode end, and then the reference pops the dangling step without
    decrementing total_step_count (end_experience only does so when finalize_last_episode returns true,
    replay.rs:115-125), after which its own episode boundaries are off by one.  The device buffer keeps the
    documented meaning of total_step_count there instead of copying that."""
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(max_steps))
    E = 70
    env, rb, models = _fill(ctx, cfg, max_steps, E, C_cap, R.HistoryDataBound(*bound), periods=7, seed=21)
    tot = [0, 0, 0]
    evicted = 0
    for e in range(E):
        m = models[e]
        got = rb.read_lane(e)
        np.testing.assert_array_equal(got["succ"], m.stored_codes(), err_msg=f"lane {e}: stored successor codes")
        np.testing.assert_array_equal(got["episode_len"], m.episode_lens(), err_msg=f"lane {e}: episodes")
        assert got["total_step_count"] == m.rb.total_step_count
        assert len(m.records) == m.rb.n
        np.testing.assert_array_equal(got["obs"], np.array([r[0] for r in m.records], np.float32).reshape(-1, 5))
        np.testing.assert_array_equal(got["action"], np.array([r[1] for r in m.records], np.uint8))
        np.testing.assert_array_equal(got["reward"], np.array([r[2] for r in m.records], np.float32))
        intr = got["succ"] == INTR
        if intr.any():
            np.testing.assert_array_equal(got["next_obs"][intr], np.array([r[3] for r in m.records], np.float32)[intr])
        tot[0] += m.rb.n
        tot[1] += m.rb.n_eps
        tot[2] += m.rb.total_step_count
        evicted += m.rb.index_offset > 0
    s = rb.stats()
    assert (s.num_steps, s.num_episodes, s.total_step_count) == tuple(tot)
    assert evicted > E // 2, "the case must exercise evictions"


def test_replay_buffer_full_error(ctx):
    """An episode longer than the capacity: WriteExperienceError::Full (replay.rs:91-95)."""
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    env = R.build_env(ctx, cfg, 64, seed=3)
    rb = R.ReplayBuffer(env, 4)
    traj = R.Trajectory(env, 40)
    R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_RANDOM), R.HistoryDataBound(40, 0), traj)
    with pytest.raises(L.RelearnB200Error) as ei:
        rb.write_experience(traj)
    assert ei.value.status == L.RL_ERR_BUFFER_FULL


def _uniform_int(seed, j, draw_index, n):
    """Uniform::new(0, n).sample on the sampler's Philox stream (oracle restatement of rand 0.8.5)."""
    reject = ((1 << 64) - n) % n
    zone = (1 << 64) - 1 - reject
    for d in range(64):
        v = O.lib().ro_philox_slot(seed, j, draw_index, L.RL_STREAM_SAMPLER, d)
        lo = (v * n) & ((1 << 64) - 1)
        if lo <= zone:
            return (v * n) >> 64
    raise AssertionError("rejection loop did not terminate")


def _oracle_minibatch(lanes, seed, draw_index, minibatch_steps):
    """sample_minibatch (dqn.rs:280-297): round-robin over buffers, uniform episode, take_while on the total."""
    E = len(lanes)
    episodes, total, j = [], 0, 0
    while total < minibatch_steps:
        lane = lanes[j % E]
        lens = lane["episode_len"]
        k = _uniform_int(seed, j, draw_index, len(lens))
        a = int(lens[:k].sum())
        b = a + int(lens[k])
        episodes.append({"obs": lane["obs"][a:b], "action": lane["action"][a:b], "reward": lane["reward"][a:b],
                         "last_succ": int(lane["succ"][b - 1]), "next_obs": lane["next_obs"][b - 1]})
        total += b - a
        j += 1
    return episodes


O.lib().ro_philox_slot.restype = C.c_uint64
O.lib().ro_philox_slot.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32]


@pytest.fixture(scope="module")
def filled(ctx):
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(15))
    E = 48
    env, rb, models = _fill(ctx, cfg, 15, E, 120, R.HistoryDataBound(30, 6), periods=6, seed=33)
    lanes = [rb.read_lane(e) for e in range(E)]
    return env, rb, lanes


@pytest.mark.parametrize("draw_index,minibatch", [(0, 200), (5, 1), (9, 1500)])
def test_sample_reward_to_go_bit_exact(ctx, filled, draw_index, minibatch):
    env, rb, lanes = filled
    seed = 77
    cfg = L.DqnCfg(minibatch, 1, 0, np.float32(0.99), seed)
    got = rb.sample(cfg, None, draw_index)
    eps = _oracle_minibatch(lanes, seed, draw_index, minibatch)
    obs = np.concatenate([e["obs"] for e in eps])
    act = np.concatenate([e["action"] for e in eps])
    tgt = np.concatenate([O.discounted_cumsum_lane(e["reward"], np.array([CONT] * (len(e["reward"]) - 1) + [TERM], np.uint8),
                                                   np.float32(0.99)) for e in eps])
    assert got["num_episodes"] == len(eps) and got["num_steps"] == len(act)
    np.testing.assert_array_equal(got["obs"], obs)
    np.testing.assert_array_equal(got["action"], act)
    np.testing.assert_array_equal(got["target"], tgt)
    assert (got["succ"][:len(act)] == 0).all() and (got["succ"][len(act):] == L.RL_PAD).all()


def test_sample_one_step_td_targets(ctx, filled):
    """r + gamma * max_a Q(next): 0 after Terminate, the stored successor observation after Interrupt.
    Tolerance: f32 forward with a different summation order than torch, 1e-5 relative."""
    env, rb, lanes = filled
    rng = np.random.default_rng(4)
    params = R.init_params(rng, 5, 128, 2)
    q = R.Mlp(ctx, 5, [128], 2)
    q.set_weights(params)
    seed, draw_index, minibatch = 5, 2, 400
    cfg = L.DqnCfg(minibatch, 1, 1, np.float32(0.99), seed)
    got = rb.sample(cfg, q, draw_index)
    eps = _oracle_minibatch(lanes, seed, draw_index, minibatch)
    tgt = []
    n_intr = 0
    for e in eps:
        v = O.mlp_forward(params, 5, 128, 2, e["obs"]).max(axis=1)
        nxt = np.zeros(len(v), np.float32)
        nxt[:-1] = v[1:]
        if e["last_succ"] == INTR:
            nxt[-1] = O.mlp_forward(params, 5, 128, 2, e["next_obs"][None, :]).max()
            n_intr += 1
        tgt.append(e["reward"] + np.float32(0.99) * nxt)
    tgt = np.concatenate(tgt)
    assert n_intr > 0, "the case must include interrupted episodes"
    np.testing.assert_allclose(got["target"], tgt, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("td", [False, True])
def test_dqn_update_matches_oracle(ctx, filled, pass_kernel, td):
    """opt_steps x {sample, mse(Q(obs)[a], target), backward, Adam} vs the torch restatement on the same sampled
    episodes: parameter delta within max(2e-4, 4x the torch-f32 run's own distance from the f64 run)."""
    env, rb, lanes = filled
    rng = np.random.default_rng(8)
    params = R.init_params(rng, 5, 128, 2)
    steps, minibatch, seed = 6, 600, 123
    agent = R.DqnConfig(minibatch_steps=minibatch, opt_steps_per_update=steps, target_one_step_td=td, sample_seed=seed,
                        buffer_capacity=120).build_agent(env)
    agent.action_value_fn.set_weights(params)
    base = _DRAWS.get(id(rb), 0)  # the buffer's sampler counter advances by one per optimizer step
    _DRAWS[id(rb)] = base + steps
    log = {}
    stats = agent.batch_update(rb, log)
    new = agent.action_value_fn.get_weights()
    mbs = [_oracle_minibatch(lanes, seed, base + s, minibatch) for s in range(steps)]
    g = float(agent.discount_factor)
    new64, losses64 = TO.dqn_update(params, 5, 128, 2, mbs, g, one_step_td=td, dtype=torch.float64)
    new32, losses32 = TO.dqn_update(params, 5, 128, 2, mbs, g, one_step_td=td, dtype=torch.float32)
    d, d64, d32 = new - params, new64 - params.astype(np.float64), new32 - params
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
    print(f"dqn {pass_kernel} td={td} delta rel err vs f64: kernel {rel(d, d64):.2e}, torch-f32 {rel(d32, d64):.2e}; "
          f"loss first/last {stats.loss_first:.6f}/{stats.loss_last:.6f} vs {losses64[0]:.6f}/{losses64[-1]:.6f}")
    assert stats.opt_steps == steps and stats.num_steps == sum(len(e["action"]) for e in mbs[-1])
    np.testing.assert_allclose(stats.loss_first, losses64[0], rtol=1e-5)
    np.testing.assert_allclose(stats.loss_last, losses64[-1], rtol=1e-4)
    assert rel(d, d64) <= max(2e-4, 4 * rel(d32, d64) + 1e-5)
    assert agent.global_steps == rb.total_step_count()


_DRAWS = {}


@pytest.mark.parametrize("hidden", [64, [32, 24]], ids=["one-hidden-layer", "two-hidden-layers"])
@pytest.mark.parametrize("td", [False, True])
def test_dqn_update_other_network_shape(ctx, td, hidden):
    """DqnAgent::batch_update with a module the default kernels do not serve (MemoryGame: 7 features, 4 actions; 64 tanh
    units -> mlp_pass_any_kernel's Q-loss pass): same sampled episodes, same bound as the default-network test."""
    cfg = R.MemoryGame(4, 3)
    E, act = 24, "tanh"
    env, rb, models = _fill(ctx, cfg, 0, E, 80, R.HistoryDataBound(24, 4), periods=3, seed=44)
    lanes = [rb.read_lane(e) for e in range(E)]
    F, A = env.num_features, env.num_actions
    params = R.init_params(np.random.default_rng(9), F, hidden, A)
    steps, minibatch, seed = 5, 300, 77
    hs = [hidden] if isinstance(hidden, int) else hidden
    agent = R.DqnConfig(action_value_fn_config=R.MlpConfig(hidden_sizes=hs, activation=act), minibatch_steps=minibatch,
                        opt_steps_per_update=steps, target_one_step_td=td, sample_seed=seed, buffer_capacity=80).build_agent(env)
    agent.action_value_fn.set_weights(params)
    stats = agent.batch_update(rb, {})
    new = agent.action_value_fn.get_weights()
    mbs = [_oracle_minibatch(lanes, seed, s, minibatch) for s in range(steps)]
    g = float(agent.discount_factor)
    with TO.mlp_activation(act):
        new64, losses64 = TO.dqn_update(params, F, hidden, A, mbs, g, one_step_td=td, dtype=torch.float64)
        new32, _ = TO.dqn_update(params, F, hidden, A, mbs, g, one_step_td=td, dtype=torch.float32)
    d, d64, d32 = new - params, new64 - params.astype(np.float64), new32 - params
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
    print(f"dqn {F}->{hidden}->{A} {act} td={td}: delta rel err vs f64 kernel {rel(d, d64):.2e}, torch-f32 {rel(d32, d64):.2e}")
    assert stats.opt_steps == steps and stats.num_steps == sum(len(e["action"]) for e in mbs[-1])
    np.testing.assert_allclose(stats.loss_first, losses64[0], rtol=1e-5)
    np.testing.assert_allclose(stats.loss_last, losses64[-1], rtol=1e-4)
    assert rel(d, d64) <= max(2e-4, 4 * rel(d32, d64) + 1e-5)


def test_dqn_learns_cartpole(ctx):
    """Behavioural check in the spirit of agents/testing.rs / dqn.rs:391-414: a few DQN periods with the
    default Monte-Carlo targets raise the greedy policy's mean episode length."""
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    E = 1024
    env = R.build_env(ctx, cfg, E, seed=2)
    agent = R.DqnConfig(minibatch_steps=20_000, opt_steps_per_update=30, buffer_capacity=600,
                        exploration_rate=R.ExplorationRateSchedule(1.0, 0.1, 200_000),
                        update_size=R.DataCollectionSchedule(64 * E, 32 * E)).build_agent(env)
    agent.action_value_fn.set_weights(R.init_params(np.random.default_rng(1), 5, 128, 2))
    rb = agent.buffer()
    lengths = []
    for period in range(12):
        bound = agent.min_update_size().divide(E)
        traj = R.Trajectory(env, bound.min_steps + bound.slack_steps)
        R.rollout(env, agent.actor(training=True), bound, traj)
        rb.write_experience(traj)
        agent.batch_update(rb, {})
        ev = R.Trajectory(env, 200)
        summ = R.rollout(env, agent.actor(training=False), R.HistoryDataBound(200, 0), ev)
        lengths.append(summ.episode_length.mean)
        traj.close(); ev.close()
    print("greedy mean episode length per period:", [round(x, 1) for x in lengths])
    assert max(lengths[-4:]) > 1.5 * lengths[0]
