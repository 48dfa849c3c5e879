"""GPU parity of the UCB1 agent (src/agents/bandits/ucb.rs) against oracle.Ucb1Oracle: the actor inside the fused rollout,
the ordered fold of batch_update, the shared-table arrangement of train_parallel, and the behavioural check of the
reference's own test (`learns_determinstic_bandit`, ucb.rs:249-256) in spirit."""
import numpy as np
import pytest

import oracle as O
import relearn_b200 as R
from relearn_b200 import _lib as L
from tests import parity as P

pytestmark = pytest.mark.gpu


def _obs_index(obs_row):
    return int(np.flatnonzero(obs_row)[-1]) if np.any(obs_row) else 0


def _random_tables(rng, R_, S, A):
    count = rng.integers(2, 60, size=(R_, S, A)).astype(np.uint64)
    mean = rng.uniform(0.0, 1.0, size=(R_, S, A))
    visits = count.sum(axis=2).astype(np.uint64)
    return mean, count, visits


@pytest.mark.parametrize("cfg", [R.Chain(), R.MemoryGame(3, 2)], ids=["chain", "memory"])
@pytest.mark.parametrize("shared", [False, True], ids=["table-per-lane", "shared-table"])
@pytest.mark.parametrize("training", [True, False], ids=["training", "evaluation"])
def test_ucb1_actor_and_update_match_the_oracle(ctx, cfg, shared, training):
    rng = np.random.default_rng(3)
    E, T = 48, 60
    env = R.build_env(ctx, cfg, E, seed=9)
    st = env.structure
    S, A = st.num_observations, st.num_actions
    reps = 1 if shared else E
    agent = R.UCB1AgentConfig(0.3).build_agent(env, num_replicas=reps)
    m0, c0, v0 = agent.get_tables()
    assert (m0 == 0.5).all() and (c0 == 2).all() and (v0 == 2 * A).all()  # ucb.rs:125-128
    mean, count, visits = _random_tables(rng, reps, S, A)
    mean[:, :, -1] = mean[:, :, 0]; count[:, :, -1] = count[:, :, 0]      # exact ties between the first and the last action
    agent.set_tables(mean, count, visits)
    words = P.random_words(rng, E, 8 * T)
    env.set_noise_replay(words, None)
    traj = R.Trajectory(env, T)
    R.rollout(env, agent.actor(training=training), R.HistoryDataBound(T, 0), traj)
    host = traj.to_host()
    # (1) every action is the oracle actor's choice for the recorded observation (the tables are frozen during a rollout)
    oracles = []
    for r in range(reps):
        o = O.Ucb1Oracle(S, A, (st.reward_lo, st.reward_hi), 0.3)
        o.mean, o.count, o.visits = mean[r].copy(), count[r].copy(), visits[r].copy()
        oracles.append(o)
    near = 0
    for e in range(E):
        o = oracles[0 if shared else e]
        for t in range(int(host["lane_len"][e])):
            s = _obs_index(host["obs"][t, e])
            want = o.act(s, training)
            if want != host["action"][t, e]:
                vals = sorted(o.ucb(s))
                assert training and abs(vals[-1] - vals[-2]) <= 1e-13 * abs(vals[-1]), (e, t, want, host["action"][t, e])
                near += 1
    assert near <= 2
    # (2) the env side replays from those actions
    ref = P.oracle_rollout(cfg, E, T, 0, actor_kind=O.ACTOR_REPLAY, actions=host["action"].copy(), env_words=words)
    P.compare_traj(host, ref, what="ucb1 rollout")
    # (3) batch_update = step_update over every stored step, lane after lane: bit-exact tables
    agent.update(traj)
    gm, gc, gv = agent.get_tables()
    for e in range(E):
        o = oracles[0 if shared else e]
        for t in range(int(host["lane_len"][e])):
            o.step_update(_obs_index(host["obs"][t, e]), int(host["action"][t, e]), float(host["reward"][t, e]))
    for r in range(reps):
        np.testing.assert_array_equal(gc[r], oracles[r].count)
        np.testing.assert_array_equal(gv[r], oracles[r].visits)
        np.testing.assert_array_equal(gm[r], oracles[r].mean)


def test_ucb1_concentrates_on_the_best_immediate_reward(ctx):
    """ucb.rs:249-256 in spirit on an env of this path.  UCB1 is a bandit rule: per state it maximises the IMMEDIATE reward
    (on Chain that is the small reward of going back, chain.rs:75-105), so after training the most selected action of the
    most visited state is the one with the highest empirical mean, and the shared tables have seen every step of every lane."""
    E, T = 64, 50
    env = R.build_env(ctx, R.Chain(), E, seed=4)
    agent = R.UCB1AgentConfig().build_agent(env)
    traj = R.Trajectory(env, T)
    steps = 0
    for _ in range(30):
        R.rollout(env, agent.actor(training=True), R.HistoryDataBound(T, 0), traj)
        steps += int(traj.to_host()["lane_len"].sum())
        agent.update(traj)
    mean, count, visits = agent.get_tables()
    assert int(visits.sum()) == steps + visits.size * 2 * count.shape[2]
    s = int(np.argmax(visits[0]))
    print(f"state {s}: counts {count[0, s]}, means {np.round(mean[0, s], 4)}")
    assert np.argmax(count[0, s]) == np.argmax(mean[0, s])
    assert count[0, s].max() > 10 * count[0, s].min()
    summ = R.rollout(env, agent.actor(training=False), R.HistoryDataBound(T, 0), traj)
    acts = traj.to_host()["action"]
    assert summ.step_reward.count == E * T and (acts[0] == np.argmax(count[0, 0])).all()  # every lane starts in state 0


def test_ucb1_rejects_unbounded_or_mismatched(ctx):
    with pytest.raises(Exception):
        R.UCB1Agent(ctx, 1, 5, 2, (0.0, float("inf")))
    env = R.build_env(ctx, R.CartPoleConfig(), 8, seed=1)   # not a finite observation space
    with pytest.raises(Exception):
        R.UCB1AgentConfig().build_agent(env)
