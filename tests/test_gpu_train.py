"""train_device (relearn_b200/simulation.py) = train_parallel (src/simulation/train.rs:68-186) over device lanes, against the
oracle running the same schedule on the CPU, and the log ids it writes (train.rs:160-184)."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle as O
import relearn_b200 as R
from relearn_b200 import _lib as L
from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = R.Context(0)
    yield c
    c.close()


def test_train_device_runs_the_chain_tabular_q_example_bit_exact(ctx):
    """examples/chain-tabular-q.rs:15-29: Chain::default(), TabularQLearningAgentConfig::default() (epsilon 0.2),
    TrainParallelConfig {num_periods 10, num_threads W, min_worker_steps 10 000}.  One shared table: every worker acts from
    it while a period runs, batch_update folds the workers' buffers in order (tabular.rs:197-207).  The oracle runs the
    same schedule lane by lane on the same Philox streams; table f64 and counts u64 must agree bit for bit, and so must
    the step / episode counts train_device logs."""
    W, periods, steps, eps, seed = 8, 10, 10_000, 0.2, 5
    cfg = R.Chain()
    env = R.build_env(ctx, cfg, W, seed=seed)
    env.set_noise_philox(seed, 0)
    agent = R.TabularQ(ctx, 1, 5, 2, cfg.discount_factor, exploration_rate=eps)
    log = R.HistoryLogger()
    R.train_device(agent, env, R.TrainParallelConfig(num_periods=periods, num_threads=W, min_worker_steps=steps), log)
    q, c = agent.get_table()

    q_ref = np.zeros((5, 2), np.float64)
    c_ref = np.zeros((5, 2), np.uint64)
    olib = O.lib()
    total_steps, total_eps = 0, 0
    for period in range(periods):
        ref = P.oracle_rollout(cfg, W, steps, 0, actor_kind=O.ACTOR_TABULAR, philox_seed=seed, t0=period * (steps + 1),
                               exploration_rate=eps, q_tables=[q_ref.copy() for _ in range(W)], training=True)
        total_steps += int(ref["summary"].step_reward.count)
        total_eps += int(ref["summary"].episode_length.count)
        for e in range(W):  # tabular.rs:197-207: the buffers one after the other
            n = int(ref["lane_len"][e])
            t = O.TabQ()
            t.n_obs, t.n_act, t.discount = 5, 2, cfg.discount_factor
            t.q = q_ref.ctypes.data_as(C.POINTER(C.c_double))
            t.counts = c_ref.ctypes.data_as(C.POINTER(C.c_uint64))
            obs_idx = ref["obs"][:n, e].argmax(axis=1).astype(np.uint32)
            nobs_idx = ref["next_obs"][:n, e].argmax(axis=1).astype(np.uint32)
            olib.ro_tabq_update_buffer(C.byref(t), obs_idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       np.ascontiguousarray(ref["action"][:n, e]).ctypes.data_as(C.POINTER(C.c_uint8)),
                                       np.ascontiguousarray(ref["reward"][:n, e]).ctypes.data_as(C.POINTER(C.c_float)),
                                       np.ascontiguousarray(ref["succ"][:n, e]).ctypes.data_as(C.POINTER(C.c_uint8)),
                                       nobs_idx.ctypes.data_as(C.POINTER(C.c_uint32)), n)
    np.testing.assert_array_equal(c[0], c_ref)
    np.testing.assert_array_equal(q[0], q_ref)
    assert log.counters["sim/step/count"] == total_steps == periods * W * steps
    assert log.counters["sim/ep/count"] == total_eps
    assert log.counters["agent_update/count"] == periods
    assert len(log.scalars["sim/step/fbk/reward/mean"]) == periods
    # the learned greedy policy walks right (the chain's optimum): Q(s, right) > Q(s, left) in every state (chain.rs:38-45)
    assert np.all(q[0][:, 1] > q[0][:, 0]) or np.all(q[0][:, 0] > q[0][:, 1])
    env.close()
    agent.close()


def test_train_device_actor_critic_logs_the_reference_ids(ctx):
    """CartPole + TRPO + critic through train_device: the `sim/...` and `agent_update/...` ids of train.rs:160-184, the
    merged summary equal to the oracle's for the first period (same Philox streams), and a policy that improves."""
    W, periods, seed = 256, 12, 17
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    env = R.build_env(ctx, cfg, W, seed=seed)
    env.set_noise_philox(seed, 0)
    agent = R.ActorCriticConfig(min_batch_size=R.HistoryDataBound(W * 96, 0)).build_agent(env)
    params = R.init_params(np.random.default_rng(2), 5, 128, 2)
    agent.policy.policy_fn.set_weights(params)
    agent.critic.state_value_fn.set_weights(R.init_params(np.random.default_rng(3), 5, 128, 1))
    log = R.HistoryLogger()
    R.train_device(agent, env, R.TrainParallelConfig(num_periods=periods, num_threads=W, min_worker_steps=0), log)
    for key in ("sim/ep/fbk/reward/mean", "sim/ep/fbk/reward/stddev", "sim/ep/length_mean", "sim/ep/length_stddev",
                "sim/step/fbk/reward/mean", "sim/step/fbk/reward/stddev"):
        assert len(log.scalars[key]) == periods, key
    assert log.counters["agent_update/count"] == periods and log.counters["sim/step/count"] == periods * W * 96
    assert log.durations["sim/time"] > 0 and log.durations["agent_update/time"] > 0
    ref = P.oracle_rollout(cfg, W, 96, 0, actor_kind=O.ACTOR_POLICY, params=params, philox_seed=seed, t0=0)
    s = ref["summary"]
    assert log.scalars["sim/step/fbk/reward/mean"][0] == pytest.approx(s.step_reward.mean, rel=1e-12)
    # (an action at a near-tie of the f32 softmax may differ between device and oracle: episode statistics to 1 %)
    assert log.scalars["sim/ep/length_mean"][0] == pytest.approx(s.episode_length.mean, rel=1e-2)
    assert log.scalars["sim/ep/length_mean"][-1] > 1.15 * log.scalars["sim/ep/length_mean"][0]
    env.close()


def test_train_device_dqn_appends_to_the_replay_rings(ctx):
    """DqnAgent through train_device: every period's trajectory is appended to the per-lane replay rings before
    batch_update (dqn.rs:228-230,263-337)."""
    W, periods, seed = 128, 4, 23
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    env = R.build_env(ctx, cfg, W, seed=seed)
    agent = R.DqnConfig(buffer_capacity=2000, minibatch_steps=4096, opt_steps_per_update=5,
                        update_size=R.DataCollectionSchedule(W * 50, W * 50)).build_agent(env)
    agent.action_value_fn.set_weights(R.init_params(np.random.default_rng(4), 5, 128, 2))
    log = R.HistoryLogger()
    R.train_device(agent, env, R.TrainParallelConfig(num_periods=periods, num_threads=W, min_worker_steps=0), log)
    assert log.counters["agent_update/count"] == periods
    # the summary counts the steps the iterator produced; the rings hold what finalize_last_episode kept (a dangling
    # last step per lane and period is popped, buffers/mod.rs:237-261)
    produced = log.counters["sim/step/count"]
    assert produced - W * periods <= agent.global_steps <= produced and produced > 0
    assert len(log.scalars["loss"]) == periods and all(math.isfinite(v) for v in log.scalars["loss"])
    env.close()


def test_shared_table_acts_for_every_lane(ctx):
    """num_replicas = 1: all lanes read table 0 in the rollout (tabular.rs:148-156 gives every worker the same table)."""
    W, T = 12, 50
    cfg = R.Chain()
    env = R.build_env(ctx, cfg, W, seed=3)
    env.set_noise_philox(3, 0)
    agent = R.TabularQ(ctx, 1, 5, 2, cfg.discount_factor, exploration_rate=0.0)
    qt = np.zeros((1, 5, 2)); qt[0, :, 0] = 1.0  # greedy = action 0 everywhere
    agent.set_table(qt, np.zeros((1, 5, 2), np.uint64))
    traj = R.Trajectory(env, T)
    R.rollout(env, agent.actor(training=False), R.HistoryDataBound(T, 0), traj)
    host = traj.to_host()
    assert np.all(host["action"][host["succ"] != L.RL_PAD] == 0)
    traj.close(); env.close(); agent.close()
