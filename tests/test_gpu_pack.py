"""GPU parity of `rl_pack_history` (LazyHistoryFeatures, src/torch/agents/features.rs:70-215, in the packed order of
src/torch/packed.rs:346-420) against oracle.pack_episodes: the reference's own known-answer test (features.rs:293-406)
loaded into a device trajectory, and trajectories produced by the fused rollout (ragged lanes, slack, Interrupts)."""
import numpy as np
import pytest

import oracle as O
import relearn_b200 as R
from relearn_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _episodes_of(host):
    """Episodes in buffer order (lane after lane), each a list of (t, lane)."""
    T, E = host["succ"].shape
    eps = []
    for e in range(E):
        cur = []
        for t in range(int(host["lane_len"][e])):
            cur.append((t, e))
            if host["succ"][t, e] != L.RL_CONTINUE:
                eps.append(cur)
                cur = []
        assert not cur  # the stored history ends with a finished (or interrupted) episode: buffers/mod.rs:237-261
    return eps


def _check(host, packed):
    eps = _episodes_of(host)
    ref = O.pack_episodes(eps)
    N, M = len(ref["steps"]), len(eps)
    assert (packed["num_steps"], packed["num_episodes"]) == (N, M)
    assert packed["max_len"] == (len(eps[ref["order"][0]]) if eps else 0)
    np.testing.assert_array_equal(packed["batch_sizes"], np.asarray(ref["batch_sizes"], np.int64))
    idx = np.asarray(ref["steps"], np.int64).reshape(-1, 2)
    np.testing.assert_array_equal(packed["obs"], host["obs"][idx[:, 0], idx[:, 1]])
    np.testing.assert_array_equal(packed["action"], host["action"][idx[:, 0], idx[:, 1]].astype(np.int64))
    np.testing.assert_array_equal(packed["reward"], host["reward"][idx[:, 0], idx[:, 1]])
    # extended observations: every episode with one more entry (features.rs:217-262)
    ext_eps = [ep + [("end", ep[-1])] for ep in eps]
    ext = O.pack_episodes(ext_eps)
    np.testing.assert_array_equal(packed["ext_batch_sizes"], np.asarray(ext["batch_sizes"], np.int64))
    assert len(ext["steps"]) == N + M
    for i, item in enumerate(ext["steps"]):
        if item[0] == "end":
            t, e = item[1]
            if host["succ"][t, e] == L.RL_INTERRUPT:
                assert not packed["ext_invalid"][i]
                np.testing.assert_array_equal(packed["ext_obs"][i], host["next_obs"][t, e])
            else:
                assert packed["ext_invalid"][i] and not packed["ext_obs"][i].any()
        else:
            assert not packed["ext_invalid"][i]
            np.testing.assert_array_equal(packed["ext_obs"][i], host["obs"][item[0], item[1]])


def test_pack_history_reference_kat(ctx):
    """features.rs:293-406: episodes of lengths 4, 6, 3, 1 -> batch sizes [4, 3, 3, 2, 1, 1] and the interleaved order."""
    episodes = [
        [(True, 0, 1.0), (True, 1, 1.0), (True, 2, 1.0), (True, 3, 1.0)],
        [(False, 10, -1.0), (False, 11, -1.0), (False, 12, 0.0), (False, 13, 0.0), (False, 14, 1.0), (False, 15, 1.0)],
        [(False, 20, 2.0), (True, 21, 2.0), (False, 22, 2.0)],
        [(True, 30, 3.0)],
    ]
    T, E, F = 6, 4, 5
    obs = np.zeros((T, E, F), np.float32); nobs = np.zeros((T, E, F), np.float32)
    action = np.zeros((T, E), np.uint8); reward = np.zeros((T, E), np.float32)
    succ = np.full((T, E), L.RL_PAD, np.uint8)
    for e, ep in enumerate(episodes):
        for t, (o, a, r) in enumerate(ep):
            obs[t, e, 0], action[t, e], reward[t, e] = float(o), a, r
            succ[t, e] = L.RL_TERMINATE if t == len(ep) - 1 else L.RL_CONTINUE
    env = R.build_env(ctx, R.CartPoleConfig().wrap(R.VisibleStepLimit(500)), E, seed=0)
    traj = R.Trajectory(env, T)
    traj.load(obs, action, reward, succ, nobs)
    p = R.pack_history(traj)
    assert p["batch_sizes"].tolist() == [4, 3, 3, 2, 1, 1]
    assert p["action"].tolist() == [10, 0, 20, 30, 11, 1, 21, 12, 2, 22, 13, 3, 14, 15]
    assert p["obs"][:, 0].tolist() == [0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 0, 1, 0, 0]
    assert p["reward"].tolist() == [-1, 1, 2, 3, -1, 1, 2, 0, 1, 2, 0, 1, 1, 1]
    assert p["ext_batch_sizes"].tolist() == [4, 4, 3, 3, 2, 1, 1] and p["ext_invalid"].sum() == 4
    _check(traj.to_host(), p)


@pytest.mark.parametrize("cfg,E,T,slack", [
    (R.CartPoleConfig().wrap(R.VisibleStepLimit(11)), 70, 48, 5),
    (R.Chain(), 33, 40, 0),
    (R.MetaEnv(R.UniformBernoulliBandits(3), 4), 20, 30, 7),
], ids=["cartpole-limit11", "chain", "bandit-meta"])
def test_pack_history_of_rollouts(ctx, cfg, E, T, slack):
    env = R.build_env(ctx, cfg, E, seed=6)
    traj = R.Trajectory(env, T + slack)
    R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_RANDOM), R.HistoryDataBound(T, slack), traj)
    host = traj.to_host()
    assert (host["succ"] == L.RL_INTERRUPT).any() or not isinstance(cfg, R.CartPoleConfig)
    _check(host, R.pack_history(traj))


def test_pack_history_of_an_empty_trajectory(ctx):
    env = R.build_env(ctx, R.Chain(), 4, seed=1)
    traj = R.Trajectory(env, 8)
    p = R.pack_history(traj)
    assert (p["num_steps"], p["num_episodes"], p["max_len"]) == (0, 0, 0) and p["obs"].shape == (0, env.num_features)
