"""Committed golden trajectories (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
the CPU suite checks that the oracle still reproduces them, the GPU suite checks the CUDA rollouts against them."""
import glob
import os

import numpy as np
import pytest

import oracle as O
import relearn_b200 as R
from relearn_b200 import _lib as L
from tests import parity as P
from tests.golden.make_golden import CASES

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return dict(np.load(os.path.join(HERE, name + ".npz")))


def test_every_case_has_a_fixture():
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "*.npz"))}
    assert have == set(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    cfg, E, T, slack = CASES[name]
    g = _load(name)
    ref = P.oracle_rollout(cfg, E, T, slack, actor_kind=O.ACTOR_REPLAY, actions=g["actions"], env_words=g["words"])
    for k in ("obs", "next_obs", "action", "reward", "succ", "lane_len"):
        np.testing.assert_array_equal(ref[k], g[k], err_msg=f"{name}: {k}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_rollout_reproduces_golden(ctx, name):
    cfg, E, T, slack = CASES[name]
    g = _load(name)
    env = R.build_env(ctx, cfg, E, seed=1)
    env.set_noise_replay(g["words"], None)
    traj = R.Trajectory(env, T + slack)
    R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_REPLAY_ACTIONS, actions=g["actions"]), R.HistoryDataBound(T, slack), traj)
    host = traj.to_host()
    f64 = name.startswith("cartpole")  # sin/cos: <= 1 ulp in f64, invisible or 1 ulp in the f32 observation
    P.compare_traj(host, g, obs_rtol=1e-6 if f64 else 0.0, obs_atol=1e-7 if f64 else 0.0, what=f"golden {name}")
