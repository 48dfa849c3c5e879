"""GPU parity for `MlpConfig::hidden_sizes` with two or three entries (mlp.rs:25-34,139-151): Module::forward, the fused
rollout with such a policy / action-value network (generic thread-per-env kernel), advantages through such a critic.
The update passes for these modules are covered in tests/test_gpu_update_shapes.py and tests/test_gpu_dqn.py."""
import numpy as np
import pytest

import oracle as O
import relearn_b200 as R
from relearn_b200 import _lib as L
from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F,hidden,A,act", [(5, [32, 16], 2, "relu"), (14, [24, 40, 12], 10, "tanh"), (7, [256, 31], 4, "sigmoid"),
                                            (36, [8, 8, 8], 32, "identity")])
def test_deep_forward_matches_numpy(ctx, F, hidden, A, act):
    rng = np.random.default_rng(F)
    params = R.init_params(rng, F, hidden, A) * 1.5
    assert params.size == R.num_params(F, hidden, A)
    net = R.Mlp(ctx, F, hidden, A, act)
    assert net.num_params == params.size
    net.set_weights(params)
    np.testing.assert_array_equal(net.get_weights(), params)
    x = rng.normal(size=(300, F)).astype(np.float32)
    got = net.forward(x)
    want = P.mlp_forward_any(params, F, hidden, A, x, act)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)


def test_mlp_create_rejects_what_is_not_built(ctx):
    with pytest.raises(Exception):
        R.Mlp(ctx, 5, [8, 8, 8, 8], 2)          # four hidden layers
    with pytest.raises(Exception):
        R.Mlp(ctx, 5, [300, 16], 2)             # a deep module's layer above 256 units
    with pytest.raises(Exception):
        R.Mlp(ctx, 5, [], 2)                    # no hidden layer


@pytest.mark.parametrize("cfg,hidden,act", [
    (R.CartPoleConfig().wrap(R.VisibleStepLimit(25)), [32, 16], "relu"),
    (R.MemoryGame(4, 3), [24, 16, 8], "tanh"),
    (R.Chain(), [16, 16], "relu"),
], ids=["cartpole", "memory", "chain"])
def test_rollout_with_deep_policy_against_the_oracle(ctx, cfg, hidden, act):
    """PolicyActor::act (policies/actor.rs:42-55) through a deep Mlp inside the fused rollout: every action is the inverse-CDF
    choice of the restated softmax on the recorded observation, and the oracle env replays the trajectory from those actions."""
    rng = np.random.default_rng(11)
    E, T = 96, 40
    env = R.build_env(ctx, cfg, E, seed=3)
    F, A = env.num_features, env.num_actions
    W = 24 * T + 64
    ewords, awords = P.random_words(rng, E, W), P.random_words(rng, E, W)
    env.set_noise_replay(ewords, awords)
    params = R.init_params(rng, F, hidden, A) * 3.0
    net = R.Mlp(ctx, F, hidden, A, act)
    net.set_weights(params)
    traj = R.Trajectory(env, T)
    R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net), R.HistoryDataBound(T, 0), traj)
    host = traj.to_host()
    checked, near = P.check_policy_consistency(host, params, hidden, A, awords, activation=act)
    assert checked > E * (T - 2) and near <= 3
    ref = P.oracle_rollout(cfg, E, T, 0, actor_kind=O.ACTOR_REPLAY, actions=host["action"].copy(), env_words=ewords)
    P.compare_traj(host, ref, what=f"deep policy on {cfg}")


def test_gae_through_a_deep_critic(ctx):
    """Critic::advantages (critics/mod.rs:101-199) with a two-hidden-layer state-value module: values of every stored
    observation (and of the successor on Interrupt) through value_forward_kernel's layer-generic branch."""
    rng = np.random.default_rng(5)
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(12))
    E, T, hidden = 50, 30, [20, 12]
    env = R.build_env(ctx, cfg, E, seed=8)
    traj = R.Trajectory(env, T)
    R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_RANDOM), R.HistoryDataBound(T, 0), traj)
    host = traj.to_host()
    vparams = R.init_params(rng, 5, hidden, 1) * 2.0
    vf = R.Mlp(ctx, 5, hidden, 1, "tanh")
    vf.set_weights(vparams)
    gamma, lam = np.float32(0.97), np.float32(0.9)
    adv, rtg = ctx.alloc(T * E * 4), ctx.alloc(T * E * 4)
    L.check(ctx._lib.rl_gae(traj.handle, vf.handle, gamma, lam, adv.c, rtg.c), ctx.handle)
    a = adv.download((T, E), np.float32)
    assert (host["succ"] == L.RL_INTERRUPT).any()
    for e in range(0, E, 7):
        n = int(host["lane_len"][e])
        v = P.mlp_forward_any(vparams, 5, hidden, 1, host["obs"][:n, e], "tanh")[:, 0]
        vn = P.mlp_forward_any(vparams, 5, hidden, 1, host["next_obs"][:n, e], "tanh")[:, 0]
        want = O.gae_lane(host["reward"][:n, e], v, vn, host["succ"][:n, e], gamma, lam)
        np.testing.assert_allclose(a[:n, e], want, rtol=2e-4, atol=2e-5)
