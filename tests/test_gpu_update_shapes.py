"""GPU parity of the update passes for modules OTHER than the reference's 5 -> 128 -> {1, 2} ReLU defaults.

`mlp_pass_any_kernel` (update.cu) serves every one-hidden-layer MlpConfig (mlp.rs:21-61): other observation sizes
(MemoryGame, the bandit meta-env), other action counts, other hidden sizes and the Tanh / Sigmoid activations
(ff/activation.rs:11) -- and, in its layer-generic form, `hidden_sizes` with two or three entries (mlp.rs:25-34,139-151).  Same oracle and same bounds as tests/test_gpu_update.py: loss / gradient / Fisher-vector product
against the f64 autograd run at rtol 1e-5; whole updates against the f64 run within max(2e-4, 4 x the torch-f32 run's own
distance from it).
"""
import numpy as np
import pytest

import oracle as O
from oracle import tensor_oracle as TO
import relearn_b200 as R
from relearn_b200 import _lib as L

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

SHAPES = [
    pytest.param(R.MemoryGame(4, 3), 64, "relu", id="memory-7-64-4-relu"),
    pytest.param(R.MetaEnv(R.UniformBernoulliBandits(10), 7), 128, "tanh", id="bandit-14-128-10-tanh"),
    pytest.param(R.Chain(), 32, "sigmoid", id="chain-5-32-2-sigmoid"),
    pytest.param(R.CartPoleConfig().wrap(R.VisibleStepLimit(40)), 256, "relu", id="cartpole-5-256-2-relu"),
    pytest.param(R.CartPoleConfig().wrap(R.LatentStepLimit(30)), 100, "relu", id="cartpole-4-100-2-relu"),
    # MlpConfig::hidden_sizes with two and three entries (even / odd layer widths, widths below and above a warp)
    pytest.param(R.CartPoleConfig().wrap(R.VisibleStepLimit(40)), [32, 16], "tanh", id="cartpole-5-32-16-2-tanh"),
    pytest.param(R.MemoryGame(4, 3), [64, 33], "relu", id="memory-7-64-33-4-relu"),
    pytest.param(R.MetaEnv(R.UniformBernoulliBandits(10), 7), [24, 40, 12], "sigmoid", id="bandit-14-24-40-12-10-sigmoid"),
]
DEEP = SHAPES[5:]


def _hs(hidden):
    return [hidden] if isinstance(hidden, int) else list(hidden)


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) /
                 max(np.linalg.norm(np.asarray(b, np.float64)), 1e-300))


def _collect(ctx, cfg, hidden, activation, E, T, seed, scale=1.5):
    rng = np.random.default_rng(seed)
    env = R.build_env(ctx, cfg, E, seed=seed)
    F, A = env.num_features, env.num_actions
    params = (R.init_params(rng, F, hidden, A) * scale).astype(np.float32)
    net = R.Mlp(ctx, F, _hs(hidden), A, activation)
    net.set_weights(params)
    traj = R.Trajectory(env, T)
    R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net), R.HistoryDataBound(T, 0), traj)
    host = traj.to_host()
    valid = host["succ"] != L.RL_PAD
    adv = rng.normal(size=(T, E)).astype(np.float32)
    return env, traj, net, params, host, valid, adv, F, A


@pytest.mark.parametrize("cfg,hidden,activation", SHAPES)
def test_policy_probe_any_shape(ctx, cfg, hidden, activation):
    """STATS / GRAD / FVP passes: entropy, loss, KL, gradient and Fisher-vector product vs f64 autograd."""
    E, T = 70, 45
    env, traj, net, params, host, valid, adv, F, A = _collect(ctx, cfg, hidden, activation, E, T, seed=5)
    policy = R.Trpo(net, R.TrpoConfig())
    vec = np.random.default_rng(6).normal(size=net.num_params).astype(np.float32)
    got = policy.probe(traj, ctx.to_device(adv), vec)
    obs, act, a = host["obs"][valid], host["action"][valid], adv[valid]
    with TO.mlp_activation(activation):
        loss64, kl64, ent64, g64, hv64 = TO.policy_loss_kl_grad_fvp(params, F, hidden, A, obs, act, a, vec, 1e-5, torch.float64)
        _, _, _, g32, hv32 = TO.policy_loss_kl_grad_fvp(params, F, hidden, A, obs, act, a, vec, 1e-5, torch.float32)
    print(f"N={valid.sum()} {F}->{hidden}->{A} {activation}: grad rel err kernel {_rel(got['grad'], g64):.2e} torch-f32 "
          f"{_rel(g32, g64):.2e}; fvp kernel {_rel(got['fvp'], hv64):.2e} torch-f32 {_rel(hv32, hv64):.2e}")
    assert abs(got["loss"] - loss64) <= 1e-6 * max(1.0, abs(loss64))
    assert abs(got["kl"]) <= 1e-7
    assert abs(got["entropy"] - ent64) <= 1e-6 * max(1.0, abs(ent64))
    assert _rel(got["grad"], g64) <= 1e-5
    assert _rel(got["fvp"], hv64) <= 1e-5


@pytest.mark.parametrize("cfg,hidden,activation", SHAPES[:3] + DEEP)
def test_trpo_update_any_shape(ctx, cfg, hidden, activation):
    """The whole trust-region step (CG, step size, line search) on a well-conditioned problem (hpv_reg_coeff 0.1)."""
    E, T, reg = 90, 60, 0.1
    env, traj, net, params, host, valid, adv, F, A = _collect(ctx, cfg, hidden, activation, E, T, seed=7)
    policy = R.Trpo(net, R.TrpoConfig(optimizer_config=R.ConjugateGradientOptimizerConfig(hpv_reg_coeff=reg)))
    log = {}
    status = policy.update(traj, ctx.to_device(adv), log)
    new = net.get_weights()
    obs, act, a = host["obs"][valid], host["action"][valid], adv[valid]
    ocfg = TO.CgConfig(hpv_reg_coeff=reg)
    with TO.mlp_activation(activation):
        new64, log64 = TO.trpo_update(params, F, hidden, A, obs, act, a, cfg=ocfg, dtype=torch.float64)
        new32, log32 = TO.trpo_update(params, F, hidden, A, obs, act, a, cfg=ocfg, dtype=torch.float32)
    d, d64, d32 = new - params, new64 - params.astype(np.float64), new32 - params
    print(f"{F}->{hidden}->{A} {activation}: status {status}, backtracks {log['num_backtracks']}/{log64['num_backtracks']}, "
          f"delta rel err vs f64: kernel {_rel(d, d64):.2e}, torch-f32 {_rel(d32, d64):.2e}")
    assert status == L.RL_OK and log64["error"] is None
    assert log["num_steps"] == int(valid.sum())
    assert log["num_backtracks"] == log64["num_backtracks"]
    np.testing.assert_allclose(log["entropy"], log64["entropy"], rtol=1e-5)
    np.testing.assert_allclose(log["loss_initial"], log64["loss_initial"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(log["loss_final"], log64["loss_final"], rtol=1e-5, atol=1e-7)
    assert _rel(d, d64) <= max(2e-5, 1.25 * _rel(d32, d64)) and _rel(d, d64) <= 1e-4


@pytest.mark.parametrize("cfg,hidden,activation", [SHAPES[0], SHAPES[1], SHAPES[4]] + DEEP)
def test_value_update_any_shape(ctx, cfg, hidden, activation):
    """ValuesOpt::update (opt.rs:100-127) with a non-default state-value module: 15 Adam steps on the reward-to-go MSE."""
    E, T, steps = 80, 50, 15
    env, traj, net, params, host, valid, adv, F, A = _collect(ctx, cfg, hidden, activation, E, T, seed=9)
    rng = np.random.default_rng(10)
    vparams = R.init_params(rng, F, hidden, 1)
    vcfg = R.ValuesOptConfig(state_value_fn_config=R.MlpConfig(hidden_sizes=_hs(hidden), activation=activation),
                             opt_steps_per_update=steps)
    critic = R.ValuesOpt(ctx, vcfg, F, float(env.discount_factor))
    critic.state_value_fn.set_weights(vparams)
    stats = critic.update(traj)
    new = critic.state_value_fn.get_weights()
    gamma = np.float32(min(0.99, float(env.discount_factor)))  # opt.rs:73
    rtg = np.zeros((T, E), np.float32)
    for e in range(E):
        n = int(host["lane_len"][e])
        rtg[:n, e] = O.discounted_cumsum_lane(host["reward"][:n, e], host["succ"][:n, e], gamma)
    obs, tgt = host["obs"][valid], rtg[valid]
    with TO.mlp_activation(activation):
        new64, losses64, _ = TO.value_update(vparams, F, hidden, obs, tgt, n_steps=steps, dtype=torch.float64)
        new32, losses32, _ = TO.value_update(vparams, F, hidden, obs, tgt, n_steps=steps, dtype=torch.float32)
    d, d64, d32 = new - vparams, new64 - vparams.astype(np.float64), new32 - vparams
    print(f"critic {F}->{hidden}->1 {activation}: delta rel err vs f64 kernel {_rel(d, d64):.2e}, torch-f32 {_rel(d32, d64):.2e}")
    assert stats.num_steps == int(valid.sum()) and stats.opt_steps == steps
    np.testing.assert_allclose(stats.loss_first, losses64[0], rtol=1e-5)
    np.testing.assert_allclose(stats.loss_last, losses64[-1], rtol=1e-4)
    assert _rel(d, d64) <= max(2e-4, 4 * _rel(d32, d64) + 1e-5)


@pytest.mark.parametrize("cfg,hidden,activation", [SHAPES[1], SHAPES[2], DEEP[0], DEEP[2]])
def test_ppo_and_reinforce_any_shape(ctx, cfg, hidden, activation):
    """Ppo::update / Reinforce::update (ppo.rs:97-147, reinforce.rs:64-89) with non-default policies."""
    E, T, steps, clip = 64, 50, 8, 0.1
    env, traj, net, params, host, valid, adv, F, A = _collect(ctx, cfg, hidden, activation, E, T, seed=11)
    adv_d = ctx.to_device(adv)
    obs, act, a = host["obs"][valid], host["action"][valid], adv[valid]
    log = {}
    R.Ppo(net, R.PpoConfig(opt_steps_per_update=steps, clip_distance=clip)).update(traj, adv_d, log)
    new = net.get_weights()
    with TO.mlp_activation(activation):
        new64, l64, ent64 = TO.ppo_update(params, F, hidden, A, obs, act, a, steps, clip, dtype=torch.float64)
        new32, l32, _ = TO.ppo_update(params, F, hidden, A, obs, act, a, steps, clip, dtype=torch.float32)
    d, d64, d32 = new - params, new64 - params.astype(np.float64), new32 - params
    print(f"ppo {F}->{hidden}->{A} {activation}: delta rel err vs f64 kernel {_rel(d, d64):.2e}, torch-f32 {_rel(d32, d64):.2e}")
    np.testing.assert_allclose(log["entropy"], ent64, rtol=1e-5)
    np.testing.assert_allclose(log["loss_first"], l64[0], rtol=1e-5, atol=1e-7)
    assert _rel(d, d64) <= max(2e-4, 4 * _rel(d32, d64) + 1e-5)
    # REINFORCE from the same starting point
    net.set_weights(params)
    log = {}
    R.Reinforce(net, R.ReinforceConfig()).update(traj, adv_d, log)
    new = net.get_weights()
    with TO.mlp_activation(activation):
        new64, l64, ent64 = TO.reinforce_update(params, F, hidden, A, obs, act, a, dtype=torch.float64)
        new32, _, _ = TO.reinforce_update(params, F, hidden, A, obs, act, a, dtype=torch.float32)
    d, d64, d32 = new - params, new64 - params.astype(np.float64), new32 - params
    np.testing.assert_allclose(log["entropy"], ent64, rtol=1e-5)
    np.testing.assert_allclose(log["loss_first"], l64, rtol=1e-5, atol=1e-7)
    assert _rel(d, d64) <= max(2e-3, 4 * _rel(d32, d64) + 1e-5)


@pytest.mark.parametrize("hidden", [[32], [24, 16]], ids=["one-hidden-layer", "two-hidden-layers"])
def test_actor_critic_learns_chain_with_small_tanh_mlp(ctx, hidden):
    """agents/testing.rs:14-64 in spirit, on an env and modules the default kernels do not serve: TRPO with 32-unit
    tanh networks on Chain (chain.rs: always going right pays 10 at the end of the chain, going left pays 2 at once)
    raises the mean step reward."""
    E, T = 512, 64
    env = R.build_env(ctx, R.Chain(), E, seed=3)
    mc = R.MlpConfig(hidden_sizes=hidden, activation="tanh")
    agent = R.ActorCriticConfig(policy_config=R.TrpoConfig(policy_fn_config=mc),
                                critic_config=R.ValuesOptConfig(state_value_fn_config=mc)).build_agent(env)
    rng = np.random.default_rng(0)
    agent.policy.policy_fn.set_weights(R.init_params(rng, env.num_features, hidden, 2))
    agent.critic.state_value_fn.set_weights(R.init_params(rng, env.num_features, hidden, 1))
    traj = R.Trajectory(env, T)
    rewards = []
    for period in range(12):
        summ = R.rollout(env, agent.actor(), R.HistoryDataBound(T, 0), traj)
        rewards.append(summ.step_reward.mean)
        agent.batch_update(traj, {})
    print("chain mean step reward per period:", [round(x, 3) for x in rewards])
    assert max(rewards[-3:]) > 1.15 * rewards[0]
