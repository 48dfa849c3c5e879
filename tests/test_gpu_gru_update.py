"""GPU parity: TRPO / critic updates and GAE through a recurrent module (Chain<Gru, Linear>, config 4) vs the
torch-CPU restatement, which differentiates the composed gru_cell with autograd exactly as the reference does with
cuDNN disabled (trpo.rs:104-108).

Tolerances as in test_gpu_update.py: each quantity is compared with the same algorithm run in f64; gradients and
Fisher-vector products within 1e-5 relative (norm-wise), parameter deltas of a whole trust-region step within 2e-5
(or 1.25x what the torch f32 run itself achieves)."""
import numpy as np
import pytest

import oracle as O
from oracle import tensor_oracle as TO
import relearn_b200 as R
from relearn_b200 import _lib as L

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) /
                 max(np.linalg.norm(np.asarray(b, np.float64)), 1e-300))


def _lane_episodes(host, e):
    n = int(host["lane_len"][e])
    ends = np.flatnonzero(host["succ"][:n, e] != L.RL_CONTINUE)
    starts = np.concatenate([[0], ends[:-1] + 1]) if len(ends) else np.zeros(0, int)
    return [(int(a), int(b) + 1) for a, b in zip(starts, ends)]


def _collect(ctx, hidden, arms, episodes, E, T, seed, scale=1.0, activation="relu"):
    rng = np.random.default_rng(seed)
    cfg = R.MetaEnv(R.UniformBernoulliBandits(arms), episodes)
    env = R.build_env(ctx, cfg, E, seed=seed)
    F, A = env.num_features, env.num_actions
    params = (R.init_gru_linear_params(rng, F, hidden, A) * scale).astype(np.float32)
    o = 3 * hidden * F + 3 * hidden * hidden
    params[o:o + 6 * hidden] = rng.normal(size=6 * hidden) * 0.2
    net = R.GruLinear(ctx, F, hidden, A, activation)
    net.set_weights(params)
    traj = R.Trajectory(env, T)
    R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, seq_net=net), R.HistoryDataBound(T, 0), traj)
    host = traj.to_host()
    spans = [(e, a, b) for e in range(E) for a, b in _lane_episodes(host, e)]
    assert sum(b - a for _, a, b in spans) == int((host["succ"] != L.RL_PAD).sum())
    return env, traj, net, params, host, spans


def _per_episode(x, spans):
    return [x[a:b, e] for e, a, b in spans]


# hidden > 8 runs K10 (gru_big.cu: tiled GEMMs over the lanes of a step); (128, 10, .) is the rl2-sized module
# (rl2-bandits.rs:379-393: GRU(14 -> 128) -> ReLU -> Linear(128 -> 10))
@pytest.mark.parametrize("hidden,arms,episodes,activation", [(4, 2, 10, "relu"), (8, 5, 4, "tanh"), (3, 3, 6, "relu"),
                                                             (24, 3, 4, "tanh"), (128, 10, 4, "relu")])
def test_seq_trpo_probe_loss_grad_fvp(ctx, hidden, arms, episodes, activation):
    E, T = (80 if hidden <= 24 else 45), 2 * (2 * episodes - 1) + 5
    env, traj, net, params, host, spans = _collect(ctx, hidden, arms, episodes, E, T, seed=hidden + arms, scale=1.5,
                                                   activation=activation)
    F, A = env.num_features, env.num_actions
    rng = np.random.default_rng(2)
    adv = rng.normal(size=(T, E)).astype(np.float32)
    vec = rng.normal(size=net.num_params).astype(np.float32)
    got = R.Trpo(net, R.TrpoConfig()).probe(traj, ctx.to_device(adv), vec)
    eps, acts, advs = _per_episode(host["obs"], spans), _per_episode(host["action"], spans), _per_episode(adv, spans)
    reg = 1e-5
    loss64, kl64, ent64, g64, hv64 = TO.seq_policy_loss_kl_grad_fvp(params, F, hidden, A, eps, acts, advs, vec, reg,
                                                                   torch.float64, activation)
    _, _, _, g32, hv32 = TO.seq_policy_loss_kl_grad_fvp(params, F, hidden, A, eps, acts, advs, vec, reg, torch.float32,
                                                        activation)
    print(f"episodes={len(spans)} grad rel err: kernel {_rel(got['grad'], g64):.2e} torch-f32 {_rel(g32, g64):.2e}; "
          f"fvp rel err: kernel {_rel(got['fvp'], hv64):.2e} torch-f32 {_rel(hv32, hv64):.2e}")
    assert abs(got["loss"] - loss64) <= 1e-6 * max(1.0, abs(loss64))
    assert abs(got["kl"]) <= 1e-7 and abs(kl64) <= 1e-12
    assert abs(got["entropy"] - ent64) <= 1e-6
    assert _rel(got["grad"], g64) <= 1e-5
    assert _rel(got["fvp"], hv64) <= 1e-5


@pytest.mark.parametrize("hidden,arms,episodes,reg", [(4, 2, 10, 0.1), (8, 4, 5, 0.1), (128, 10, 3, 0.1)])
def test_seq_trpo_update_matches_f64(ctx, hidden, arms, episodes, reg):
    E, T = (96 if hidden <= 8 else 40), 2 * (2 * episodes - 1) + 3
    env, traj, net, params, host, spans = _collect(ctx, hidden, arms, episodes, E, T, seed=7 + hidden, scale=1.5)
    F, A = env.num_features, env.num_actions
    rng = np.random.default_rng(11)
    adv = rng.normal(size=(T, E)).astype(np.float32)
    policy = R.Trpo(net, R.TrpoConfig(optimizer_config=R.ConjugateGradientOptimizerConfig(hpv_reg_coeff=reg)))
    log = {}
    status = policy.update(traj, ctx.to_device(adv), log)
    new = net.get_weights()
    eps, acts, advs = _per_episode(host["obs"], spans), _per_episode(host["action"], spans), _per_episode(adv, spans)
    ocfg = TO.CgConfig(hpv_reg_coeff=reg)
    new64, log64 = TO.seq_trpo_update(params, F, hidden, A, eps, acts, advs, cfg=ocfg, dtype=torch.float64)
    new32, log32 = TO.seq_trpo_update(params, F, hidden, A, eps, acts, advs, cfg=ocfg, dtype=torch.float32)
    d, d64, d32 = new - params, new64 - params.astype(np.float64), new32 - params
    print(f"status={status} backtracks kernel/f64/f32 = {log['num_backtracks']}/{log64['num_backtracks']}/"
          f"{log32['num_backtracks']}; delta rel err vs f64: kernel {_rel(d, d64):.2e}, torch-f32 {_rel(d32, d64):.2e}")
    assert status == L.RL_OK and log64["error"] is None
    assert log["num_steps"] == sum(b - a for _, a, b in spans)
    assert log["num_backtracks"] == log64["num_backtracks"]
    np.testing.assert_allclose(log["entropy"], log64["entropy"], rtol=1e-5)
    np.testing.assert_allclose(log["step_size"], log64["step_size"], rtol=2e-5)
    np.testing.assert_allclose(log["loss_initial"], log64["loss_initial"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(log["loss_final"], log64["loss_final"], rtol=1e-5, atol=1e-7)
    assert _rel(d, d64) <= max(2e-5, 1.25 * _rel(d32, d64)) and _rel(d, d64) <= 1e-4


def _value_params(rng, F, hidden):
    p = R.init_gru_linear_params(rng, F, hidden, 1)
    o = 3 * hidden * F + 3 * hidden * hidden
    p[o:o + 6 * hidden] = rng.normal(size=6 * hidden) * 0.2
    return p.astype(np.float32)


@pytest.mark.parametrize("hidden,arms,episodes,E,steps", [(4, 2, 6, 72, 20), (128, 10, 4, 40, 6)])
def test_seq_gae_and_value_update(ctx, hidden, arms, episodes, E, steps):
    """ValuesOpt<Chain<Gru, Linear>>: GAE over SeqPacked values (interrupted trials bootstrap from one more step of
    the same sequence, critics/mod.rs:116-131) and Adam steps on the reward-to-go targets."""
    T = 2 * (2 * episodes - 1) + 4
    env, traj, net, params, host, spans = _collect(ctx, hidden, arms, episodes, E, T, seed=21)
    F = env.num_features
    rng = np.random.default_rng(5)
    vparams = _value_params(rng, F, hidden)
    critic = R.ValuesOpt(ctx, R.ValuesOptConfig(state_value_fn_config=R.GruLinearConfig(hidden_dim=hidden),
                                                opt_steps_per_update=steps), F, env.discount_factor)
    critic.state_value_fn.set_weights(vparams)
    gamma, lam = critic.discount_factor, np.float32(0.95)
    adv = critic.advantages(traj).download((T, E), np.float32)
    # oracle: extended state values per episode, then the reference scan per lane
    v = np.zeros((T, E), np.float32)
    vn = np.zeros((T, E), np.float32)
    n_intr = 0
    for e, a, b in spans:
        obs = host["obs"][a:b, e]
        if host["succ"][b - 1, e] == L.RL_INTERRUPT:
            ext = np.concatenate([obs, host["next_obs"][b - 1:b, e]])
            out = TO.gru_linear_episode(vparams, F, hidden, 1, ext)[:, 0]
            v[a:b, e], vn[b - 1, e] = out[:-1], out[-1]
            n_intr += 1
        else:
            v[a:b, e] = TO.gru_linear_episode(vparams, F, hidden, 1, obs)[:, 0]
    assert n_intr > E  # every finished trial ends in Interrupt
    for e in range(E):
        n = int(host["lane_len"][e])
        ref = O.gae_lane(host["reward"][:n, e], v[:n, e], vn[:n, e], host["succ"][:n, e], gamma, lam)
        np.testing.assert_allclose(adv[:n, e], ref, rtol=2e-5, atol=2e-6)
    # critic update
    stats = critic.update(traj)
    new = critic.state_value_fn.get_weights()
    rtg = np.zeros((T, E), np.float32)
    for e in range(E):
        n = int(host["lane_len"][e])
        rtg[:n, e] = O.discounted_cumsum_lane(host["reward"][:n, e], host["succ"][:n, e], gamma)
    eps, tg = _per_episode(host["obs"], spans), _per_episode(rtg, spans)
    new64, losses64 = TO.seq_value_update(vparams, F, hidden, eps, tg, n_steps=steps, dtype=torch.float64)
    new32, losses32 = TO.seq_value_update(vparams, F, hidden, eps, tg, n_steps=steps, dtype=torch.float32)
    d, d64, d32 = new - vparams, new64 - vparams.astype(np.float64), new32 - vparams
    print(f"GRU critic delta rel err vs f64: kernel {_rel(d, d64):.2e}, torch-f32 {_rel(d32, d64):.2e}; "
          f"loss {stats.loss_first:.6f}->{stats.loss_last:.6f} vs {losses64[0]:.6f}->{losses64[-1]:.6f}")
    assert stats.opt_steps == steps and stats.num_steps == sum(b - a for _, a, b in spans)
    np.testing.assert_allclose(stats.loss_first, losses64[0], rtol=1e-5)
    np.testing.assert_allclose(stats.loss_last, losses64[-1], rtol=1e-4)
    assert _rel(d, d64) <= max(2e-4, 4 * _rel(d32, d64) + 1e-5)


def test_gemm_formulation_serves_the_small_modules_too():
    """K10 against the oracle on the modules K9 (one thread per lane) normally serves: the same tests in a child process
    with RL_SEQ_FORCE_BIG=1 (the switch is read once per process)."""
    import os
    import subprocess
    import sys

    if os.environ.get("RL_SEQ_FORCE_BIG") == "1":
        pytest.skip("already the forced run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RL_SEQ_FORCE_BIG="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_gru_update.py"), "-x", "-q", "-m", "gpu",
                        "-k", "(probe or update_matches or gae_and_value) and not 128 and not 24"], capture_output=True, text=True,
                       env=env, cwd=root, timeout=1200)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0


@pytest.mark.parametrize("hidden", [8, 128])
def test_rl2_actor_critic_learns_bandits(ctx, hidden):
    """Behavioural check in the spirit of rl2-bandits.rs / agents/testing.rs: TRPO with a GRU policy and a GRU critic
    on 2-armed Bernoulli bandit trials raises the mean per-step reward of the meta-episodes (hidden 128 = the module size
    of rl2-bandits.rs:379-393, served by K10)."""
    arms, episodes, E = 2, 10, 2048
    T = 2 * episodes - 1
    env = R.build_env(ctx, R.MetaEnv(R.UniformBernoulliBandits(arms), episodes), E, seed=4)
    gcfg = R.GruLinearConfig(hidden_dim=hidden)
    agent = R.ActorCriticConfig(policy_config=R.TrpoConfig(policy_fn_config=gcfg),
                                critic_config=R.ValuesOptConfig(state_value_fn_config=gcfg, opt_steps_per_update=20)
                                ).build_agent(env)
    rng = np.random.default_rng(0)
    agent.policy.policy_fn.set_weights(R.init_gru_linear_params(rng, env.num_features, hidden, arms))
    agent.critic.state_value_fn.set_weights(R.init_gru_linear_params(rng, env.num_features, hidden, 1))
    traj = R.Trajectory(env, T)
    means = []
    for period in range(25):
        summ = R.rollout(env, agent.actor(), R.HistoryDataBound(T, 0), traj)
        means.append(summ.step_reward.mean)
        agent.batch_update(traj, {})
    print("mean step reward per period:", [round(x, 4) for x in means])
    # random play earns 0.5 * (episodes / T); a policy that exploits the better arm earns noticeably more
    assert np.mean(means[-3:]) > np.mean(means[:3]) * 1.08
