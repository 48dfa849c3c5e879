"""Run each hot kernel a few times so that `ncu -k regex:<name>` can capture it (one GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R
from relearn_b200 import _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = R.Context(0)
cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
rng = np.random.default_rng(0)
net = R.Mlp(ctx, 5, [128], 2)
net.set_weights(R.init_params(rng, 5, 128, 2))

if which in ("all", "rollout"):
    for lanes in (0, 1):
        E, T = (4096, 256) if lanes == 0 else (1 << 20, 32)
        env = R.build_env(ctx, cfg, E, seed=1)
        traj = R.Trajectory(env, T)
        for _ in range(3):
            R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net, lanes_per_env=lanes),
                      R.HistoryDataBound(T, 0), traj, want_summary=False)
        ctx.synchronize()
        traj.close(); env.close()

if which in ("all", "step"):
    E = 1 << 22
    env = R.build_env(ctx, cfg, E, seed=7)
    env.reset_all()
    actions = ctx.to_device(rng.integers(0, 2, E).astype(np.uint8))
    for _ in range(4):
        env.step_device(actions)
    ctx.synchronize()
    env.close()

if which in ("all", "update", "scan"):
    E, T = 4096, 256
    env = R.build_env(ctx, cfg, E, seed=1)
    agent = R.ActorCriticConfig().build_agent(env)
    agent.policy.policy_fn.set_weights(R.init_params(rng, 5, 128, 2))
    agent.critic.state_value_fn.set_weights(R.init_params(rng, 5, 128, 1))
    traj = R.Trajectory(env, T)
    R.rollout(env, agent.actor(), R.HistoryDataBound(T, 0), traj, want_summary=False)
    if which in ("all", "scan"):
        E2 = 1 << 17
        env2 = R.build_env(ctx, cfg, E2, seed=8)
        traj2 = R.Trajectory(env2, T)
        R.rollout(env2, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net), R.HistoryDataBound(T, 0), traj2, want_summary=False)
        adv, rtg = ctx.alloc(T * E2 * 4), ctx.alloc(T * E2 * 4)
        for _ in range(3):
            L.check(ctx._lib.rl_gae(traj2.handle, None, 0.99, 0.95, adv.c, rtg.c), ctx.handle)
        ctx.synchronize()
    if which in ("all", "update"):
        agent.critic.cfg.opt_steps_per_update = 3
        for _ in range(2):
            agent.batch_update(traj, {})
        ctx.synchronize()
print("done", which)
