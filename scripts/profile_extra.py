"""DQN (config 3) and GRU (config 4) kernels for an ncu capture: python scripts/profile_extra.py [dqn|gru]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R
from relearn_b200 import _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "dqn"
ctx = R.Context(0)
rng = np.random.default_rng(0)
if which == "dqn":
    E, cfg = 65536, R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    env = R.build_env(ctx, cfg, E, seed=5)
    agent = R.DqnConfig(buffer_capacity=762, opt_steps_per_update=3).build_agent(env)
    agent.action_value_fn.set_weights(R.init_params(rng, 5, 128, 2))
    rb = agent.buffer()
    bound = R.HistoryDataBound(16, 5)
    traj = R.Trajectory(env, 21)
    for _ in range(3):
        R.rollout(env, agent.actor(), bound, traj, want_summary=False)
        rb.write_experience(traj)
        agent.batch_update(rb)
else:
    env = R.build_env(ctx, R.MetaEnv(R.UniformBernoulliBandits(2), 10), 131072, seed=6)
    net = R.GruLinear(ctx, env.num_features, 4, env.num_actions)
    net.set_weights(R.init_gru_linear_params(rng, env.num_features, 4, env.num_actions))
    traj = R.Trajectory(env, 190)
    for _ in range(3):
        R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, seq_net=net), R.HistoryDataBound(190, 0), traj,
                  want_summary=False)
ctx.synchronize()
print("done", which)
