"""One fused CartPole rollout (E = 4096, T = 256) with a 5-64-64-2 tanh policy (K2g, a warp per env), for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R
from relearn_b200 import _lib as L

ctx = R.Context(0)
hidden = [int(x) for x in os.environ.get("HIDDEN", "64,64").split(",")]
env = R.build_env(ctx, R.CartPoleConfig().wrap(R.VisibleStepLimit(500)), 4096, seed=1)
net = R.Mlp(ctx, 5, hidden, 2, os.environ.get("ACT", "tanh"))
net.set_weights(R.init_params(np.random.default_rng(0), 5, hidden, 2))
traj = R.Trajectory(env, 256)
for _ in range(3):
    R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net), R.HistoryDataBound(256, 0), traj, want_summary=False)
ctx.synchronize()
