"""One fused CartPole rollout configuration for an ncu capture: python scripts/profile_rollout_small.py E T lanes"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R
from relearn_b200 import _lib as L

E, T, lanes = (int(x) for x in sys.argv[1:4])
ctx = R.Context(0)
cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
net = R.Mlp(ctx, 5, [128], 2)
net.set_weights(R.init_params(np.random.default_rng(0), 5, 128, 2))
env = R.build_env(ctx, cfg, E, seed=1)
traj = R.Trajectory(env, T)
for _ in range(4):
    R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net, lanes_per_env=lanes), R.HistoryDataBound(T, 0),
              traj, want_summary=False)
ctx.synchronize()
print("done")
