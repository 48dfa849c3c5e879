"""Fused CartPole rollout (E = 4096, T = 256) with non-default policy modules: ms per period by the kernel `lanes_per_env = 0` picks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R
from relearn_b200 import _lib as L

ctx = R.Context(0)
E, T = int(os.environ.get("E", 4096)), 256
cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
rng = np.random.default_rng(0)
for hidden, act in (([128], "relu"), ([128], "tanh"), ([64], "relu"), ([256], "relu"), ([64, 64], "tanh"), ([32, 32, 32], "relu")):
    env = R.build_env(ctx, cfg, E, seed=1)
    net = R.Mlp(ctx, 5, hidden, 2, act)
    net.set_weights(R.init_params(rng, 5, hidden, 2))
    traj = R.Trajectory(env, T)
    spec = R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net)
    for _ in range(2):
        R.rollout(env, spec, R.HistoryDataBound(T, 0), traj, want_summary=False)
    e0 = ctx.event().record()
    for _ in range(3):
        R.rollout(env, spec, R.HistoryDataBound(T, 0), traj, want_summary=False)
    e1 = ctx.event().record()
    ms = e0.elapsed_ms(e1) / 3
    print(f"5-{'-'.join(map(str, hidden))}-2 {act}: {ms:.3f} ms per period, {E * T / ms / 1e6:.1f} G env-steps/s", flush=True)
    traj.close(); env.close()
