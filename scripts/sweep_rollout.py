"""Fused rollout throughput vs lanes and threads-per-env (BASELINE config 5 sweep, one GPU)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R
from relearn_b200 import _lib as L

ctx = R.Context(0)
cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
net = R.Mlp(ctx, 5, [128], 2)
net.set_weights(R.init_params(np.random.default_rng(0), 5, 128, 2))
Es = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1024, 4096, 16384, 65536, 262144, 1 << 20]
LANES = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8, 16, 32, L.RL_LANES_TENSOR_CORE]
out = []
for E in Es:
    T = 256 if E <= 65536 else (64 if E <= (1 << 20) else 16)
    env = R.build_env(ctx, cfg, E, seed=1)
    traj = R.Trajectory(env, T)
    for lanes in LANES:
        if lanes != L.RL_LANES_TENSOR_CORE and E * lanes > (1 << 25):
            continue
        spec = R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net, lanes_per_env=lanes)
        for _ in range(2):
            R.rollout(env, spec, R.HistoryDataBound(T, 0), traj, want_summary=False)
        reps = 5
        e0 = ctx.event().record()
        for _ in range(reps):
            R.rollout(env, spec, R.HistoryDataBound(T, 0), traj, want_summary=False)
        e1 = ctx.event().record()
        ms = e0.elapsed_ms(e1) / reps
        rec = {"envs": E, "horizon": T, "lanes_per_env": lanes, "ms": round(ms, 4), "env_steps_per_s": E * T / (ms * 1e-3)}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    traj.close()
    env.close()
