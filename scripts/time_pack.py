"""Time rl_pack_history on a [256][131072] CartPole trajectory (33.5 M steps): ms and GB/s of the 79 algorithmic bytes per step."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R
from relearn_b200 import _lib as L

ctx = R.Context(0)
lib = ctx._lib
T, E = int(os.environ.get("T", 256)), int(os.environ.get("E", 1 << 17))
cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
env = R.build_env(ctx, cfg, E, seed=8)
traj = R.Trajectory(env, T)
net = R.Mlp(ctx, 5, [128], 2)
net.set_weights(R.init_params(np.random.default_rng(1), 5, 128, 2))
R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net), R.HistoryDataBound(T, 0), traj, want_summary=False)
cap = T * E
bufs = [ctx.alloc(cap * 20), ctx.alloc(2 * cap * 20), ctx.alloc(2 * cap), ctx.alloc(cap * 8), ctx.alloc(cap * 4), ctx.alloc((T + 1) * 8),
        ctx.alloc((T + 2) * 8)]
info = L.PackedInfo()
call = lambda: L.check(lib.rl_pack_history(traj.handle, *[b.c for b in bufs], C.byref(info)), ctx.handle)
for _ in range(2):
    call()
e0 = ctx.event().record()
for _ in range(5):
    call()
e1 = ctx.event().record()
ms = e0.elapsed_ms(e1) / 5
print(f"rl_pack_history: {info.num_steps} steps, {info.num_episodes} episodes, longest {info.max_len}: {ms:.3f} ms, "
      f"{79 * info.num_steps / (ms * 1e-3) / 1e9:.0f} GB/s of 79 B per step")
