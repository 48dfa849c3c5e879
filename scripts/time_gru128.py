"""Rollout throughput of the rl2-sized recurrent policy (bandit meta-env, k = 10 arms, n = 100 episodes per trial,
Chain<Gru(14 -> 128), Linear(128 -> 10)>, rl2-bandits.rs:46-51,379-393): K8h (tiled GEMM) against K8a (thread per
env).  CUDA events on the library stream; prints one JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relearn_b200 as R  # noqa: E402
from relearn_b200 import _lib as L  # noqa: E402


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64 * 2
    arms, episodes = 10, 100
    T = 2 * episodes - 1
    ctx = R.Context(0)
    env = R.build_env(ctx, R.MetaEnv(R.UniformBernoulliBandits(arms), episodes), E, seed=4)
    F, A = env.num_features, env.num_actions
    net = R.GruLinear(ctx, F, 128, A)
    net.set_weights(R.init_gru_linear_params(np.random.default_rng(0), F, 128, A))
    traj = R.Trajectory(env, T)
    spec = R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, seq_net=net)
    out = {"envs": E, "horizon": T, "features": F, "hidden": 128, "actions": A, "flop_per_env_step": 2 * 3 * 128 * (F + 128) + 2 * 128 * A}
    kernels = sys.argv[2].split(",") if len(sys.argv) > 2 else ["stepped", "tile", "tile32", "thread"]
    for kernel, reps in (("stepped", 3), ("tile", 3), ("tile32", 3), ("thread", 1)):
        if kernel not in kernels:
            continue
        os.environ["RL_GRU_KERNEL"] = kernel
        R.rollout(env, spec, R.HistoryDataBound(T, 0), traj, want_summary=False)
        e0 = ctx.event().record()
        for _ in range(reps):
            summ = R.rollout(env, spec, R.HistoryDataBound(T, 0), traj)
        e1 = ctx.event().record()
        ms = e0.elapsed_ms(e1) / reps
        out[kernel] = {"ms": ms, "env_steps_per_s": E * T / (ms * 1e-3), "fp32_tflops": out["flop_per_env_step"] * E * T / (ms * 1e-3) / 1e12,
                       "mean_step_reward": summ.step_reward.mean}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
