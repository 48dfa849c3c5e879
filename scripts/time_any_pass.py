"""Time the run-time-sized pass kernel (mlp_pass_any_kernel, update.cu) on non-default modules: a probe = statistics +
loss / KL + gradient + one Fisher-vector product, and a whole critic update of 80 Adam steps, on a 262 144-step CartPole
batch; the default 5-128-2 / 5-128-1 modules (tensor-core passes) beside them."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R
from relearn_b200 import _lib as L

ctx = R.Context(0)
E, T = 4096, 64
cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
env = R.build_env(ctx, cfg, E, seed=1)
rng = np.random.default_rng(0)
for hidden, act in (([128], "relu"), ([64], "tanh"), ([256], "relu"), ([64, 64], "tanh"), ([32, 32, 32], "relu")):
    mc = R.MlpConfig(hidden_sizes=hidden, activation=act)
    agent = R.ActorCriticConfig(policy_config=R.TrpoConfig(policy_fn_config=mc),
                                critic_config=R.ValuesOptConfig(state_value_fn_config=mc)).build_agent(env)
    agent.policy.policy_fn.set_weights(R.init_params(rng, 5, hidden, 2))
    agent.critic.state_value_fn.set_weights(R.init_params(rng, 5, hidden, 1))
    traj = R.Trajectory(env, T)
    R.rollout(env, agent.actor(), R.HistoryDataBound(T, 0), traj, want_summary=False)
    adv = ctx.to_device(rng.normal(size=(T, E)).astype(np.float32))
    vec = rng.normal(size=agent.policy.policy_fn.num_params).astype(np.float32)
    agent.policy.probe(traj, adv, vec)
    e0 = ctx.event().record()
    for _ in range(3):
        agent.policy.probe(traj, adv, vec)
    e1 = ctx.event().record()
    probe_ms = e0.elapsed_ms(e1) / 3
    log = {}
    agent.policy.update(traj, adv, log)
    st = agent.critic.update(traj, {})
    print(f"5-{'-'.join(map(str, hidden))}-2 {act}: probe (stats + loss/KL + grad + FVP) {probe_ms:.3f} ms, TRPO update "
          f"{log['policy/update_time'] * 1e3:.2f} ms ({log['cg_iterations']} CG iterations, {log['num_backtracks']} backtracks), "
          f"critic 80 Adam steps {st.update_ms:.2f} ms  [{T * E} steps]", flush=True)
    traj.close()
