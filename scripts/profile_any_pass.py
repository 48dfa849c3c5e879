"""One critic update (3 Adam steps) through mlp_pass_any_kernel on a 5-64-64-1 tanh module, for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R

ctx = R.Context(0)
E, T = 4096, 64
hidden = [int(x) for x in os.environ.get("HIDDEN", "64,64").split(",")]
env = R.build_env(ctx, R.CartPoleConfig().wrap(R.VisibleStepLimit(500)), E, seed=1)
rng = np.random.default_rng(0)
mc = R.MlpConfig(hidden_sizes=hidden, activation="tanh")
agent = R.ActorCriticConfig(policy_config=R.TrpoConfig(policy_fn_config=mc),
                            critic_config=R.ValuesOptConfig(state_value_fn_config=mc, opt_steps_per_update=3)).build_agent(env)
agent.policy.policy_fn.set_weights(R.init_params(rng, 5, hidden, 2))
agent.critic.state_value_fn.set_weights(R.init_params(rng, 5, hidden, 1))
traj = R.Trajectory(env, T)
R.rollout(env, agent.actor(), R.HistoryDataBound(T, 0), traj, want_summary=False)
agent.critic.update(traj, {})
ctx.synchronize()
