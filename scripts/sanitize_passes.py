"""Small run of every tensor-core pass (critic, TRPO stats / grad / FVP / eval, PPO, REINFORCE) for compute-sanitizer:
`compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_passes.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import relearn_b200 as R

ctx = R.Context(0)
cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
E, T = 300, 37  # 11 100 samples: 87 tiles, the last one partly padding
env = R.build_env(ctx, cfg, E, seed=1)
rng = np.random.default_rng(0)
for pc in (R.TrpoConfig(), R.PpoConfig(), R.ReinforceConfig()):
    agent = R.ActorCriticConfig(policy_config=pc, critic_config=R.ValuesOptConfig(opt_steps_per_update=3)).build_agent(env)
    agent.policy.policy_fn.set_weights(R.init_params(rng, 5, 128, 2))
    agent.critic.state_value_fn.set_weights(R.init_params(rng, 5, 128, 1))
    traj = R.Trajectory(env, T)
    R.rollout(env, agent.actor(), R.HistoryDataBound(T, 0), traj, want_summary=False)
    log = {}
    status = agent.batch_update(traj, log)
    ctx.synchronize()
    print(type(agent.policy).__name__, "status", status, "steps", log.get("num_steps"))
print("done")
