"""Data-parallel check, run under torchrun with WORLD_SIZE ranks (one per GPU):

1. rl_ctx_allreduce_f64 sums across ranks (NCCL over NVLink);
2. sharded rollouts equal the corresponding lanes of a single full-size env (Philox is keyed by the global lane);
3. GAE + TRPO + critic on lane shards with all-reduced sums match the same update on the full batch on one GPU.

Exit code 0 = all checks passed on every rank.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import relearn_b200 as R
from relearn_b200 import _lib as L


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = R.Context(local)
    ids = [R.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)

    # 1. all-reduce
    buf = ctx.to_device(np.arange(5, dtype=np.float64) + rank)
    L.check(ctx._lib.rl_ctx_allreduce_f64(ctx.handle, buf.c, 5), ctx.handle)
    got = buf.download((5,), np.float64)
    want = world * np.arange(5, dtype=np.float64) + sum(range(world))
    assert np.array_equal(got, want), (got, want)

    # 2. + 3. sharded vs full
    E, T, seed = 512, 96, 11
    REG = float(os.environ.get("RL_CHECK_HPV_REG", "0.1"))
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    rng = np.random.default_rng(0)
    pparams, vparams = R.init_params(rng, 5, 128, 2), R.init_params(rng, 5, 128, 1)

    def run(context, n_lanes, offset):
        env = R.build_env(context, cfg, n_lanes, seed=seed, lane_offset=offset)
        # hpv_reg_coeff 0.1: ten f32 CG iterations stay well conditioned, so the comparison measures the reduction over the
        # ranks and not the amplification of its rounding by the reference's own 1e-5 (tests/test_gpu_update.py; with 1e-5
        # this 48 k-step batch gives 0 at 2 ranks and 3e-3 at 8)
        agent = R.ActorCriticConfig(policy_config=R.TrpoConfig(optimizer_config=R.ConjugateGradientOptimizerConfig(hpv_reg_coeff=REG)),
                                    critic_config=R.ValuesOptConfig(opt_steps_per_update=10)).build_agent(env)
        agent.policy.policy_fn.set_weights(pparams)
        agent.critic.state_value_fn.set_weights(vparams)
        traj = R.Trajectory(env, T)
        R.rollout(env, agent.actor(), R.HistoryDataBound(T, 0), traj)
        log = {}
        status = agent.batch_update(traj, log)
        return traj.to_host(), agent.policy.policy_fn.get_weights(), agent.critic.state_value_fn.get_weights(), status, log

    per = E // world
    host_s, p_s, v_s, st_s, log_s = run(ctx, per, rank * per)
    assert log_s["num_steps"] > host_s["num_steps"] or world == 1, "the update must see the global batch"
    ok = True
    if rank == 0:
        solo = R.Context(local)  # no communicator: the whole batch on one GPU
        host_f, p_f, v_f, st_f, log_f = run(solo, E, 0)
        for k in ("obs", "action", "reward", "succ"):
            assert np.array_equal(host_f[k][:, :per], host_s[k]), f"shard 0 differs from the full env in {k}"
        rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b - (pparams if b.size == pparams.size else vparams)), 1e-30))
        dp, dv = rel(p_s, p_f), rel(v_s, v_f)
        print(f"world={world} hpv_reg_coeff={REG}: N={log_s['num_steps']} (full {log_f['num_steps']}), status {st_s}/{st_f}, backtracks "
              f"{log_s['num_backtracks']}/{log_f['num_backtracks']}, policy delta rel diff {dp:.2e}, critic delta rel diff {dv:.2e}",
              flush=True)
        ok = (log_s["num_steps"] == log_f["num_steps"] and st_s == st_f and log_s["num_backtracks"] == log_f["num_backtracks"]
              and dp < 1e-4 and dv < 1e-4)
    peer = ctx.comm_peer_info()
    want_peer = os.environ.get("RL_XREDUCE", "")[:1] not in ("n", "0")
    if rank == 0:
        print("update reductions:", "fused row-reduction + NVLink peer exchange" if peer["peer_mailboxes"] else "reduce + ncclAllReduce",
              flush=True)
    ok = ok and not peer["timed_out"] and (peer["peer_mailboxes"] or not want_peer)
    # every rank must hold identical parameters afterwards (no broadcast is needed by construction)
    t = torch.tensor(np.concatenate([p_s, v_s]), device="cuda")
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(t, ref))
    flag = torch.tensor([1.0 if (ok and same) else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("parameters bit-identical across ranks:", same, "| all checks:", bool(flag.item() == 1.0), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
