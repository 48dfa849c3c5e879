"""world_size-2 `gloo` tests (CPU) of the N>1 path's host logic: lane sharding is invariant to the number
of ranks, per-rank summaries merge to the single-rank summary, and the update's "all-reduce sums, divide by
the global count" contract reproduces the full-batch gradient when ranks hold unequal numbers of steps.

The compute on each rank is the CPU oracle (this is a test of the sharding/reduction logic, not of the
kernels; the same logic runs over NCCL on the GPUs)."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rollout_shard(lanes, offset, params, horizon, seed):
    import ctypes as C

    import oracle as O

    out = []
    total = O.Summary()
    cfg = O.cartpole_cfg(500)
    for e in range(offset, offset + lanes):
        actor = O.policy_actor(params, 5, 128, 2)
        r = O.rollout_lane(cfg, actor, horizon, 0, O.PhiloxRng(seed, e, 0))
        out.append((r["obs"][:r["n"]].copy(), r["action"][:r["n"]].copy(), r["succ"][:r["n"]].copy()))
        O.lib().ro_summary_merge(C.byref(total), C.byref(r["summary"]))
    return out, total


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from relearn_b200 import init_params
    from relearn_b200.parallel import MeanVar, global_mean_from_sums, merge_summaries, shard_lanes

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        E, T, seed = 7, 40, 99  # 7 lanes over 2 ranks: unequal shards (4 + 3)
        params = init_params(np.random.default_rng(0), 5, 128, 2)
        count, offset = shard_lanes(E, rank, world)
        mine, summ = _rollout_shard(count, offset, params, T, seed)

        # (1) trajectories: gather every rank's lanes on rank 0 and compare with the one-rank run
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        # (2) summaries: all-gather the 3 x (mean, M2, count) partials and Chan-merge them
        part = torch.tensor(np.stack([MeanVar(s.mean, s.m2, s.count).to_array()
                                      for s in (summ.step_reward, summ.episode_reward, summ.episode_length)]))
        parts = [torch.zeros_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        merged = merge_summaries([[MeanVar.from_array(r) for r in p.numpy()] for p in parts])

        # (3) update reductions: per-rank SUMS of the loss gradient + count -> all-reduce -> divide by global N
        from oracle import tensor_oracle as TO

        obs = np.concatenate([o for o, _a, _s in mine])
        act = np.concatenate([a for _o, a, _s in mine])
        adv = np.cos(np.arange(len(act)) + 1000 * rank).astype(np.float64)
        loss, _kl, _ent, g, _hv = TO.policy_loss_kl_grad_fvp(params, 5, 128, 2, obs, act, adv, np.zeros(1026))
        n_local = len(act)

        def all_reduce_sum(a):
            t = torch.tensor(a)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t.numpy()

        g_global = global_mean_from_sums(np.concatenate([g * n_local, [loss * n_local]]), n_local, all_reduce_sum)
        payload = [(obs, act, adv)]
        allp = [None] * world
        dist.all_gather_object(allp, payload)

        if rank == 0:
            ref, ref_summ = _rollout_shard(E, 0, params, T, seed)
            flat = [lane for shard in gathered for lane in shard]
            assert len(flat) == E
            for (o1, a1, s1), (o2, a2, s2) in zip(flat, ref):
                assert np.array_equal(o1, o2) and np.array_equal(a1, a2) and np.array_equal(s1, s2)
            for m, s in zip(merged, (ref_summ.step_reward, ref_summ.episode_reward, ref_summ.episode_length)):
                assert m.count == s.count
                assert abs(m.mean - s.mean) < 1e-12 and abs(m.squared_residual_sum - s.m2) < 1e-9
            fo = np.concatenate([p[0][0] for p in allp])
            fa = np.concatenate([p[0][1] for p in allp])
            fadv = np.concatenate([p[0][2] for p in allp])
            floss, _k, _e, fg, _h = TO.policy_loss_kl_grad_fvp(params, 5, 128, 2, fo, fa, fadv, np.zeros(1026))
            np.testing.assert_allclose(g_global[:-1], fg, rtol=1e-10, atol=1e-14)
            assert abs(g_global[-1] - floss) < 1e-12
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback

        q.put((rank, "".join(traceback.format_exception(type(e), e, e.__traceback__))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_two_rank_sharding_and_reductions():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=200) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
