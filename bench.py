#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused rollout (+ policy act) hot path, TRPO update ms, roofline.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`; for N > 1 the driver
launches it under torchrun (one rank per GPU, RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* in the env).  One JSON
line on stdout from rank 0.

Workload = BASELINE.json configs[1] "cartpole-trpo": CartPole + VisibleStepLimit(500), MLP policy 5->128->2,
E = 4096 lanes per GPU, T = 256 steps per lane per period (N = 1 048 576 env-steps per GPU per step of this
bench).  A "step" is one collection period: fresh episodes on every lane, policy forward + categorical
sample + dynamics + trajectory write for E*T env-steps.  Weak scaling: every rank owns its own E lanes
(global lane ids rank*E ..), no data-path collective.

`--impl reference` times the reference's CPU implementation of the same path: the C oracle port of
Steps::step + PolicyActor::act on all host cores (the Rust reference cannot be built in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
B_PER_STEP_ROLLOUT = 26      # obs 5 x f32 + action u8 + reward f32 + succ u8
B_PER_STEP_UNFUSED = 98      # K1 (f64 state)
B_PER_STEP_SCAN = 17         # K3
FLOP_PER_STEP_POLICY = 1792  # 2 * (5*128 + 128*2)
# K6 critic pass: forward 1536 + backward 3072 FLOP per sample per Adam iteration (the network's own
# arithmetic).  What the tcgen05 kernel issues per 128-sample tile is larger (bf16-piece operands, padded N):
# MMA1 128x128x48, MMA3 128x32x128, MMA2 128x32x128  =>  2 * (786432 + 524288 + 524288) / 128 = 28672 FLOP per sample.
FLOP_PER_SAMPLE_CRITIC = 4608
FLOP_ISSUED_PER_SAMPLE_CRITIC_TC = 28672
# dram__bytes_read.sum + dram__bytes_write.sum of one K2w launch (the kernel `lanes_per_env = 0` picks) at E = 4096,
# T = 256 from the round-1 `ncu --set full` capture of the final kernel (profiles/r1i_k2w_e4096.md): 50.7 KB + 17.4 KB
# (the write figure is whatever part of the trajectory L2 happened to write back during the launch: 2.8 .. 22 KB over captures).  The 27 MB trajectory of a
# period stays in the 126 MB L2 while the kernel runs and drains afterwards, so the in-kernel DRAM traffic
# is far BELOW the algorithmic 26 B/env-step, not above it.
K2C_NCU_DRAM_BYTES_PER_LAUNCH = 68096


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """Dense bf16 TFLOP/s: the sustained figure (the critic pass runs 80 times back to back inside a long step)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), "measured sustained (MEASURED_PEAKS.json)"
    return 1500.0, "fallback (B200_PROFILING.md)"


def cpu_update_baseline(n_steps: int, seed: int = 0):
    """The reference's update on the host: torch-CPU restatement of Trpo::update + ValuesOpt::update (oracle/
    tensor_oracle.py, all host threads) on a bounded sample of the same batch distribution."""
    import torch

    from oracle import tensor_oracle as TO
    from relearn_b200.modules import init_params

    rng = np.random.default_rng(seed)
    obs = rng.uniform(-1, 1, size=(n_steps, 5)).astype(np.float32)
    act = rng.integers(0, 2, n_steps)
    adv = rng.normal(size=n_steps).astype(np.float32)
    tgt = rng.uniform(0, 50, size=n_steps).astype(np.float32)
    pp, vp = init_params(rng, 5, 128, 2), init_params(rng, 5, 128, 1)
    t0 = time.perf_counter()
    TO.trpo_update(pp, 5, 128, 2, obs, act, adv)
    t1 = time.perf_counter()
    TO.value_update(vp, 5, 128, obs, tgt, n_steps=80)
    t2 = time.perf_counter()
    return {"batch_steps": n_steps, "trpo_policy_ms": (t1 - t0) * 1e3, "critic_80_adam_ms": (t2 - t1) * 1e3,
            "threads": torch.get_num_threads(), "kind": "port (torch-CPU autograd restatement, f32)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax.append(float(p[2]))
                except ValueError:
                    continue
                for n, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        return out


def dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist  # noqa: F811

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def max_over_ranks(dist, local, value: float) -> float:
    if dist is None:
        return value
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist, ctx):
    ctx.synchronize()
    if dist is not None:
        import torch

        dist.barrier()
        torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------
# CPU arm (the reference's own CPU path, as the C oracle port)
# ------------------------------------------------------------------------------------------------
def cpu_rollout_rate(params: np.ndarray, lanes: int, horizon: int, threads: int, seed: int = 1, t0: int = 0):
    import ctypes as C

    import oracle as O

    cfg = O.cartpole_cfg(500)
    mlp = O.mlp_struct(params, 5, 128, 2) if params is not None else None  # None: RandomAgent (env only)
    summ = O.Summary()
    t = time.perf_counter()
    O.lib().ro_rollout_lanes_philox(C.byref(cfg), C.byref(mlp) if mlp is not None else None, lanes, 0, horizon, 0, seed, t0,
                                    threads, C.byref(summ))
    dt = time.perf_counter() - t
    return lanes * horizon / dt, dt, summ


def run_reference(args):
    rank, world, local, _ = dist_setup() if int(os.environ.get("WORLD_SIZE", "1")) > 1 else (0, 1, 0, None)
    if rank != 0:
        return
    import oracle as O  # noqa: F401
    from relearn_b200.modules import init_params

    cores = os.cpu_count() or 1
    params = init_params(np.random.default_rng(0), 5, 128, 2)
    horizon = args.horizon
    # bounded sample: calibrate so that one step is ~1 s of wall time on all cores
    rate, _, _ = cpu_rollout_rate(params, max(cores, 8), horizon, cores)
    lanes = int(max(cores, min(args.envs, rate * 1.0 / horizon)))
    for _ in range(args.warmup):
        cpu_rollout_rate(params, lanes, horizon, cores)
    t = time.perf_counter()
    for i in range(args.steps):
        cpu_rollout_rate(params, lanes, horizon, cores, t0=i * (horizon + 1))
    dt = time.perf_counter() - t
    value = lanes * horizon * args.steps / dt
    sample = f"{lanes} lanes x {horizon} steps per step on {cores} threads (C oracle port of Steps::step + PolicyActor::act)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 dynamics / f32 policy", "data": "synthetic",
        "config": {"workload": "cartpole-trpo rollout (CartPole+VisibleStepLimit(500), MLP 5-128-2 policy)",
                   "envs_per_step": lanes, "horizon": horizon, "host": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import relearn_b200 as R
    from relearn_b200 import _lib as L

    rank, world, local, dist = dist_setup()
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    ctx = R.Context(local)
    info = ctx.device_info()
    E, T = args.envs, args.horizon
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    env = R.build_env(ctx, cfg, E, seed=1234, lane_offset=rank * E)
    rng = np.random.default_rng(0)
    params = R.init_params(rng, 5, 128, 2)
    vparams = R.init_params(rng, 5, 128, 1)
    agent = R.ActorCriticConfig().build_agent(env)
    agent.policy.policy_fn.set_weights(params)
    agent.critic.state_value_fn.set_weights(vparams)
    traj = R.Trajectory(env, T)
    bound = R.HistoryDataBound(T, 0)
    actor = agent.actor(args.lanes)
    flush = ctx.alloc(256 << 20)  # > 126 MB L2

    if world > 1:  # data-parallel group for the update's all-reduces
        ids = [R.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)

    def one_period(timed_events=None):
        flush.zero()
        if timed_events is not None:
            timed_events[0].record()
        R.rollout(env, actor, bound, traj, want_summary=False)
        if timed_events is not None:
            timed_events[1].record()

    for _ in range(max(args.warmup, 3)):
        one_period()
    barrier(dist, ctx)
    clocks = ClockSampler(local)
    clocks.start()
    evs = [(ctx.event(), ctx.event()) for _ in range(args.steps)]
    launches0 = ctx.launch_count
    barrier(dist, ctx)
    wall0 = time.perf_counter()
    for k in range(args.steps):
        one_period(evs[k])
    barrier(dist, ctx)
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count - launches0
    kernel_ms = sum(a.elapsed_ms(b) for a, b in evs)
    total_ms = max_over_ranks(dist, local, kernel_ms)
    clock_info = clocks.stop()
    steps_per_period = E * T
    value = world * steps_per_period * args.steps / (total_ms * 1e-3)

    # ---- e2e through the C ABI with host buffers: weights H2D (pinned) -> rollout -> summary D2H ----
    w_pinned = ctx.pinned_array((params.size,), np.float32)
    w_pinned[:] = params
    e2e_s = 0.0
    for k in range(max(3, min(args.warmup, 5)) + args.steps):
        flush.zero()
        ctx.synchronize()
        t0 = time.perf_counter()
        agent.policy.policy_fn.set_weights_async(w_pinned)
        summ = R.rollout(env, actor, bound, traj, want_summary=True)
        dt = time.perf_counter() - t0
        if k >= max(3, min(args.warmup, 5)):
            e2e_s += dt
    e2e_s = max_over_ranks(dist, local, e2e_s)
    e2e_value = world * steps_per_period * args.steps / e2e_s
    h2d = int(params.nbytes)
    d2h = 10 * 8

    # ---- roofline of the dominant kernel (fused rollout: trajectory write stream) ----
    hbm_peak, peak_src = measured_peaks()
    achieved = B_PER_STEP_ROLLOUT * steps_per_period * args.steps / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": K2C_NCU_DRAM_BYTES_PER_LAUNCH if (E == 4096 and T == 256) else None,
                "traffic_note": "ncu dram bytes of one launch (profiles/r1_summary.md s13); trajectory is L2-resident during the kernel",
                "peak_source": peak_src, "kernel": "rollout_cartpole_ws_kernel K2w (fused step+policy+sample; policy and dynamics on different warps)",
                "note": "K2w moves 26 B of trajectory per env-step, so HBM is not its bound: at 4096 envs the period is 256 x the latency of the dynamics warp's dependent f64 step (~1200 clk, DESIGN section 4); the HBM-bound kernels of the path are in kernels[] (77-81 % of peak)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 dynamics / f32 policy", "data": "synthetic",
        "config": {"workload": "cartpole-trpo rollout (CartPole+VisibleStepLimit(500), MLP 5-128-2 policy, fused)",
                   "envs_per_gpu": E, "horizon": T, "env_steps_per_step": world * steps_per_period,
                   "lanes_per_env": args.lanes, "l2": "flushed between timed iterations (256 MiB memset)",
                   "noise": "philox4x32-10", "parallelism": f"dp{world} (lanes sharded, no collective)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "rl_mlp_set_weights_async(pinned host) + rl_rollout + StepsSummary read-back per period"},
        "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline,
        "wall_s_timed_region": wall, "sm_count": info["sm_count"],
    }

    if rank == 0 and not args.quick:
        fp32_peak = ctx.fp32_peak_tflops()
        line["fp32"] = {"policy_tflops": FLOP_PER_STEP_POLICY * steps_per_period * args.steps / (kernel_ms * 1e-3) / 1e12,
                        "measured_fma_peak_tflops": fp32_peak}
    # ---- the update that consumes the rollout: GAE + TRPO + critic (ms per iteration) ----
    if not args.no_update:
        upd = {"adv_est_ms": [], "policy_ms": [], "critic_ms": [], "status": []}
        for it in range(1 + args.update_iters):
            R.rollout(env, actor, bound, traj, want_summary=False)
            barrier(dist, ctx)
            e0, e1 = ctx.event().record(), None
            adv = agent.critic.advantages(traj)
            e1 = ctx.event().record()
            log = {}
            status = agent.policy.update(traj, adv, log)
            cstats = agent.critic.update(traj, log)
            if it > 0:
                upd["adv_est_ms"].append(e0.elapsed_ms(e1))
                upd["policy_ms"].append(log["policy/update_time"] * 1e3)
                upd["critic_ms"].append(cstats.update_ms)
                upd["status"].append(int(status))
        adv_ms = float(np.mean(upd["adv_est_ms"]))
        pol_ms = max_over_ranks(dist, local, float(np.mean(upd["policy_ms"])))
        cri_ms = max_over_ranks(dist, local, float(np.mean(upd["critic_ms"])))
        tpeak, tpeak_src = measured_tensor_peak()
        fp32_peak = line.get("fp32", {}).get("measured_fma_peak_tflops")
        alg = FLOP_PER_SAMPLE_CRITIC * steps_per_period * 80 / (cri_ms * 1e-3) / 1e12
        issued = FLOP_ISSUED_PER_SAMPLE_CRITIC_TC * steps_per_period * 80 / (cri_ms * 1e-3) / 1e12
        line["update"] = {
            "batch_steps_per_gpu": steps_per_period, "adv_est_ms": adv_ms, "trpo_policy_ms": pol_ms,
            "critic_80_adam_ms": cri_ms, "total_ms": adv_ms + pol_ms + cri_ms, "status": upd["status"],
            "pass_kernel": os.environ.get("RL_PASS_KERNEL", "tcgen05"),
            "all_reduce": ("none (1 GPU)" if world == 1 else
                           "fused row-reduction + NVLink peer-mailbox exchange (+ Adam), one kernel per pass"
                           if ctx.comm_peer_info()["peer_mailboxes"] else "reduce + ncclAllReduce (f64 sums)"),
            # the critic's 80 passes + Adam steps as one region: the network's own FLOPs per second (compare with the
            # FP32 FMA peak, which bounds any non-tensor implementation) and the bf16 FLOPs the MMAs issue
            "critic_roofline": {"bound": "tensor", "achieved": issued, "peak": tpeak, "unit": "TFLOP/s", "frac": issued / tpeak,
                                "peak_source": tpeak_src, "algorithmic_tflops": alg, "fp32_fma_peak_tflops": fp32_peak,
                                "note": "achieved = bf16 FLOPs issued by the three MMAs per tile (pieces + padding, 28 672 per "
                                        "sample); algorithmic_tflops = 4608 FLOP per sample per iteration of the f32 network"},
        }
    # ---- per-kernel rooflines for the HBM-bound kernels (rank 0, 1 GPU only) ----
    if rank == 0 and world == 1 and not args.quick:
        line["kernels"] = kernel_rooflines(ctx, R, L, hbm_peak)
        # CPU baseline on this box's host cores (bounded sample)
        cores = os.cpu_count() or 1
        cpu_rollout_rate(params, max(cores, 8), T, cores)
        periods, cpu_s = 0, 0.0
        while cpu_s < 10.0 and periods < 10000:  # ~10 s of wall time on all cores
            _, dt, _ = cpu_rollout_rate(params, E, T, cores, t0=periods * (T + 1))
            cpu_s += dt
            periods += 1
        line["cpu_baseline"] = {"value": periods * E * T / cpu_s, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{periods} periods of {E} lanes x {T} steps ({cpu_s:.1f} s wall) of the same "
                                          f"workload, C oracle port of Steps::step + PolicyActor::act, {cores} threads"}
        # env-only (RandomAgent: no policy network), SURVEY 8d's second CPU number: ~3 s of the same lanes on all cores
        env_s, env_periods = 0.0, 0
        while env_s < 3.0:
            _, dt, _ = cpu_rollout_rate(None, E, T, cores, t0=env_periods * (T + 1))
            env_s += dt
            env_periods += 1
        line["cpu_baseline"]["env_only"] = {"value": env_periods * E * T / env_s, "unit": UNIT,
                                            "sample": f"{env_periods} periods of {E} lanes x {T} steps, RandomAgent, {cores} threads"}
        if not args.no_update:
            try:
                line["update"]["cpu_baseline"] = cpu_update_baseline(1 << 16)
            except Exception as exc:  # torch missing on the box: report, do not fail the bench
                line["update"]["cpu_baseline"] = {"unavailable": repr(exc)[:200]}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def kernel_rooflines(ctx, R, L, hbm_peak):
    """Unfused step kernel (K1) and GAE scan (K3) at sizes larger than L2, timed alone with CUDA events."""
    import ctypes as C

    out = []
    E = 1 << 22
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    env = R.build_env(ctx, cfg, E, seed=7)
    env.reset_all()
    actions = ctx.to_device(np.random.default_rng(0).integers(0, 2, E).astype(np.uint8))
    for _ in range(5):
        env.step_device(actions)
    reps = 20
    e0 = ctx.event().record()
    for _ in range(reps):
        env.step_device(actions)
    e1 = ctx.event().record()
    ms = e0.elapsed_ms(e1) / reps
    gbs = B_PER_STEP_UNFUSED * E / (ms * 1e-3) / 1e9
    out.append({"kernel": "env_step_kernel<CartPole> (unfused, f64 SoA state)", "envs": E, "bytes_per_step": B_PER_STEP_UNFUSED,
                "ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak, "env_steps_per_s": E / (ms * 1e-3),
                "working_set_mb": E * (B_PER_STEP_UNFUSED + 20) / 1e6})
    env.close()
    # GAE / reward-to-go scan over [T, E] = [256, 131072] (33.5 M steps, 570 MB)
    T, E2 = 256, 1 << 17
    env2 = R.build_env(ctx, cfg, E2, seed=8)
    traj = R.Trajectory(env2, T)
    net = R.Mlp(ctx, 5, [128], 2)
    net.set_weights(R.init_params(np.random.default_rng(1), 5, 128, 2))
    R.rollout(env2, R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net), R.HistoryDataBound(T, 0), traj, want_summary=False)
    adv, rtg = ctx.alloc(T * E2 * 4), ctx.alloc(T * E2 * 4)
    x = ctx.alloc(T * E2 * 4)
    x.zero()
    v = traj.view()
    lib = ctx._lib
    for _ in range(3):
        L.check(lib.rl_gae(traj.handle, None, 0.99, 0.95, adv.c, rtg.c), ctx.handle)
    reps = 10
    e0 = ctx.event().record()
    for _ in range(reps):
        L.check(lib.rl_gae(traj.handle, None, 0.99, 0.95, adv.c, rtg.c), ctx.handle)
    e1 = ctx.event().record()
    ms = e0.elapsed_ms(e1) / reps
    # without a critic the pass reads reward 4 + succ 1 and writes adv 4 + rtg 4 = 13 B/step
    gbs = 13 * T * E2 / (ms * 1e-3) / 1e9
    out.append({"kernel": "gae_scan_kernel<no critic> (reward-to-go + advantages)", "steps": T * E2, "bytes_per_step": 13,
                "ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak})
    for _ in range(3):
        L.check(lib.rl_discounted_cumsum(ctx.handle, x.c, C.c_void_p(v.succ), T, E2, 0.99, adv.c), ctx.handle)
    e0 = ctx.event().record()
    for _ in range(reps):
        L.check(lib.rl_discounted_cumsum(ctx.handle, x.c, C.c_void_p(v.succ), T, E2, 0.99, adv.c), ctx.handle)
    e1 = ctx.event().record()
    ms = e0.elapsed_ms(e1) / reps
    gbs = 9 * T * E2 / (ms * 1e-3) / 1e9
    out.append({"kernel": "cumsum_kernel (discounted_cumsum_from_end)", "steps": T * E2, "bytes_per_step": 9, "ms": ms,
                "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak})
    # large-E fused rollout (config 5's throughput end): K2c with one thread per env on the FP32 pipe, and K2t with the
    # policy's hidden layer on tcgen05 (what `lanes_per_env = 0` picks at this size)
    E3, T3 = 1 << 20, 64
    env3 = R.build_env(ctx, cfg, E3, seed=9)
    traj3 = R.Trajectory(env3, T3)
    for lanes, name in ((1, "rollout_cartpole_group_kernel<1> K2c (fused, thread per env, FP32 pipe)"),
                        (L.RL_LANES_TENSOR_CORE, "rollout_cartpole_tc_kernel K2t (fused, 128-env tiles, hidden layer on tcgen05)")):
        spec = R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, net=net, lanes_per_env=lanes)
        for _ in range(2):
            R.rollout(env3, spec, R.HistoryDataBound(T3, 0), traj3, want_summary=False)
        reps = 5
        e0 = ctx.event().record()
        for _ in range(reps):
            R.rollout(env3, spec, R.HistoryDataBound(T3, 0), traj3, want_summary=False)
        e1 = ctx.event().record()
        ms = e0.elapsed_ms(e1) / reps
        rec = {"kernel": name, "envs": E3, "horizon": T3, "ms": ms, "env_steps_per_s": E3 * T3 / (ms * 1e-3),
               "trajectory_write_gbs": B_PER_STEP_ROLLOUT * E3 * T3 / (ms * 1e-3) / 1e9,
               "policy_tflops": FLOP_PER_STEP_POLICY * E3 * T3 / (ms * 1e-3) / 1e12}
        if lanes == L.RL_LANES_TENSOR_CORE:
            # bf16 FLOPs the MMAs issue: [128 x 48] . [48 x 128] per 128 env-steps
            rec["mma_bf16_tflops"] = 2 * 48 * 128 * E3 * T3 / (ms * 1e-3) / 1e12
        out.append(rec)
    out += other_configs(ctx, R, L, hbm_peak)
    return out


def other_configs(ctx, R, L, hbm_peak):
    """BASELINE configs 3 and 4 at their per-GPU sizes (timed alone, CUDA events): DQN collection + replay append +
    update at 65 536 envs, and the bandit meta-env with the rnn.rs-sized GRU policy at 131 072 envs."""
    out = []
    rng = np.random.default_rng(3)
    # ---- config 3: cartpole-dqn, 65 536 envs ----
    E, cfg = 65536, R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    env = R.build_env(ctx, cfg, E, seed=5)
    agent = R.DqnConfig(buffer_capacity=762).build_agent(env)  # 50 M steps of replay per GPU (examples/cartpole-dqn.rs:34-42)
    agent.action_value_fn.set_weights(R.init_params(rng, 5, 128, 2))
    rb = agent.buffer()
    bound = R.HistoryDataBound(16, 5)  # first update: 1 M steps over 65 536 lanes
    traj = R.Trajectory(env, bound.min_steps + bound.slack_steps)
    times = {"rollout_ms": [], "append_ms": [], "update_ms": []}
    for it in range(4):
        e0 = ctx.event().record()
        R.rollout(env, agent.actor(), bound, traj, want_summary=False)
        e1 = ctx.event().record()
        rb.write_experience(traj)
        e2 = ctx.event().record()
        stats = agent.batch_update(rb)
        if it > 0:
            times["rollout_ms"].append(e0.elapsed_ms(e1))
            times["append_ms"].append(e1.elapsed_ms(e2))
            times["update_ms"].append(stats.update_ms)
    st = rb.stats()
    steps = traj.view().num_steps
    app_ms = float(np.mean(times["append_ms"]))
    out.append({"kernel": "config 3 cartpole-dqn: eps-greedy rollout + replay append + 50 x (sample 100k, Q loss, Adam)",
                "envs": E, "steps_per_period": int(steps), "rollout_ms": float(np.mean(times["rollout_ms"])),
                "append_ms": app_ms, "append_gbs": 2 * 26 * steps / (app_ms * 1e-3) / 1e9,
                "dqn_update_50_steps_ms": float(np.mean(times["update_ms"])), "replay_steps": int(st.num_steps),
                "replay_episodes": int(st.num_episodes)})
    traj.close(); rb.close(); env.close()
    # ---- config 4: bandit meta-env (k = 2 arms, n = 10 episodes per trial) + GRU(6 -> 4) -> Linear(4 -> 2) ----
    E4, trials = 131072, 10
    mcfg = R.MetaEnv(R.UniformBernoulliBandits(2), 10)
    env4 = R.build_env(ctx, mcfg, E4, seed=6)
    net = R.GruLinear(ctx, env4.num_features, 4, env4.num_actions)
    net.set_weights(R.init_gru_linear_params(rng, env4.num_features, 4, env4.num_actions))
    T4 = trials * 19
    traj4 = R.Trajectory(env4, T4)
    spec = R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, seq_net=net)
    for _ in range(2):
        R.rollout(env4, spec, R.HistoryDataBound(T4, 0), traj4, want_summary=False)
    reps = 5
    e0 = ctx.event().record()
    for _ in range(reps):
        R.rollout(env4, spec, R.HistoryDataBound(T4, 0), traj4, want_summary=False)
    e1 = ctx.event().record()
    ms = e0.elapsed_ms(e1) / reps
    bytes_per_step = 4 * env4.num_features + 6  # obs planes + action + reward + succ
    out.append({"kernel": "config 4 bandit meta-env + GRU(6->4)->Linear policy, fused rollout (thread per env)", "envs": E4,
                "horizon": T4, "ms": ms, "env_steps_per_s": E4 * T4 / (ms * 1e-3), "bytes_per_step": bytes_per_step,
                "trajectory_write_gbs": bytes_per_step * E4 * T4 / (ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": bytes_per_step * E4 * T4 / (ms * 1e-3) / 1e9 / hbm_peak})
    traj4.close()
    # the update that consumes it: GAE + TRPO + 80-step critic through Chain<Gru, Linear> on one trial per lane (T = 19)
    g = R.GruLinearConfig(hidden_dim=4)
    agent4 = R.ActorCriticConfig(policy_config=R.TrpoConfig(policy_fn_config=g),
                                 critic_config=R.ValuesOptConfig(state_value_fn_config=g)).build_agent(env4)
    agent4.policy.policy_fn.set_weights(R.init_gru_linear_params(rng, env4.num_features, 4, env4.num_actions))
    agent4.critic.state_value_fn.set_weights(R.init_gru_linear_params(rng, env4.num_features, 4, 1))
    traj5 = R.Trajectory(env4, 19)
    upd = []
    for it in range(3):
        R.rollout(env4, agent4.actor(), R.HistoryDataBound(19, 0), traj5, want_summary=False)
        e0 = ctx.event().record()
        adv = agent4.critic.advantages(traj5)
        e1 = ctx.event().record()
        log = {}
        agent4.policy.update(traj5, adv, log)
        cs = agent4.critic.update(traj5, log)
        if it > 0:
            upd.append((e0.elapsed_ms(e1), log["policy/update_time"] * 1e3, cs.update_ms))
    out.append({"kernel": "config 4 update: GAE + TRPO + 80-step critic through GRU(6->4)->Linear (BPTT pass kernel)", "envs": E4,
                "batch_steps": E4 * 19, "adv_est_ms": float(np.mean([u[0] for u in upd])),
                "trpo_policy_ms": float(np.mean([u[1] for u in upd])), "critic_80_adam_ms": float(np.mean([u[2] for u in upd]))})
    traj5.close(); env4.close()
    # ---- config 4, rl2-sized: k = 10 arms, n = 100 episodes per trial, GRU(14 -> 128) -> Linear(128 -> 10) (K8h) ----
    E6, T6 = 148 * 64 * 2, 199
    env6 = R.build_env(ctx, R.MetaEnv(R.UniformBernoulliBandits(10), 100), E6, seed=7)
    net6 = R.GruLinear(ctx, env6.num_features, 128, env6.num_actions)
    net6.set_weights(R.init_gru_linear_params(rng, env6.num_features, 128, env6.num_actions))
    traj6 = R.Trajectory(env6, T6)
    spec6 = R.ActorSpec(kind=L.RL_ACTOR_CATEGORICAL_POLICY, seq_net=net6)
    R.rollout(env6, spec6, R.HistoryDataBound(T6, 0), traj6, want_summary=False)
    e0 = ctx.event().record()
    for _ in range(3):
        R.rollout(env6, spec6, R.HistoryDataBound(T6, 0), traj6, want_summary=False)
    e1 = ctx.event().record()
    ms = e0.elapsed_ms(e1) / 3
    flop = 2 * 3 * 128 * (env6.num_features + 128) + 2 * 128 * env6.num_actions
    fp32_peak = ctx.fp32_peak_tflops()
    out.append({"kernel": "config 4 rl2-sized: bandit meta-env (10 arms x 100 episodes) + GRU(14->128)->Linear(128->10), fused rollout "
                          "K8h (64-env tiles, FP32 FFMA2 GEMM per step, weights streamed by cp.async.bulk)",
                "envs": E6, "horizon": T6, "ms": ms, "env_steps_per_s": E6 * T6 / (ms * 1e-3), "flop_per_env_step": flop,
                "fp32_tflops": flop * E6 * T6 / (ms * 1e-3) / 1e12, "measured_fma_peak_tflops": fp32_peak,
                "frac_of_fp32_peak": flop * E6 * T6 / (ms * 1e-3) / 1e12 / fp32_peak})
    traj6.close(); env6.close()
    return out


_REAL_STDOUT = None


def claim_stdout():
    """Route everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr; the one JSON line
    of the contract goes to the original stdout through emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=4096, help="lanes per GPU")
    ap.add_argument("--horizon", type=int, default=256, help="steps per lane per period")
    ap.add_argument("--lanes", type=int, default=0, help="threads per env in the fused kernel (0 = auto)")
    ap.add_argument("--update-iters", type=int, default=3)
    ap.add_argument("--no-update", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the per-kernel rooflines and the CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
