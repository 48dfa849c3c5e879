"""Rollouts: Steps + TakeAlignedSteps + write_experience + StepsSummary (src/simulation).

`rollout()` is one collection period of `train_parallel` (src/simulation/train.rs:108-158) with one
GPU lane per reference worker thread.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .envs import BatchedEnv
from .runtime import Context, DeviceBuffer


@dataclass
class HistoryDataBound:
    """src/agents/buffers/mod.rs:25-113"""

    min_steps: int = 0
    slack_steps: int = 0

    @staticmethod
    def with_default_slack(min_steps: int) -> "HistoryDataBound":
        return HistoryDataBound(min_steps, min(max(min_steps // 100, 5), 1000))

    def max(self, other: "HistoryDataBound") -> "HistoryDataBound":
        return HistoryDataBound(max(self.min_steps, other.min_steps), max(self.slack_steps, other.slack_steps))

    def divide(self, n: int) -> "HistoryDataBound":
        return HistoryDataBound(-(-self.min_steps // n), self.slack_steps)


@dataclass
class ActorSpec:
    """Agent::actor(mode) flattened for the rollout kernel (src/agents/mod.rs:48-114)."""

    kind: int = L.RL_ACTOR_RANDOM
    net: object = None                 # modules.Mlp
    actions: np.ndarray | None = None  # [T, E] u8, REPLAY_ACTIONS
    table: object = None               # agents.TabularQ
    exploration_rate: float = 0.0
    training: bool = True
    lanes_per_env: int = 0
    seq_net: object = None             # modules.GruLinear (recurrent CATEGORICAL_POLICY)


class Trajectory:
    """Device-resident history buffer for all lanes (the batched analogue of VecBuffer, vec.rs:15)."""

    def __init__(self, env: BatchedEnv, step_capacity: int):
        self.env, self.ctx, self._lib = env, env.ctx, env.ctx._lib
        h = C.c_void_p()
        L.check(self._lib.rl_traj_create(env.handle, step_capacity, C.byref(h)), self.ctx.handle)
        self.handle = h
        self.step_capacity = step_capacity

    def close(self):
        if self.handle and self.ctx.handle:
            self._lib.rl_traj_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def view(self) -> L.TrajView:
        v = L.TrajView()
        L.check(self._lib.rl_traj_view_of(self.handle, C.byref(v)), self.ctx.handle)
        return v

    def to_host(self) -> dict:
        """Read the whole buffer back: obs/next_obs [T, E, F], action/reward/succ [T, E], lane_len [E]."""
        v = self.view()
        T, E, F = v.step_capacity, v.num_lanes, v.num_features
        rd = self.ctx.read
        return {
            "obs": rd(v.obs, (T, F, E), np.float32).transpose(0, 2, 1).copy(),
            "next_obs": rd(v.next_obs, (T, F, E), np.float32).transpose(0, 2, 1).copy(),
            "action": rd(v.action, (T, E), np.uint8),
            "reward": rd(v.reward, (T, E), np.float32),
            "succ": rd(v.succ, (T, E), np.uint8),
            "lane_len": rd(v.lane_len, (E,), np.uint32),
            "num_steps": int(v.num_steps),
        }

    def load(self, obs, action, reward, succ, next_obs=None):
        """Fill from host arrays obs [T, E, F], action/reward/succ [T, E] (parity entry point)."""
        T, E, F = obs.shape
        bufs = [self.ctx.to_device(np.ascontiguousarray(np.asarray(obs, np.float32).transpose(0, 2, 1))),
                self.ctx.to_device(np.asarray(action, np.uint8)),
                self.ctx.to_device(np.asarray(reward, np.float32)),
                self.ctx.to_device(np.asarray(succ, np.uint8))]
        nb = None
        if next_obs is not None:
            nb = self.ctx.to_device(np.ascontiguousarray(np.asarray(next_obs, np.float32).transpose(0, 2, 1)))
        L.check(self._lib.rl_traj_load(self.handle, T, bufs[0].c, bufs[1].c, bufs[2].c, bufs[3].c,
                                       nb.c if nb else None), self.ctx.handle)
        self.ctx.synchronize()
        for b in bufs + ([nb] if nb else []):
            b.free()


def rollout(env: BatchedEnv, actor: ActorSpec, bound: HistoryDataBound, traj: Trajectory, want_summary: bool = True):
    """One period of experience for every lane.  Returns StepsSummary (or None)."""
    lib = env.ctx._lib
    cfg = L.ActorCfg()
    cfg.kind = actor.kind
    cfg.net = actor.net.handle if actor.net is not None else None
    keep = None
    if actor.actions is not None:
        if isinstance(actor.actions, DeviceBuffer):
            cfg.actions_dev = actor.actions.c
        else:
            keep = env.ctx.to_device(np.ascontiguousarray(actor.actions, np.uint8))
            cfg.actions_dev = keep.c
    cfg.table = actor.table.handle if actor.table is not None else None
    cfg.exploration_rate = actor.exploration_rate
    cfg.training = 1 if actor.training else 0
    cfg.lanes_per_env = actor.lanes_per_env
    cfg.seq_net = actor.seq_net.handle if actor.seq_net is not None else None
    summ = L.StepsSummary() if want_summary else None
    L.check(lib.rl_rollout(env.handle, C.byref(cfg), L.Bound(bound.min_steps, bound.slack_steps), traj.handle,
                           C.byref(summ) if want_summary else None), env.ctx.handle)
    if keep is not None:
        env.ctx.synchronize()
        keep.free()
    return summ
