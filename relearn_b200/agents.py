"""Agents on the hot path (src/agents, src/torch/agents): tabular Q-learning for now; the neural
agents (ActorCritic/TRPO, DQN) live in `torch_agents.py`."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .runtime import Context
from .simulation import ActorSpec, Trajectory


class TabularQ:
    """BaseTabularQLearningAgent (src/agents/tabular.rs:84-232), one table per replica (= lane)."""

    def __init__(self, ctx: Context, num_replicas: int, num_observations: int, num_actions: int,
                 discount_factor: float, exploration_rate: float = 0.2):
        self.ctx, self._lib = ctx, ctx._lib
        self.shape = (num_replicas, num_observations, num_actions)
        self.exploration_rate = exploration_rate
        h = C.c_void_p()
        L.check(self._lib.rl_tabq_create(ctx.handle, num_replicas, num_observations, num_actions, discount_factor,
                                         C.byref(h)), ctx.handle)
        self.handle = h

    def close(self):
        if self.handle and self.ctx.handle:
            self._lib.rl_tabq_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def actor(self, training: bool = True) -> ActorSpec:
        """Agent::actor(mode) (tabular.rs:148-156)."""
        return ActorSpec(kind=L.RL_ACTOR_TABULAR_EPS_GREEDY, table=self, exploration_rate=self.exploration_rate,
                         training=training)

    def update(self, traj: Trajectory):
        """BatchUpdate::batch_update (tabular.rs:197-207)."""
        L.check(self._lib.rl_tabq_update(self.handle, traj.handle), self.ctx.handle)

    batch_update = update

    def get_table(self):
        q = np.empty(self.shape, np.float64)
        c = np.empty(self.shape, np.uint64)
        L.check(self._lib.rl_tabq_get_table(self.handle, q.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)),
                self.ctx.handle)
        return q, c

    def set_table(self, q, counts):
        q = np.ascontiguousarray(q, np.float64)
        c = np.ascontiguousarray(counts, np.uint64)
        L.check(self._lib.rl_tabq_set_table(self.handle, q.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)),
                self.ctx.handle)
