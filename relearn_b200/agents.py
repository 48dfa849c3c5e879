"""Finite-space agents (src/agents): tabular Q-learning and UCB1; the neural agents (ActorCritic / TRPO / PPO, DQN) live
in `torch_agents.py`."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .runtime import Context
from .simulation import ActorSpec, Trajectory


class TabularQ:
    """BaseTabularQLearningAgent (src/agents/tabular.rs:84-232).

    `num_replicas = 1` over an env with E > 1 lanes is the reference under `train_parallel`: every lane (worker) acts
    from the one shared table, frozen while a period runs, and `batch_update` folds the lanes' buffers into it one after
    the other in lane order (tabular.rs:197-207) -- bit-identical to the reference's sequential fold.
    `num_replicas = E` trains E INDEPENDENT agents, one table per lane, each seeing only its own lane's data: replicas
    for sweeps, not a faster way to train one agent (sample efficiency per table is that of a single worker)."""

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

    def update(self, traj: Trajectory, logger=None):
        """BatchUpdate::batch_update (tabular.rs:197-207)."""
        L.check(self._lib.rl_tabq_update(self.handle, traj.handle), self.ctx.handle)

    batch_update = update

    def min_update_size(self):
        """tabular.rs:190-195: one step (the training loop's min_worker_steps sets the period)."""
        from .simulation import HistoryDataBound

        return HistoryDataBound(1, 0)

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


@dataclass
class UCB1AgentConfig:
    """ucb.rs:20-40: `exploration_rate` scales the confidence interval (0.2 after Audibert and Munos)."""

    exploration_rate: float = 0.2

    def build_agent(self, env, num_replicas: int = 1) -> "UCB1Agent":
        """BuildAgent::build_agent (ucb.rs:42-76): tables sized by the env's finite observation / action spaces, rewards
        scaled by its feedback range.  `num_replicas`: 1 = one agent shared by every lane (train_parallel), env.num_envs =
        one independent agent per lane."""
        st = env.structure
        if st.num_observations <= 0:
            raise ValueError("UCB1 needs a finite observation space")
        return UCB1Agent(env.ctx, num_replicas, st.num_observations, st.num_actions, (st.reward_lo, st.reward_hi),
                         self.exploration_rate)


class UCB1Agent:
    """BaseUCB1Agent (src/agents/bandits/ucb.rs:78-243): UCB1 (Auer 2002) applied independently to each state."""

    def __init__(self, ctx: Context, num_replicas: int, num_observations: int, num_actions: int, reward_range,
                 exploration_rate: float = 0.2):
        self.ctx, self._lib = ctx, ctx._lib
        self.shape = (num_replicas, num_observations, num_actions)
        self.exploration_rate = exploration_rate
        h = C.c_void_p()
        L.check(self._lib.rl_ucb1_create(ctx.handle, num_replicas, num_observations, num_actions, float(reward_range[0]),
                                         float(reward_range[1]), exploration_rate, C.byref(h)), ctx.handle)
        self.handle = h

    def close(self):
        if self.handle and self.ctx.handle:
            self._lib.rl_ucb1_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def actor(self, training: bool = True) -> ActorSpec:
        """Agent::actor(mode) (ucb.rs:162-171): Training maximises the upper confidence bound, Evaluation the count."""
        return ActorSpec(kind=L.RL_ACTOR_UCB1, ucb=self, training=training)

    def update(self, traj: Trajectory, logger=None):
        """BatchUpdate::batch_update (ucb.rs:186-199)."""
        L.check(self._lib.rl_ucb1_update(self.handle, traj.handle), self.ctx.handle)

    batch_update = update

    def min_update_size(self):
        """ucb.rs:179-184."""
        from .simulation import HistoryDataBound

        return HistoryDataBound(1, 0)

    def get_tables(self):
        """(state_action_mean_reward f64 [R, S, A], state_action_count u64 [R, S, A], state_visit_count u64 [R, S])."""
        m = np.empty(self.shape, np.float64)
        c = np.empty(self.shape, np.uint64)
        v = np.empty(self.shape[:2], np.uint64)
        L.check(self._lib.rl_ucb1_get_tables(self.handle, m.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                                             v.ctypes.data_as(C.c_void_p)), self.ctx.handle)
        return m, c, v

    def set_tables(self, mean, count, visits):
        m = np.ascontiguousarray(mean, np.float64)
        c = np.ascontiguousarray(count, np.uint64)
        v = np.ascontiguousarray(visits, np.uint64)
        L.check(self._lib.rl_ucb1_set_tables(self.handle, m.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                                             v.ctypes.data_as(C.c_void_p)), self.ctx.handle)
