"""Environment configs and batched environment handles.

Mirrors the reference's `src/envs` surface for the environments on the hot path: the config
structs keep the reference's names/defaults (`CartPoleConfig`, `Chain`, `MemoryGame`,
`MetaEnv(UniformBernoulliBandits)` + `TrialEpisodeLimit`, `VisibleStepLimit`) and `build_env`
(src/envs/builders.rs:17) returns a `BatchedEnv`: `num_envs` independent instances stepped in
lockstep on the GPU.  `EnvStructure` (src/envs/mod.rs:165-193) is exposed as `.structure`.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from .runtime import Context, DeviceBuffer


class Successor:
    """src/envs/mod.rs:257-269"""

    CONTINUE, TERMINATE, INTERRUPT, PAD = L.RL_CONTINUE, L.RL_TERMINATE, L.RL_INTERRUPT, L.RL_PAD


@dataclass
class VisibleStepLimit:
    """src/envs/wrappers/step_limit.rs:97-123"""

    max_steps_per_episode: int = 100


@dataclass
class LatentStepLimit:
    """src/envs/wrappers/step_limit.rs:13-90: episodes are interrupted after the limit, the observation does not show it"""

    max_steps_per_episode: int = 100


@dataclass
class CartPoleConfig:
    """PhysicalConstants + EnvironmentParams (src/envs/cartpole.rs:157-216); `.wrap(VisibleStepLimit(n))` or
    `.wrap(LatentStepLimit(n))`."""

    gravity: float = 9.8
    mass_cart: float = 1.0
    mass_pole: float = 0.1
    length_half_pole: float = 0.5
    friction_cart: float = 0.01
    friction_pole: float = 0.01
    time_step: float = 0.02
    action_force: float = 10.0
    max_pos: float = 2.4
    max_angle: float = 12.0 * (math.pi / 180.0)
    discount_factor: float = 0.99
    max_steps_per_episode: int = 0
    step_limit_visible: int = 1

    kind = L.RL_ENV_CARTPOLE

    def wrap(self, limit) -> "CartPoleConfig":
        out = CartPoleConfig(**{k: getattr(self, k) for k in self.__dataclass_fields__})
        out.max_steps_per_episode = limit.max_steps_per_episode
        out.step_limit_visible = 0 if isinstance(limit, LatentStepLimit) else 1
        return out

    def c_cfg(self):
        c = L.CartPoleCfg()
        for k in self.__dataclass_fields__:
            setattr(c, k, getattr(self, k))
        return c


CartPole = CartPoleConfig


@dataclass
class Chain:
    """src/envs/chain.rs:21-45"""

    size: int = 5
    discount_factor: float = 0.95
    kind = L.RL_ENV_CHAIN

    def c_cfg(self):
        return L.ChainCfg(self.size, self.discount_factor)


@dataclass
class MemoryGame:
    """src/envs/memory.rs:24-55"""

    num_actions: int = 2
    history_len: int = 1
    kind = L.RL_ENV_MEMORY_GAME

    def c_cfg(self):
        return L.MemoryCfg(self.num_actions, self.history_len)


@dataclass
class PartitionGame:
    """src/envs/partition.rs: classify 10-bit elements the way a hidden axis-aligned supervisor does (no parameters)."""

    kind = L.RL_ENV_PARTITION_GAME

    def c_cfg(self):
        return C.c_uint64(0)  # rl_env_create ignores cfg for this kind


@dataclass
class UniformBernoulliBandits:
    """src/envs/bandits.rs:128-181"""

    num_arms: int = 2


@dataclass
class OneHotBandits:
    """src/envs/bandits.rs:187-243: deterministic bandits, one uniformly chosen arm pays 1"""

    num_arms: int = 2


@dataclass
class TrialEpisodeLimit:
    """src/envs/meta.rs:541-566"""

    episodes_per_trial: int = 10


@dataclass
class MetaEnv:
    """MetaEnv<UniformBernoulliBandits | OneHotBandits>.wrap(TrialEpisodeLimit) (src/envs/meta.rs:49-203,568-617)."""

    env_distribution: UniformBernoulliBandits = field(default_factory=UniformBernoulliBandits)
    episodes_per_trial: int = 10
    kind = L.RL_ENV_BANDIT_META

    def wrap(self, limit: TrialEpisodeLimit) -> "MetaEnv":
        return MetaEnv(self.env_distribution, limit.episodes_per_trial)

    def c_cfg(self):
        return L.BanditMetaCfg(self.env_distribution.num_arms, self.episodes_per_trial,
                               1 if isinstance(self.env_distribution, OneHotBandits) else 0)


class BatchedEnv:
    """`num_envs` lanes of one environment on one GPU; global lane ids start at `lane_offset`."""

    def __init__(self, ctx: Context, config, num_envs: int, seed: int = 0, lane_offset: int = 0):
        self.ctx, self.config, self.num_envs, self.lane_offset = ctx, config, num_envs, lane_offset
        self._lib = ctx._lib
        cfg = config.c_cfg()
        h = C.c_void_p()
        L.check(self._lib.rl_env_create(ctx.handle, config.kind, C.byref(cfg), num_envs, lane_offset, seed,
                                        C.byref(h)), ctx.handle)
        self.handle = h
        s = L.EnvStructure()
        L.check(self._lib.rl_env_structure_of(self.handle, C.byref(s)), ctx.handle)
        self.structure = s
        self._noise_bufs = None

    # EnvStructure
    @property
    def num_features(self):
        return self.structure.num_features

    @property
    def num_actions(self):
        return self.structure.num_actions

    @property
    def discount_factor(self):
        return self.structure.discount_factor

    def close(self):
        if self.handle and self.ctx.handle:
            self._lib.rl_env_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_noise_replay(self, env_words: np.ndarray | None, actor_words: np.ndarray | None):
        """Parity mode: per-lane u32 word streams [num_envs, words_per_lane] replace the Philox source."""
        wpl = 0
        bufs = []
        ptrs = []
        for w in (env_words, actor_words):
            if w is None:
                ptrs.append(None)
                continue
            w = np.ascontiguousarray(w, dtype=np.uint32)
            assert w.shape[0] == self.num_envs
            wpl = max(wpl, w.shape[1])
            b = self.ctx.to_device(w)
            bufs.append(b)
            ptrs.append(b.c)
        if env_words is not None and actor_words is not None:
            assert env_words.shape[1] == actor_words.shape[1]
        self._noise_bufs = bufs
        L.check(self._lib.rl_env_set_noise_replay(self.handle, ptrs[0], ptrs[1], wpl), self.ctx.handle)

    def set_noise_philox(self, seed: int, step_counter: int = 0):
        L.check(self._lib.rl_env_set_noise_philox(self.handle, seed, step_counter), self.ctx.handle)

    # unfused path ----------------------------------------------------------------------------
    def reset_all(self) -> np.ndarray:
        """Environment::initial_state + observe for every lane; returns obs [num_envs, F]."""
        L.check(self._lib.rl_env_reset_all(self.handle), self.ctx.handle)
        return self.observation()

    def observation(self) -> np.ndarray:
        p = C.c_void_p()
        L.check(self._lib.rl_env_observation(self.handle, C.byref(p)), self.ctx.handle)
        return self.ctx.read(p, (self.num_features, self.num_envs), np.float32).T.copy()

    def step_device(self, actions_dev) -> L.StepOut:
        out = L.StepOut()
        a = actions_dev.c if isinstance(actions_dev, DeviceBuffer) else C.c_void_p(actions_dev)
        L.check(self._lib.rl_env_step(self.handle, a, C.byref(out)), self.ctx.handle)
        return out

    def step(self, actions: np.ndarray) -> dict:
        """Environment::step + observe with auto-reset.  Host arrays in, host arrays out (parity/debug)."""
        buf = self.ctx.to_device(np.ascontiguousarray(actions, dtype=np.uint8))
        out = self.step_device(buf)
        E, F = self.num_envs, self.num_features
        res = {
            "obs": self.ctx.read(out.obs, (F, E), np.float32).T.copy(),
            "reward": self.ctx.read(out.reward, (E,), np.float32),
            "succ": self.ctx.read(out.succ, (E,), np.uint8),
            "next_obs": self.ctx.read(out.next_obs, (F, E), np.float32).T.copy(),
        }
        buf.free()
        return res

    def get_state(self):
        E = self.num_envs
        planes = 4 if self.config.kind == L.RL_ENV_CARTPOLE else (
            self.config.env_distribution.num_arms if self.config.kind == L.RL_ENV_BANDIT_META else 0)
        f64 = np.zeros((max(planes, 1), E), np.float64)
        u32 = np.zeros(E, np.uint32)
        L.check(self._lib.rl_env_get_state(self.handle, f64.ctypes.data_as(C.c_void_p) if planes else None,
                                           u32.ctypes.data_as(C.c_void_p)), self.ctx.handle)
        return f64, u32

    def set_state(self, f64: np.ndarray | None, u32: np.ndarray | None):
        f = np.ascontiguousarray(f64, np.float64) if f64 is not None else None
        u = np.ascontiguousarray(u32, np.uint32) if u32 is not None else None
        L.check(self._lib.rl_env_set_state(self.handle, f.ctypes.data_as(C.c_void_p) if f is not None else None,
                                           u.ctypes.data_as(C.c_void_p) if u is not None else None), self.ctx.handle)


def build_env(ctx: Context, config, num_envs: int, seed: int = 0, lane_offset: int = 0) -> BatchedEnv:
    """BuildEnv::build_env (src/envs/builders.rs:17), batched."""
    return BatchedEnv(ctx, config, num_envs, seed, lane_offset)
