"""ctypes binding of the CPU oracle (``oracle/relearn_oracle.c``).

TEST INFRASTRUCTURE ONLY: importable from ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product package
``relearn_b200`` never imports this module.

The oracle restates the reference (edlanglois/relearn) algorithms on the CPU; see the header of
``relearn_oracle.h`` for what pins it ("parity unpinned" for the RNG word->sample rules, which live
in the un-vendored rand 0.8.5 crate).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librelearn_oracle.so")

CONTINUE, TERMINATE, INTERRUPT = 0, 1, 2
STREAM_ENV_STEP, STREAM_ENV_RESET, STREAM_ACTOR = 0, 1, 2
ENV_CARTPOLE, ENV_CHAIN, ENV_MEMORY, ENV_BANDIT_META, ENV_PARTITION = 0, 1, 2, 3, 4
BANDIT_UNIFORM_BERNOULLI, BANDIT_ROUND_ROBIN_DETERMINISTIC, BANDIT_ONE_HOT = 0, 1, 2
ACTOR_REPLAY, ACTOR_RANDOM, ACTOR_POLICY, ACTOR_EPS_GREEDY_Q, ACTOR_TABULAR = 0, 1, 2, 3, 4
ACT_IDENTITY, ACT_RELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3
MAX_ARMS = 32
MAX_FEATURES = 64


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``make -C oracle``)."""
    src = os.path.join(_HERE, "relearn_oracle.c")
    hdr = os.path.join(_HERE, "relearn_oracle.h")
    stale = (
        force
        or not os.path.exists(_SO)
        or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


class Rng(C.Structure):
    _fields_ = [
        ("mode", C.c_int),
        ("words", C.POINTER(C.c_uint32)),
        ("n_words", C.c_size_t),
        ("cursor", C.c_size_t),
        ("exhausted", C.c_int),
        ("seed", C.c_uint64),
        ("lane", C.c_uint64),
        ("t", C.c_uint32),
        ("draw", C.c_uint32 * 3),
    ]


class UniformF64(C.Structure):
    _fields_ = [("low", C.c_double), ("scale", C.c_double)]


class EnvCfg(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("gravity", C.c_double),
        ("mass_cart", C.c_double),
        ("mass_pole", C.c_double),
        ("length_half_pole", C.c_double),
        ("friction_cart", C.c_double),
        ("friction_pole", C.c_double),
        ("time_step", C.c_double),
        ("action_force", C.c_double),
        ("max_pos", C.c_double),
        ("max_angle", C.c_double),
        ("discount_factor", C.c_double),
        ("max_steps_per_episode", C.c_uint64),
        ("step_limit_visible", C.c_int),
        ("chain_size", C.c_uint64),
        ("num_actions", C.c_uint64),
        ("history_len", C.c_uint64),
        ("num_arms", C.c_uint64),
        ("episodes_per_trial", C.c_uint64),
        ("bandit_dist", C.c_int),
    ]


class State(C.Structure):
    _fields_ = [
        ("x", C.c_double),
        ("xd", C.c_double),
        ("th", C.c_double),
        ("thd", C.c_double),
        ("flag", C.c_int),
        ("steps_remaining", C.c_uint64),
        ("s", C.c_uint64),
        ("s_init", C.c_uint64),
        ("means", C.c_double * MAX_ARMS),
        ("inner_done", C.c_int),
        ("has_prev", C.c_int),
        ("prev_action", C.c_uint64),
        ("prev_reward", C.c_double),
        ("remaining_episodes", C.c_uint64),
    ]


class Env(C.Structure):
    _fields_ = [
        ("cfg", EnvCfg),
        ("total_weight", C.c_double),
        ("inv_total_mass", C.c_double),
        ("mass_length_pole", C.c_double),
        ("reset_dist", UniformF64),
        ("mean_dist", UniformF64),
        ("rr_good_arm", C.c_uint64),
    ]


class Mlp(C.Structure):
    _fields_ = [
        ("in_", C.c_int),
        ("hidden", C.c_int),
        ("out", C.c_int),
        ("act", C.c_int),
        ("params", C.POINTER(C.c_float)),
    ]


class Actor(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("actions", C.POINTER(C.c_uint8)),
        ("n_actions", C.c_size_t),
        ("cursor", C.c_size_t),
        ("mlp", Mlp),
        ("exploration_rate", C.c_double),
        ("training", C.c_int),
        ("q_table", C.POINTER(C.c_double)),
        ("n_obs", C.c_int),
        ("n_act", C.c_int),
        ("act_fn", C.c_void_p),
        ("act_ud", C.c_void_p),
    ]


class Omv(C.Structure):
    _fields_ = [("mean", C.c_double), ("m2", C.c_double), ("count", C.c_uint64)]

    def as_tuple(self):
        return (self.mean, self.m2, self.count)


class Summary(C.Structure):
    _fields_ = [
        ("step_reward", Omv),
        ("episode_reward", Omv),
        ("episode_length", Omv),
        ("cur_len", C.c_uint64),
        ("cur_reward", C.c_double),
    ]


class LaneOut(C.Structure):
    _fields_ = [
        ("obs", C.POINTER(C.c_float)),
        ("action", C.POINTER(C.c_uint8)),
        ("reward", C.POINTER(C.c_float)),
        ("succ", C.POINTER(C.c_uint8)),
        ("next_obs", C.POINTER(C.c_float)),
        ("cap", C.c_size_t),
        ("n_taken", C.c_size_t),
    ]


class Replay(C.Structure):
    _fields_ = [
        ("capacity", C.c_size_t),
        ("succ", C.POINTER(C.c_uint8)),
        ("n", C.c_size_t),
        ("episode_ends", C.POINTER(C.c_uint64)),
        ("n_eps", C.c_size_t),
        ("eps_cap", C.c_size_t),
        ("index_offset", C.c_uint64),
        ("total_step_count", C.c_uint64),
    ]


class TabQ(C.Structure):
    _fields_ = [
        ("n_obs", C.c_int),
        ("n_act", C.c_int),
        ("discount", C.c_double),
        ("q", C.POINTER(C.c_double)),
        ("counts", C.POINTER(C.c_uint64)),
    ]


_lib = None
_aten = None
_SO_ATEN = os.path.join(_HERE, "librelearn_oracle_aten.so")


def build_aten(force: bool = False) -> str:
    """Compile the ATen-actor variant of the CPU rollout baseline (``make -C oracle aten``): the oracle's env loop
    with ``PolicyActor::act`` as batch-1 ATen calls against the torch wheel's libtorch_cpu.so."""
    build(force)
    src = os.path.join(_HERE, "aten_actor.cpp")
    stale = force or not os.path.exists(_SO_ATEN) or os.path.getmtime(_SO_ATEN) < max(os.path.getmtime(src), os.path.getmtime(_SO))
    if stale:
        r = subprocess.run(["make", "-C", _HERE, "aten"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building oracle/aten_actor.cpp failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return _SO_ATEN


def aten_lib() -> C.CDLL:
    """ctypes handle of librelearn_oracle_aten.so (ro_rollout_lanes_aten); imports torch first so that the wheel's
    shared libraries are the ones already loaded."""
    global _aten
    if _aten is None:
        import torch  # noqa: F401

        lib()
        A = C.CDLL(build_aten())
        A.ro_rollout_lanes_aten.restype = C.c_uint64
        A.ro_rollout_lanes_aten.argtypes = [C.POINTER(EnvCfg), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                            C.c_size_t, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(Summary)]
        _aten = A
    return _aten


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    P = C.POINTER
    sig = {
        "ro_rng_script": (None, [P(Rng), P(C.c_uint32), C.c_size_t]),
        "ro_rng_philox": (None, [P(Rng), C.c_uint64, C.c_uint64, C.c_uint32]),
        "ro_rng_set_step": (None, [P(Rng), C.c_uint32]),
        "ro_next_u32": (C.c_uint32, [P(Rng), C.c_int]),
        "ro_next_u64": (C.c_uint64, [P(Rng), C.c_int]),
        "ro_philox4x32_10": (None, [P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]),
        "ro_philox_slot": (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32]),
        "ro_gen_f32": (C.c_float, [P(Rng), C.c_int]),
        "ro_gen_f64": (C.c_double, [P(Rng), C.c_int]),
        "ro_gen_range": (C.c_uint64, [P(Rng), C.c_int, C.c_uint64]),
        "ro_uniform_int": (C.c_uint64, [P(Rng), C.c_int, C.c_uint64]),
        "ro_gen_bool": (C.c_int, [P(Rng), C.c_int, C.c_double]),
        "ro_uniform_inclusive": (UniformF64, [C.c_double, C.c_double]),
        "ro_uniform_sample": (C.c_double, [P(UniformF64), P(Rng), C.c_int]),
        "ro_u64_to_uniform": (C.c_double, [P(UniformF64), C.c_uint64]),
        "ro_u32_to_f32": (C.c_float, [C.c_uint32]),
        "ro_u64_to_f64": (C.c_double, [C.c_uint64]),
        "ro_cfg_cartpole_default": (None, [P(EnvCfg), C.c_uint64]),
        "ro_cfg_chain_default": (None, [P(EnvCfg)]),
        "ro_cfg_memory": (None, [P(EnvCfg), C.c_uint64, C.c_uint64]),
        "ro_cfg_bandit_meta": (None, [P(EnvCfg), C.c_uint64, C.c_uint64, C.c_int]),
        "ro_cfg_partition": (None, [P(EnvCfg)]),
        "ro_env_init": (None, [P(Env), P(EnvCfg)]),
        "ro_env_num_features": (C.c_int, [P(Env)]),
        "ro_env_num_actions": (C.c_int, [P(Env)]),
        "ro_env_num_observations": (C.c_int, [P(Env)]),
        "ro_env_discount": (C.c_double, [P(Env)]),
        "ro_env_reward_range": (None, [P(Env), P(C.c_double), P(C.c_double)]),
        "ro_env_initial_state": (None, [P(Env), P(State), P(Rng)]),
        "ro_env_observe": (None, [P(Env), P(State), P(C.c_float)]),
        "ro_env_observe_index": (C.c_uint64, [P(Env), P(State)]),
        "ro_env_step": (C.c_int, [P(Env), P(State), C.c_uint64, P(Rng), P(C.c_double)]),
        "ro_cartpole_next_state": (None, [P(Env), P(State), C.c_double, P(State)]),
        "ro_mlp_num_params": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
        "ro_mlp_forward": (None, [P(Mlp), P(C.c_float), P(C.c_float), P(C.c_float)]),
        "ro_log_softmax": (None, [P(C.c_float), C.c_int, P(C.c_float)]),
        "ro_categorical_sample": (C.c_int, [P(C.c_float), C.c_int, C.c_float]),
        "ro_omv_push": (None, [P(Omv), C.c_double]),
        "ro_omv_add": (Omv, [Omv, Omv]),
        "ro_summary_push": (None, [P(Summary), C.c_double, C.c_int]),
        "ro_summary_merge": (None, [P(Summary), P(Summary)]),
        "ro_rollout_lane": (
            C.c_size_t,
            [P(Env), P(Actor), C.c_size_t, C.c_size_t, P(Rng), P(Rng), C.c_uint32, P(LaneOut), P(Summary)],
        ),
        "ro_take_aligned_steps": (C.c_size_t, [P(C.c_uint8), C.c_size_t, C.c_size_t, C.c_size_t]),
        "ro_finalize_last_episode": (C.c_size_t, [P(C.c_uint8), C.c_size_t, P(C.c_int)]),
        "ro_default_slack": (C.c_size_t, [C.c_size_t]),
        "ro_div_ceil": (C.c_size_t, [C.c_size_t, C.c_size_t]),
        "ro_replay_init": (C.c_int, [P(Replay), C.c_size_t]),
        "ro_replay_free": (None, [P(Replay)]),
        "ro_replay_write_step": (C.c_int, [P(Replay), C.c_uint8]),
        "ro_replay_end_experience": (None, [P(Replay)]),
        "ro_discounted_cumsum_packed_f64": (
            None,
            [P(C.c_double), C.c_size_t, P(C.c_size_t), C.c_size_t, C.c_double],
        ),
        "ro_discounted_cumsum_packed_f32": (
            None,
            [P(C.c_float), C.c_size_t, P(C.c_size_t), C.c_size_t, C.c_float],
        ),
        "ro_discounted_cumsum_lane_f32": (
            None,
            [P(C.c_float), P(C.c_uint8), C.c_size_t, C.c_float, P(C.c_float)],
        ),
        "ro_gae_lane_f32": (
            None,
            [P(C.c_float), P(C.c_float), P(C.c_float), P(C.c_uint8), C.c_size_t, C.c_float, C.c_float, P(C.c_float)],
        ),
        "ro_td_lane_f32": (
            None,
            [P(C.c_float), P(C.c_float), P(C.c_float), P(C.c_uint8), C.c_size_t, C.c_float, P(C.c_float)],
        ),
        "ro_tabq_step_update": (None, [P(TabQ), C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_uint64]),
        "ro_tabq_update_buffer": (
            None,
            [P(TabQ), P(C.c_uint32), P(C.c_uint8), P(C.c_float), P(C.c_uint8), P(C.c_uint32), C.c_size_t],
        ),
        "ro_argmax_f64": (C.c_int, [P(C.c_double), C.c_int]),
        "ro_rollout_lanes_philox": (
            C.c_uint64,
            [P(EnvCfg), P(Mlp), C.c_uint64, C.c_uint64, C.c_size_t, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int, P(Summary)],
        ),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


# --------------------------------------------------------------------------------------
# numpy helpers
# --------------------------------------------------------------------------------------
def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def cartpole_cfg(step_limit: int = 500) -> EnvCfg:
    c = EnvCfg()
    lib().ro_cfg_cartpole_default(C.byref(c), step_limit)
    return c


def chain_cfg(size: int = 5, discount: float = 0.95) -> EnvCfg:
    c = EnvCfg()
    lib().ro_cfg_chain_default(C.byref(c))
    c.chain_size = size
    c.discount_factor = discount
    return c


def memory_cfg(num_actions: int = 2, history_len: int = 1) -> EnvCfg:
    c = EnvCfg()
    lib().ro_cfg_memory(C.byref(c), num_actions, history_len)
    return c


def bandit_meta_cfg(num_arms: int = 2, episodes_per_trial: int = 10, dist: int = BANDIT_UNIFORM_BERNOULLI) -> EnvCfg:
    c = EnvCfg()
    lib().ro_cfg_bandit_meta(C.byref(c), num_arms, episodes_per_trial, dist)
    return c


def partition_cfg() -> EnvCfg:
    c = EnvCfg()
    lib().ro_cfg_partition(C.byref(c))
    return c


def make_env(cfg: EnvCfg) -> Env:
    e = Env()
    lib().ro_env_init(C.byref(e), C.byref(cfg))
    return e


class ScriptRng:
    """Sequential u32 word stream (how rand_core's BlockRng hands out ChaCha words)."""

    def __init__(self, words):
        self.words = np.ascontiguousarray(words, dtype=np.uint32)
        self.rng = Rng()
        lib().ro_rng_script(C.byref(self.rng), _ptr(self.words, C.c_uint32), self.words.size)

    @property
    def ref(self):
        return C.byref(self.rng)


class PhiloxRng:
    def __init__(self, seed: int, lane: int, t: int = 0):
        self.rng = Rng()
        lib().ro_rng_philox(C.byref(self.rng), seed, lane, t)

    @property
    def ref(self):
        return C.byref(self.rng)


def mlp_struct(params: np.ndarray, n_in: int, hidden: int, n_out: int, act: int = ACT_RELU) -> Mlp:
    assert params.dtype == np.float32 and params.flags.c_contiguous
    assert params.size == hidden * n_in + hidden + n_out * hidden + n_out
    m = Mlp()
    m.in_, m.hidden, m.out, m.act = n_in, hidden, n_out, act
    m.params = _ptr(params, C.c_float)
    return m


def mlp_forward(params: np.ndarray, n_in: int, hidden: int, n_out: int, x: np.ndarray, act: int = ACT_RELU):
    m = mlp_struct(params, n_in, hidden, n_out, act)
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, n_in)
    out = np.zeros((x.shape[0], n_out), dtype=np.float32)
    h = np.zeros(hidden, dtype=np.float32)
    L = lib()
    for i in range(x.shape[0]):
        L.ro_mlp_forward(C.byref(m), _ptr(x[i], C.c_float), _ptr(out[i], C.c_float), _ptr(h, C.c_float))
    return out


def rollout_lane(cfg: EnvCfg, actor: Actor, min_steps: int, slack: int, rng_env, rng_actor=None, t0: int = 0,
                 env: Env | None = None):
    """Run one lane; returns dict of numpy arrays trimmed to the stored length + Summary."""
    L = lib()
    env = env if env is not None else make_env(cfg)
    F = L.ro_env_num_features(C.byref(env))
    cap = max(min_steps + slack, 1)
    obs = np.zeros((cap, F), np.float32)
    nobs = np.zeros((cap, F), np.float32)
    act = np.zeros(cap, np.uint8)
    rew = np.zeros(cap, np.float32)
    succ = np.full(cap, 255, np.uint8)
    out = LaneOut(_ptr(obs, C.c_float), _ptr(act, C.c_uint8), _ptr(rew, C.c_float), _ptr(succ, C.c_uint8),
                  _ptr(nobs, C.c_float), cap, 0)
    summ = Summary()
    ra = rng_actor if rng_actor is not None else rng_env
    n = L.ro_rollout_lane(C.byref(env), C.byref(actor), min_steps, slack, rng_env.ref, ra.ref, t0, C.byref(out),
                          C.byref(summ))
    return {
        "n": n,
        "n_taken": out.n_taken,
        "obs": obs,
        "next_obs": nobs,
        "action": act,
        "reward": rew,
        "succ": succ,
        "summary": summ,
    }


def replay_actor(actions) -> tuple[Actor, np.ndarray]:
    a = Actor()
    arr = np.ascontiguousarray(actions, dtype=np.uint8)
    a.kind = ACTOR_REPLAY
    a.actions = _ptr(arr, C.c_uint8)
    a.n_actions = arr.size
    return a, arr


def policy_actor(params: np.ndarray, n_in: int, hidden: int, n_out: int, kind: int = ACTOR_POLICY,
                 exploration_rate: float = 0.0) -> Actor:
    a = Actor()
    a.kind = kind
    a.mlp = mlp_struct(params, n_in, hidden, n_out)
    a.exploration_rate = exploration_rate
    return a


def discounted_cumsum_lane(x, succ, d):
    x = np.ascontiguousarray(x, np.float32)
    succ = np.ascontiguousarray(succ, np.uint8)
    y = np.zeros_like(x)
    lib().ro_discounted_cumsum_lane_f32(_ptr(x, C.c_float), _ptr(succ, C.c_uint8), x.size, d, _ptr(y, C.c_float))
    return y


def gae_lane(reward, v, v_next_intr, succ, gamma, lam):
    reward = np.ascontiguousarray(reward, np.float32)
    v = np.ascontiguousarray(v, np.float32)
    vn = np.ascontiguousarray(v_next_intr, np.float32)
    succ = np.ascontiguousarray(succ, np.uint8)
    adv = np.zeros_like(reward)
    lib().ro_gae_lane_f32(_ptr(reward, C.c_float), _ptr(v, C.c_float), _ptr(vn, C.c_float), _ptr(succ, C.c_uint8),
                          reward.size, gamma, lam, _ptr(adv, C.c_float))
    return adv


def pack_episodes(episodes):
    """LazyHistoryFeatures::new + PackedStructure (src/torch/agents/features.rs:70-125, packed.rs:346-420):
    episodes sorted by length descending (stable here; the reference's sort_unstable leaves ties unordered),
    then interleaved time-major.  `episodes` is a list of lists of per-step tuples.
    Returns {"steps": packed list of the tuples, "batch_sizes": [...], "order": episode order}."""
    order = sorted(range(len(episodes)), key=lambda i: -len(episodes[i]))
    longest = len(episodes[order[0]]) if episodes else 0
    steps, batch_sizes = [], []
    for t in range(longest):
        row = [episodes[i][t] for i in order if len(episodes[i]) > t]
        steps += row
        batch_sizes.append(len(row))
    return {"steps": steps, "batch_sizes": batch_sizes, "order": order}


# ------------------------------------------------------------------------------------------------
# UCB1 (src/agents/bandits/ucb.rs) -- small-case Python restatement (scalar f64 arithmetic, libm log as f64::ln)
# ------------------------------------------------------------------------------------------------
class Ucb1Oracle:
    """BaseUCB1Agent + BaseUCB1Actor (ucb.rs:78-243)."""

    def __init__(self, num_observations: int, num_actions: int, reward_range, exploration_rate: float = 0.2):
        lo, hi = float(reward_range[0]), float(reward_range[1])
        width = hi - lo
        if not np.isfinite(width):
            raise ValueError("UnboundedReward")  # ucb.rs:112-117
        self.exploration_rate = float(exploration_rate)
        self.reward_scale_factor = 1.0 / width   # :118
        self.reward_shift = -lo                  # :119
        # one success and one failure for each arm (:125-128)
        self.mean = np.full((num_observations, num_actions), 0.5, np.float64)
        self.count = np.full((num_observations, num_actions), 2, np.uint64)
        self.visits = np.full((num_observations,), 2 * num_actions, np.uint64)

    def step_update(self, obs: int, action: int, reward: float):  # :143-160
        scaled = (float(reward) + self.reward_shift) * self.reward_scale_factor
        self.visits[obs] += np.uint64(1)
        self.count[obs, action] += np.uint64(1)
        m = float(self.mean[obs, action])
        self.mean[obs, action] = m + (scaled - m) / float(self.count[obs, action])

    def ucb(self, obs: int):  # :219-231
        import math

        lsv = 2.0 * math.log(float(self.visits[obs]))
        return [math.sqrt(lsv / float(c)) * self.exploration_rate + float(m) for c, m in zip(self.count[obs], self.mean[obs])]

    def act(self, obs: int, training: bool = True) -> int:  # :214-243; argmax_by returns the LAST maximum (cmp.rs:58-76)
        vals = self.ucb(obs) if training else [int(c) for c in self.count[obs]]
        best, bv = 0, vals[0]
        for k, v in enumerate(vals):
            if v >= bv:
                best, bv = k, v
        return best
