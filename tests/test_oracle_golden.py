"""CPU tests: the oracle (oracle/) against every known-answer vector the reference's own tests hold for the
hot path, transcribed value for value with the reference file:line beside each.

The reference is Rust and cannot be built or imported in this image, so these transcriptions are what pins
the oracle.  What no reference test pins (the rand 0.8.5 word -> sample rules, CartPole trajectories, MLP /
Adam numerics) is cross-checked against independent restatements written here and is flagged "unpinned".
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import pytest

import oracle as O

CONT, TERM, INTR = O.CONTINUE, O.TERMINATE, O.INTERRUPT


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


# ------------------------------------------------------------------------------------------------
# scans: src/torch/packed.rs:979-1008
# ------------------------------------------------------------------------------------------------
def _pack(seqs):
    """PackedTensor::from_sorted_sequences (packed.rs:604-625): time-major interleave, longest first."""
    n_b = len(seqs[0])
    data, sizes = [], []
    for t in range(n_b):
        row = [s[t] for s in seqs if len(s) > t]
        data += row
        sizes.append(len(row))
    return data, sizes


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_discounted_cumsum_from_end_kat(dtype):
    seqs = [[1.0, 2.0, 3.0, 4.0], [5.0, 6.0], [7.0, 8.0]]
    data, sizes = _pack(seqs)
    assert sizes == [3, 3, 1, 1]  # packed.rs:1010-1015 batch_sizes_tensor_values
    x = np.array(data, dtype)
    bs = np.array(sizes, np.uint64)
    fn = O.lib().ro_discounted_cumsum_packed_f64 if dtype == np.float64 else O.lib().ro_discounted_cumsum_packed_f32
    ctype = C.c_double if dtype == np.float64 else C.c_float
    fn(x.ctypes.data_as(C.POINTER(ctype)), x.size, bs.ctypes.data_as(C.POINTER(C.c_size_t)), bs.size, dtype(0.1))
    expected, _ = _pack([[1.234, 2.34, 3.4, 4.0], [5.6, 6.0], [7.8, 8.0]])
    tol = 1e-8 if dtype == np.float64 else 1e-6
    np.testing.assert_allclose(x, np.array(expected), rtol=tol, atol=tol)


def test_discounted_cumsum_lane_matches_packed():
    rng = np.random.default_rng(0)
    lens = [7, 5, 5, 2, 1]
    seqs = [rng.normal(size=n).astype(np.float32) for n in lens]
    data, sizes = _pack([list(s) for s in seqs])
    x = np.array(data, np.float32)
    bs = np.array(sizes, np.uint64)
    O.lib().ro_discounted_cumsum_packed_f32(x.ctypes.data_as(C.POINTER(C.c_float)), x.size,
                                            bs.ctypes.data_as(C.POINTER(C.c_size_t)), bs.size, np.float32(0.9))
    # the same episodes laid out as one lane with successor codes
    lane = np.concatenate(seqs)
    succ = np.concatenate([[CONT] * (n - 1) + [TERM] for n in lens]).astype(np.uint8)
    y = O.discounted_cumsum_lane(lane, succ, np.float32(0.9))
    # un-pack the packed result
    off = np.concatenate([[0], np.cumsum(sizes)])
    for s_i, n in enumerate(lens):
        got = np.array([x[off[t] + s_i] for t in range(n)], np.float32)
        start = int(np.sum(lens[:s_i]))
        np.testing.assert_array_equal(got, y[start:start + n])  # bit-exact: same f32 recurrence


# ------------------------------------------------------------------------------------------------
# step limit wrapper: src/envs/wrappers/step_limit.rs:244-300
# ------------------------------------------------------------------------------------------------
def _chain_with_limit(limit, visible=1):
    cfg = O.chain_cfg()
    cfg.max_steps_per_episode = limit
    cfg.step_limit_visible = visible
    return O.make_env(cfg)


@pytest.mark.parametrize("visible", [0, 1])
def test_step_limit_kat(visible):
    env = _chain_with_limit(2, visible)
    L = O.lib()
    rng = O.ScriptRng(np.full(16, 0xFFFFFFFF, np.uint32))  # no slips
    st = O.State()
    L.ro_env_initial_state(C.byref(env), C.byref(st), rng.ref)
    F = L.ro_env_num_features(C.byref(env))
    assert F == (6 if visible else 5)
    obs = np.zeros(F, np.float32)
    L.ro_env_observe(C.byref(env), C.byref(st), obs.ctypes.data_as(C.POINTER(C.c_float)))
    if visible:
        assert obs[5] == 1.0
    r = C.c_double()
    assert L.ro_env_step(C.byref(env), C.byref(st), 0, rng.ref, C.byref(r)) == CONT  # Move::Left
    L.ro_env_observe(C.byref(env), C.byref(st), obs.ctypes.data_as(C.POINTER(C.c_float)))
    if visible:
        assert obs[5] == 0.5
    assert L.ro_env_step(C.byref(env), C.byref(st), 0, rng.ref, C.byref(r)) == INTR
    assert st.steps_remaining == 0


# ------------------------------------------------------------------------------------------------
# meta environment transcript: src/envs/meta.rs:642-769 (RoundRobinDeterministicBandits(2), limit 3)
# ------------------------------------------------------------------------------------------------
def _meta_obs(env, st):
    F = O.lib().ro_env_num_features(C.byref(env))
    o = np.zeros(F, np.float32)
    O.lib().ro_env_observe(C.byref(env), C.byref(st), o.ctypes.data_as(C.POINTER(C.c_float)))
    return o


def _expect_meta(k, inner_some, prev, done):
    """features of MetaObservation{inner_observation, prev_step, episode_done} (meta.rs:357-363)."""
    o = np.zeros(k + 4, np.float32)
    o[0] = 0.0 if inner_some else 1.0          # Option<()>: is_none
    if prev is None:
        o[1] = 1.0
    else:
        a, r = prev
        o[2 + a] = 1.0
        o[2 + k] = r
    o[3 + k] = 1.0 if done else 0.0
    return o


def test_meta_env_expected_steps():
    cfg = O.bandit_meta_cfg(2, 3, O.BANDIT_ROUND_ROBIN_DETERMINISTIC)
    env = O.make_env(cfg)
    L = O.lib()
    rng = O.ScriptRng(np.zeros(4, np.uint32))
    st = O.State()
    r = C.c_double()
    L.ro_env_initial_state(C.byref(env), C.byref(st), rng.ref)
    np.testing.assert_array_equal(_meta_obs(env, st), _expect_meta(2, True, None, False))
    # Trial 0; Ep 0; Step 0: action 0 -> reward 1, inner terminal
    assert L.ro_env_step(C.byref(env), C.byref(st), 0, rng.ref, C.byref(r)) == CONT and r.value == 1.0
    np.testing.assert_array_equal(_meta_obs(env, st), _expect_meta(2, False, (0, 1.0), True))
    # Ep 1 init: action ignored
    assert L.ro_env_step(C.byref(env), C.byref(st), 0, rng.ref, C.byref(r)) == CONT and r.value == 0.0
    np.testing.assert_array_equal(_meta_obs(env, st), _expect_meta(2, True, None, False))
    # Ep 1 step: action 1 -> reward 0
    assert L.ro_env_step(C.byref(env), C.byref(st), 1, rng.ref, C.byref(r)) == CONT and r.value == 0.0
    np.testing.assert_array_equal(_meta_obs(env, st), _expect_meta(2, False, (1, 0.0), True))
    # Ep 2 init
    assert L.ro_env_step(C.byref(env), C.byref(st), 1, rng.ref, C.byref(r)) == CONT and r.value == 0.0
    np.testing.assert_array_equal(_meta_obs(env, st), _expect_meta(2, True, None, False))
    # Ep 2 step: last inner episode of the trial -> Interrupt
    assert L.ro_env_step(C.byref(env), C.byref(st), 0, rng.ref, C.byref(r)) == INTR and r.value == 1.0
    np.testing.assert_array_equal(_meta_obs(env, st), _expect_meta(2, False, (0, 1.0), True))
    # Trial 1: the good arm is now 1
    L.ro_env_initial_state(C.byref(env), C.byref(st), rng.ref)
    np.testing.assert_array_equal(_meta_obs(env, st), _expect_meta(2, True, None, False))
    assert L.ro_env_step(C.byref(env), C.byref(st), 0, rng.ref, C.byref(r)) == CONT and r.value == 0.0
    np.testing.assert_array_equal(_meta_obs(env, st), _expect_meta(2, False, (0, 0.0), True))


def test_one_hot_bandits_meta_env():
    # OneHotBandits (bandits.rs:187-243): one gen_range draw per trial picks the arm that pays exactly 1;
    # DeterministicBandit::step draws nothing (bandits.rs:116-126, rewards test bandits.rs:296-305)
    k, n = 3, 4
    cfg = O.bandit_meta_cfg(k, n, O.BANDIT_ONE_HOT)
    env = O.make_env(cfg)
    L = O.lib()
    words = (np.arange(64, dtype=np.uint64) * 2654435761 % (1 << 32)).astype(np.uint32)
    rng = O.ScriptRng(words)
    st = O.State()
    r = C.c_double()
    for _trial in range(3):
        L.ro_env_initial_state(C.byref(env), C.byref(st), rng.ref)
        means = [st.means[i] for i in range(k)]
        assert sorted(means) == [0.0] * (k - 1) + [1.0]
        good = means.index(1.0)
        for ep in range(n):
            a = (ep + _trial) % k
            code = L.ro_env_step(C.byref(env), C.byref(st), a, rng.ref, C.byref(r))
            assert r.value == (1.0 if a == good else 0.0)
            assert code == (INTR if ep == n - 1 else CONT)
            np.testing.assert_array_equal(_meta_obs(env, st), _expect_meta(k, False, (a, r.value), True))
            if ep < n - 1:
                assert L.ro_env_step(C.byref(env), C.byref(st), 0, rng.ref, C.byref(r)) == CONT and r.value == 0.0


def test_partition_game_oracle():
    # partition.rs: supervisor axis = gen_range(0..10); element = gen::<[bool; 10]>() (MSB of ten u32 words, index
    # order); reward +1 iff the action equals element[axis]; observation = (element, Some((previous element, label)))
    # as [10 bools; is_none; 10 bools; one-hot(Left, Right)] (power.rs:106-115, option.rs:88-116); never ends.
    cfg = O.partition_cfg()
    env = O.make_env(cfg)
    L = O.lib()
    assert L.ro_env_num_features(C.byref(env)) == 23 and L.ro_env_num_actions(C.byref(env)) == 2
    gen = np.random.default_rng(3)
    words = gen.integers(0, 2**32, size=400, dtype=np.uint64).astype(np.uint32)
    rng = O.ScriptRng(words)
    st = O.State()
    r = C.c_double()
    L.ro_env_initial_state(C.byref(env), C.byref(st), rng.ref)
    axis = int(st.s_init)
    assert 0 <= axis < 10
    # gen_range(0..10) consumed one u64 (two words) unless rejected; find where the element words start
    for start in (2, 4, 6):
        bits = [int(w >> 31) for w in words[start:start + 10]]
        if sum(b << i for i, b in enumerate(bits)) == int(st.s):
            break
    else:
        raise AssertionError("first element does not match the MSBs of ten consecutive words")
    obs = _meta_obs(env, st)
    np.testing.assert_array_equal(obs[:10], bits)
    assert obs[10] == 1.0 and not obs[11:].any()
    cursor = start + 10
    for t in range(12):
        element = [int((st.s >> i) & 1) for i in range(10)]
        action = t % 2
        code = L.ro_env_step(C.byref(env), C.byref(st), action, rng.ref, C.byref(r))
        assert code == CONT
        assert r.value == (1.0 if action == element[axis] else -1.0)
        nxt = [int(w >> 31) for w in words[cursor:cursor + 10]]
        cursor += 10
        obs = _meta_obs(env, st)
        np.testing.assert_array_equal(obs[:10], nxt)
        assert obs[10] == 0.0
        np.testing.assert_array_equal(obs[11:21], element)
        np.testing.assert_array_equal(obs[21:], [1.0, 0.0] if element[axis] == 0 else [0.0, 1.0])
        assert int(st.s_init) == axis


def test_meta_trial_length_is_2n_minus_1():
    # a trial of n inner episodes is 2n-1 meta-steps
    for n in (1, 2, 10):
        cfg = O.bandit_meta_cfg(3, n)
        actor, _k = O.replay_actor(np.zeros(64, np.uint8))
        out = O.rollout_lane(cfg, actor, 2 * n - 1, 0, O.ScriptRng(np.arange(4096, dtype=np.uint32) * 2654435761))
        assert out["n"] == 2 * n - 1
        assert list(out["succ"][:2 * n - 1]) == [CONT] * (2 * n - 2) + [INTR]


# ------------------------------------------------------------------------------------------------
# TakeAlignedSteps: src/simulation/take_steps.rs:139-228
# ------------------------------------------------------------------------------------------------
STEPS = _u8([CONT, TERM, CONT, CONT, TERM, CONT, CONT, INTR])


@pytest.mark.parametrize("min_steps,slack,expected", [
    (0, 2, 0),    # take_no_steps
    (100, 2, 8),  # take_all_steps
    (5, 0, 5),    # take_aligned_no_slack
    (5, 2, 5),    # take_aligned_slack
    (3, 0, 3),    # take_unaligned_no_slack
    (3, 2, 5),    # take_unaligned_slack
])
def test_take_aligned_steps(min_steps, slack, expected):
    got = O.lib().ro_take_aligned_steps(STEPS.ctypes.data_as(C.POINTER(C.c_uint8)), STEPS.size, min_steps, slack)
    assert got == expected


# ------------------------------------------------------------------------------------------------
# VecBuffer finalisation: src/agents/buffers/vec.rs:157-276, buffers/mod.rs:237-261
# ------------------------------------------------------------------------------------------------
def test_vec_buffer_finalize_kat():
    succ = _u8([TERM, CONT, TERM, CONT, CONT])
    new_ep = C.c_int()
    n = O.lib().ro_finalize_last_episode(succ.ctypes.data_as(C.POINTER(C.c_uint8)), succ.size, C.byref(new_ep))
    assert n == 4                                   # num_steps: the last step is dropped
    assert list(succ[:4]) == [TERM, CONT, TERM, INTR]  # steps(): step(3, Interrupt(4))
    assert new_ep.value == 1
    assert int(np.sum(succ[:4] != CONT)) == 3       # num_episodes


@pytest.mark.parametrize("codes,n_after,new_ep", [
    ([], 0, 0),
    ([TERM], 1, 0),
    ([CONT], 0, 0),            # popped, nothing before it
    ([TERM, CONT], 1, 0),      # popped step was alone in its episode
    ([CONT, CONT], 1, 1),
    ([CONT, INTR], 2, 0),
])
def test_finalize_last_episode_edges(codes, n_after, new_ep):
    succ = _u8(codes + [0])  # keep a valid pointer for the empty case
    flag = C.c_int()
    n = O.lib().ro_finalize_last_episode(succ.ctypes.data_as(C.POINTER(C.c_uint8)), len(codes), C.byref(flag))
    assert (n, flag.value) == (n_after, new_ep)


# ------------------------------------------------------------------------------------------------
# ReplayBuffer: src/agents/buffers/replay.rs:207-326
# ------------------------------------------------------------------------------------------------
def _write(rb, codes):
    for c in codes:
        rc = O.lib().ro_replay_write_step(C.byref(rb), c)
        if rc != 0:
            return rc
    O.lib().ro_replay_end_experience(C.byref(rb))
    return 0


def _stored(rb):
    return [rb.succ[i] for i in range(rb.n)]


def test_replay_buffer_comprehensive():
    rb = O.Replay()
    assert O.lib().ro_replay_init(C.byref(rb), 7) == 0
    ep1 = [CONT, CONT, TERM]
    assert _write(rb, ep1) == 0
    assert (rb.n, rb.n_eps) == (3, 1) and _stored(rb) == ep1
    assert _write(rb, [CONT, CONT, CONT]) == 0      # ep2 not terminated -> [Continue, Interrupt]
    ep2 = [CONT, INTR]
    assert (rb.n, rb.n_eps) == (5, 2) and _stored(rb) == ep1 + ep2
    ep3 = [CONT, CONT, TERM]
    assert _write(rb, ep3) == 0                      # overflow drops the first episode
    assert (rb.n, rb.n_eps) == (5, 2) and _stored(rb) == ep2 + ep3
    assert _write(rb, [TERM, TERM]) == 0
    assert (rb.n, rb.n_eps) == (7, 4) and _stored(rb) == ep2 + ep3 + [TERM, TERM]
    # episode boundaries (Episodes::get, replay.rs:283-315)
    ends = [rb.episode_ends[i] - rb.index_offset for i in range(rb.n_eps)]
    assert ends == [2, 5, 6, 7]
    # total_step_count excludes the dropped dangling step, includes evicted ones (replay.rs:22-26)
    assert rb.total_step_count == 3 + 2 + 3 + 2
    O.lib().ro_replay_free(C.byref(rb))


def test_replay_buffer_episode_too_large():
    rb = O.Replay()
    O.lib().ro_replay_init(C.byref(rb), 7)
    assert _write(rb, [CONT] * 100) == -1  # WriteExperienceError::Full
    O.lib().ro_replay_free(C.byref(rb))


# ------------------------------------------------------------------------------------------------
# OnlineMeanVariance: src/utils/stats.rs:217-260
# ------------------------------------------------------------------------------------------------
def _collect(vals):
    s = O.Omv()
    for v in vals:
        O.lib().ro_omv_push(C.byref(s), v)
    return s


def test_online_mean_variance_kat():
    s = _collect([1.0, 2.0, 3.0, 4.0])
    assert abs(s.mean - 2.5) < 1e-8
    assert abs(s.m2 / s.count - 1.25) < 1e-8
    a, b = _collect([1.0, 2.0]), _collect([3.0, 4.0])
    c = O.lib().ro_omv_add(a, b)
    assert c.as_tuple() == s.as_tuple()  # stats.rs:236-241 `a + b == c` exactly


def test_steps_summary_matches_numpy():
    rng = np.random.default_rng(3)
    s = O.Summary()
    rewards, succ = rng.normal(size=200), rng.choice([CONT, CONT, CONT, TERM, INTR], size=200)
    for r, c in zip(rewards, succ):
        O.lib().ro_summary_push(C.byref(s), float(r), int(c))
    ends = np.flatnonzero(succ != CONT)
    starts = np.concatenate([[0], ends[:-1] + 1])
    ep_ret = np.array([rewards[a:b + 1].sum() for a, b in zip(starts, ends)])
    ep_len = (ends - starts + 1).astype(float)
    assert s.step_reward.count == 200 and abs(s.step_reward.mean - rewards.mean()) < 1e-12
    assert s.episode_reward.count == len(ends) and abs(s.episode_reward.mean - ep_ret.mean()) < 1e-12
    assert abs(s.episode_length.m2 / s.episode_length.count - ep_len.var()) < 1e-9


# ------------------------------------------------------------------------------------------------
# Categorical: src/torch/distributions/categorical.rs:133-267 (isclose 1e-6)
# ------------------------------------------------------------------------------------------------
NI = float("-inf")


def test_categorical_log_probs_kat():
    import torch

    from oracle import tensor_oracle as TO

    z = torch.tensor([[NI, 0, NI], [NI, 0, NI], [NI, 0, 0], [NI, 0, 0], [-1, 0, 1], [-1, 0, 1], [-1, 0, 1], [0, 0, 0]],
                     dtype=torch.float32)
    elements = torch.tensor([1, 0, 2, 0, 0, 1, 2, 0])
    ln = math.log(math.exp(-1.0) + 1.0 + math.exp(1.0))
    expected = torch.tensor([0.0, NI, -math.log(2.0), NI, -1.0 - ln, -ln, 1.0 - ln, math.log(1 / 3)], dtype=torch.float32)
    actual = TO.Categorical(z).log_prob(elements)
    assert torch.isclose(expected, actual, rtol=1e-6, atol=1e-6).all()
    # the C restatement used by the rollout oracle agrees on the finite rows
    for row, el, exp in zip(z.numpy(), elements.numpy(), expected.numpy()):
        if not np.isfinite(row).all():
            continue
        out = np.zeros(3, np.float32)
        O.lib().ro_log_softmax(row.ctypes.data_as(C.POINTER(C.c_float)), 3, out.ctypes.data_as(C.POINTER(C.c_float)))
        assert abs(out[el] - exp) < 1e-6


def test_categorical_entropy_kat():
    import torch

    from oracle import tensor_oracle as TO

    z = torch.tensor([[NI, 0, NI], [NI, 0, 0], [0, 0, 0], [math.log(0.1), math.log(0.3), math.log(0.6)]],
                     dtype=torch.float32)
    expected = torch.tensor([0.0, -math.log(0.5), -math.log(1 / 3),
                             -0.1 * math.log(0.1) - 0.3 * math.log(0.3) - 0.6 * math.log(0.6)], dtype=torch.float32)
    assert torch.isclose(expected, TO.Categorical(z).entropy(), rtol=1e-6, atol=1e-6).all()


def test_categorical_kl_kat():
    import torch

    from oracle import tensor_oracle as TO

    a = TO.Categorical(torch.tensor([[0.2, 0.3, 0.5], [0.2, 0.3, 0.5], [0.0, 1.0, 0.0]], dtype=torch.float32).log())
    b = TO.Categorical(torch.tensor([[0.2, 0.3, 0.5], [0.7, 0.2, 0.1], [0.2, 0.3, 0.5]], dtype=torch.float32).log())
    expected = torch.tensor([0.0, 0.2 * math.log(0.2 / 0.7) + 0.3 * math.log(0.3 / 0.2) + 0.5 * math.log(0.5 / 0.1),
                             math.log(1.0 / 0.3)], dtype=torch.float32)
    assert torch.isclose(expected, a.kl_divergence_from(b), rtol=1e-6, atol=1e-6).all()


def test_categorical_sample_inverse_cdf():
    # categorical.rs:118-131 `sample`: degenerate rows must return their only support point
    for row, want in (([0.0, NI, NI], 0), ([NI, 0.0, NI], 1), ([NI, NI, 0.0], 2)):
        z = np.array([v if v != NI else -1e30 for v in row], np.float32)
        for u in (0.0, 0.3, 0.999999):
            assert O.lib().ro_categorical_sample(z.ctypes.data_as(C.POINTER(C.c_float)), 3, u) == want
    z = np.zeros(3, np.float32)
    got = [O.lib().ro_categorical_sample(z.ctypes.data_as(C.POINTER(C.c_float)), 3, u) for u in (0.1, 0.4, 0.9)]
    assert got == [0, 1, 2]


# ------------------------------------------------------------------------------------------------
# Conjugate gradient optimizer: src/torch/optimizers/conjugate_gradient.rs:411-558, optimizers/mod.rs:171-214
# ------------------------------------------------------------------------------------------------
def test_hvp_quadratic_kat():
    import torch

    from oracle import tensor_oracle as TO

    m = torch.tensor([[1.0, -1.0], [-1.0, 2.0]])
    b = torch.tensor([2.0, -3.0])
    x = torch.zeros(2, requires_grad=True)
    y = m.mv(x).dot(x) / 2 + b.dot(x)
    hvp = TO.HessianVectorProduct(y, [x], 0.0)
    assert torch.equal(hvp.mat_vec_mul(torch.tensor([1.0, 0.0])), torch.tensor([1.0, -1.0]))
    assert torch.equal(hvp.mat_vec_mul(torch.tensor([0.0, 1.0])), torch.tensor([-1.0, 2.0]))


def test_cg_solve_2x2_kat():
    import torch

    from oracle import tensor_oracle as TO

    a = torch.tensor([[1.0, -1.0], [-1.0, 2.0]], dtype=torch.float64)
    b = torch.tensor([-1.0, 4.0], dtype=torch.float64)
    x, _ = TO.solve_conjugate_gradient(TO.MatrixProduct(a), b, 10, 1e-4)
    assert float((x - torch.tensor([2.0, 3.0], dtype=torch.float64)).norm()) < 1e-4


def test_trust_region_optimizes_quadratic():
    import torch

    from oracle import tensor_oracle as TO

    m = torch.tensor([[1.0, -1.0], [-1.0, 2.0]])
    b = torch.tensor([2.0, -3.0])
    x = torch.zeros(2, requires_grad=True)
    x_last = x.detach().clone()

    def fn():
        return m.mv(x).dot(x) / 2 + b.dot(x), (x - x_last).square().sum()

    for _ in range(500):
        x_last.copy_(x.detach())
        try:
            TO.trust_region_backward_step([x], fn, 0.001, TO.CgConfig(), {})
        except TO.OptimizerStepError as e:
            assert e.kind == "LossNotImproving"
            break
    assert float((x.detach() - torch.tensor([-1.0, 1.0])).norm()) < 1e-3


def test_trust_region_unused_params():
    import torch

    from oracle import tensor_oracle as TO

    x = torch.ones(2, requires_grad=True)
    unused = torch.zeros(3, requires_grad=True)
    x_prev = x.detach().clone()

    def fn():
        return x.square().sum(), (x - x_prev).square().sum()

    for _ in range(100):
        x_prev.copy_(x.detach())
        try:
            TO.trust_region_backward_step([x, unused], fn, 0.1, TO.CgConfig(), {})
        except TO.OptimizerStepError as e:
            assert e.kind == "LossNotImproving"
            break
    assert float(x.detach().norm()) < 0.1
    assert torch.equal(unused.detach(), torch.zeros(3))


# ------------------------------------------------------------------------------------------------
# History packing: src/torch/agents/features.rs:293-406
# ------------------------------------------------------------------------------------------------
def test_history_features_packing_kat():
    episodes = [
        [(True, 0, 1.0), (True, 1, 1.0), (True, 2, 1.0), (True, 3, 1.0)],
        [(False, 10, -1.0), (False, 11, -1.0), (False, 12, 0.0), (False, 13, 0.0), (False, 14, 1.0), (False, 15, 1.0)],
        [(False, 20, 2.0), (True, 21, 2.0), (False, 22, 2.0)],
        [(True, 30, 3.0)],
    ]
    packed = O.pack_episodes(episodes)
    assert packed["batch_sizes"] == [4, 3, 3, 2, 1, 1]
    assert [s[1] for s in packed["steps"]] == [10, 0, 20, 30, 11, 1, 21, 12, 2, 22, 13, 3, 14, 15]
    assert [float(s[0]) for s in packed["steps"]] == [0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 0, 1, 0, 0]
    assert [s[2] for s in packed["steps"]] == [-1, 1, 2, 3, -1, 1, 2, 0, 1, 2, 0, 1, 1, 1]


# ------------------------------------------------------------------------------------------------
# Bandit rewards: src/envs/bandits.rs:296-305 (deterministic, exact) and :257-282 (Bernoulli, 3.5 sigma)
# ------------------------------------------------------------------------------------------------
def test_round_robin_deterministic_bandit_rewards():
    cfg = O.bandit_meta_cfg(2, 1, O.BANDIT_ROUND_ROBIN_DETERMINISTIC)
    env = O.make_env(cfg)
    st, r = O.State(), C.c_double()
    rng = O.ScriptRng(np.zeros(1, np.uint32))
    for trial in range(4):
        O.lib().ro_env_initial_state(C.byref(env), C.byref(st), rng.ref)
        O.lib().ro_env_step(C.byref(env), C.byref(st), 0, rng.ref, C.byref(r))
        assert r.value == (1.0 if trial % 2 == 0 else 0.0)  # envs/testing.rs:147-160


def test_bernoulli_bandit_statistics():
    # bandits.rs:257-282: mean reward within 3.5 sigma of p over 1000 pulls
    cfg = O.bandit_meta_cfg(2, 1000)
    env = O.make_env(cfg)
    st, r = O.State(), C.c_double()
    rng = O.PhiloxRng(5, 0, 0)
    O.lib().ro_env_initial_state(C.byref(env), C.byref(st), rng.ref)
    p = st.means[1]
    total, n = 0.0, 1000
    for i in range(n):
        O.lib().ro_rng_set_step(rng.ref, i + 1)
        O.lib().ro_env_step(C.byref(env), C.byref(st), 1, rng.ref, C.byref(r))   # pull
        total += r.value
        O.lib().ro_env_step(C.byref(env), C.byref(st), 0, rng.ref, C.byref(r))   # reset step
    assert abs(total / n - p) < 3.5 * math.sqrt(p * (1 - p) / n) + 1e-12


# ------------------------------------------------------------------------------------------------
# CartPole: no reference KAT exists; cross-check the C oracle against an independent
# numpy restatement of cartpole.rs:306-446 written here, and against the scratch vector of SURVEY A.1.
# ------------------------------------------------------------------------------------------------
def _py_cartpole(x, xd, th, thd, flag, force):
    g, mc, mp, l, muc, mup, dt = 9.8, 1.0, 0.1, 0.5, 0.01, 0.01, 0.02
    tw, inv_m, ml = g * (mc + mp), 1.0 / (mc + mp), mp * l
    s, c, w2 = math.sin(th), math.cos(th), thd * thd

    def acc_nf(mu):
        alpha = (-force - ml * w2 * (s + mu * c)) * inv_m
        beta = mup * thd / ml
        num = g * s + c * (alpha + g * mu) - beta
        den = l * (4.0 / 3.0 - mp * c * inv_m * (c - mu))
        a = num / den
        return a, tw - ml * (a * s + w2 * c)

    mu = muc if flag else -muc
    a, nf = acc_nf(mu)
    positive = math.copysign(1.0, nf * xd) > 0
    if positive != flag:
        mu = -mu
        a, nf = acc_nf(mu)
    xacc = (force + ml * (w2 * s + a * c) + (-mu * nf)) * inv_m
    xd2 = xd + dt * xacc
    return x + dt * xd2, xd2, th + dt * thd, thd + dt * a, positive


def test_cartpole_next_state_against_independent_restatement():
    env = O.make_env(O.cartpole_cfg(500))
    rng = np.random.default_rng(11)
    for _ in range(2000):
        x, xd, th, thd = rng.uniform(-2.4, 2.4), rng.uniform(-3, 3), rng.uniform(-0.21, 0.21), rng.uniform(-3, 3)
        flag, force = bool(rng.integers(0, 2)), float(rng.choice([-10.0, 10.0]))
        st, out = O.State(), O.State()
        st.x, st.xd, st.th, st.thd, st.flag = x, xd, th, thd, int(flag)
        O.lib().ro_cartpole_next_state(C.byref(env), C.byref(st), force, C.byref(out))
        want = _py_cartpole(x, xd, th, thd, flag, force)
        assert (out.x, out.xd, out.th, out.thd, bool(out.flag)) == want  # bit-exact f64


def test_cartpole_survey_vector():
    env = O.make_env(O.cartpole_cfg(500))
    st = O.State()
    st.x, st.xd, st.th, st.thd, st.flag = 0.01, -0.02, 0.03, -0.04, 1
    expected = [
        (0.013015601195599434, 0.15078005977997169, 0.0292, -0.32617393275485651, 0),
        (0.012634562755588809, -0.019051922000531241, 0.022676521344902868, -0.019399713993250411, 1),
        (0.015666775663285573, 0.15161064538483812, 0.022288527065037859, -0.30808300550101747, 0),
    ]
    for force, exp in zip((10.0, -10.0, 10.0), expected):
        out = O.State()
        O.lib().ro_cartpole_next_state(C.byref(env), C.byref(st), force, C.byref(out))
        np.testing.assert_allclose([out.x, out.xd, out.th, out.thd], exp[:4], rtol=1e-13, atol=0)
        assert out.flag == exp[4]
        st = out


def test_cartpole_episode_protocol():
    # cartpole.rs:128-153 + step_limit.rs:202-223: reward 1 per step, Terminate on |x|>2.4 or |theta|>12deg,
    # Interrupt after 500 steps (never reached by a constant-force policy)
    cfg = O.cartpole_cfg(500)
    actor, _k = O.replay_actor(np.ones(600, np.uint8))
    out = O.rollout_lane(cfg, actor, 600, 0, O.PhiloxRng(1, 0, 0))
    n = out["n"]
    assert np.all(out["reward"][:n] == 1.0)
    first_end = int(np.flatnonzero(out["succ"][:n] != CONT)[0])
    assert out["succ"][first_end] == TERM and first_end < 60
    assert out["obs"][0, 4] == 1.0 and out["obs"][1, 4] == np.float32(499 / 500)
    assert np.all(np.abs(out["obs"][:n, 0]) <= 2.4) and np.all(np.abs(out["obs"][:n, 2]) <= math.radians(12.0))
    # a fresh episode starts right after the terminal step: remaining back to 1.0, state within +-0.05
    assert out["obs"][first_end + 1, 4] == 1.0 and np.all(np.abs(out["obs"][first_end + 1, :4]) <= 0.05)


# ------------------------------------------------------------------------------------------------
# Noise: Philox4x32-10 known answers (Random123 kat_vectors) and the rand 0.8.5 conversion rules
# (third-party crate, not vendored by the reference: PARITY UNPINNED, checked against the published rules)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ctr,key,out", [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
])
def test_philox4x32_10_known_answers(ctr, key, out):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    O.lib().ro_philox4x32_10(c, k, o)
    assert tuple(o) == out


def test_rand_conversions():
    L = O.lib()
    assert L.ro_u32_to_f32(0) == 0.0
    assert L.ro_u32_to_f32(0xFFFFFFFF) == (2 ** 24 - 1) / 2 ** 24       # 24 high bits
    assert L.ro_u32_to_f32(0x000000FF) == 0.0
    assert L.ro_u64_to_f64(0xFFFFFFFFFFFFFFFF) == (2 ** 53 - 1) / 2 ** 53  # 53 high bits
    u = L.ro_uniform_inclusive(-0.05, 0.05)
    lo = L.ro_u64_to_uniform(C.byref(u), 0)
    hi = L.ro_u64_to_uniform(C.byref(u), 0xFFFFFFFFFFFFFFFF)
    assert lo == -0.05 and hi <= 0.05 and hi > 0.05 - 1e-15              # new_inclusive: max maps to <= high
    # gen_bool: p == 1 consumes nothing and is always true; p == 0 never true
    rng = O.ScriptRng(np.array([0xFFFFFFFF, 0xFFFFFFFF], np.uint32))
    assert L.ro_gen_bool(rng.ref, 0, 1.0) == 1 and rng.rng.cursor == 0
    assert L.ro_gen_bool(rng.ref, 0, 0.0) == 0 and rng.rng.cursor == 2
    # gen_range: widening multiply, result = high word
    rng = O.ScriptRng(np.array([0, 0x80000000], np.uint32))               # u64 = 2^63
    assert L.ro_gen_range(rng.ref, 0, 2) == 1
    rng = O.ScriptRng(np.array([0xFFFFFFFF, 0x7FFFFFFF], np.uint32))      # u64 = 2^63 - 1
    assert L.ro_gen_range(rng.ref, 0, 2) == 0
    # next_u64 = two consecutive words, low first (rand_core BlockRng)
    rng = O.ScriptRng(np.array([0x11111111, 0x22222222], np.uint32))
    assert L.ro_next_u64(rng.ref, 0) == 0x2222222211111111


def test_gen_range_is_uniform_and_in_bounds():
    rng = O.PhiloxRng(9, 0, 0)
    counts = np.zeros(5, int)
    for i in range(5000):
        O.lib().ro_rng_set_step(rng.ref, i)
        counts[O.lib().ro_gen_range(rng.ref, O.STREAM_ACTOR, 5)] += 1
    assert counts.sum() == 5000 and np.all(np.abs(counts - 1000) < 5 * math.sqrt(1000 * 0.8))


# ------------------------------------------------------------------------------------------------
# Tabular Q: src/agents/tabular.rs:159-179 (hand-computed fold) and behavioural pin :243-284
# ------------------------------------------------------------------------------------------------
def test_tabular_q_fold_by_hand():
    q = np.zeros((2, 2), np.float64)
    cnt = np.zeros((2, 2), np.uint64)
    t = O.TabQ(2, 2, 0.5, q.ctypes.data_as(C.POINTER(C.c_double)), cnt.ctypes.data_as(C.POINTER(C.c_uint64)))
    L = O.lib()
    L.ro_tabq_step_update(C.byref(t), 0, 1, 4.0, 1, 1)   # Q[0,1] = 0*(1-1) + 1*(4 + .5*max Q[1]) = 4
    assert q[0, 1] == 4.0 and cnt[0, 1] == 1
    L.ro_tabq_step_update(C.byref(t), 1, 0, 2.0, 1, 0)   # Q[1,0] = 2 + .5*4 = 4
    assert q[1, 0] == 4.0
    L.ro_tabq_step_update(C.byref(t), 0, 1, 0.0, 0, 0)   # terminal: w = 1/2 -> Q[0,1] = 4*.5 + .5*0 = 2
    assert q[0, 1] == 2.0 and cnt[0, 1] == 2
    assert L.ro_argmax_f64(np.zeros(3).ctypes.data_as(C.POINTER(C.c_double)), 3) == 0  # ties -> first index


def test_tabular_q_learns_deterministic_bandit():
    # agents/testing.rs:14-64 via tabular.rs:243-284: >= 90% optimal arm after training
    cfg = O.bandit_meta_cfg(2, 1, O.BANDIT_ROUND_ROBIN_DETERMINISTIC)
    q = np.zeros((1, 2), np.float64)
    cnt = np.zeros((1, 2), np.uint64)
    t = O.TabQ(1, 2, 0.0, q.ctypes.data_as(C.POINTER(C.c_double)), cnt.ctypes.data_as(C.POINTER(C.c_uint64)))
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = int(rng.integers(0, 2))
        O.lib().ro_tabq_step_update(C.byref(t), 0, a, 1.0 if a == 1 else 0.0, 0, 0)
    assert O.lib().ro_argmax_f64(q[0].ctypes.data_as(C.POINTER(C.c_double)), 2) == 1


# ------------------------------------------------------------------------------------------------
# Chain<Gru, Linear>: packed == iterated steps (src/torch/modules/testing.rs:124-157)
# ------------------------------------------------------------------------------------------------
def test_gru_linear_packed_matches_iterated_steps():
    from oracle import tensor_oracle as TO

    rng = np.random.default_rng(5)
    F, H, A = 6, 4, 2
    n = 3 * H * F + 3 * H * H + 6 * H + A * H + A
    flat = (rng.normal(size=n) * 0.5).astype(np.float32)
    eps = [rng.normal(size=(L, F)).astype(np.float32) for L in (5, 3, 7, 1)]
    iterated = [TO.gru_linear_episode(flat, F, H, A, e) for e in eps]
    packed = TO.gru_packed_episodes(flat, F, H, A, eps)
    for a, b in zip(iterated, packed):
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-6)
    # identical inputs in different episodes give identical outputs (modules/testing.rs:80-122)
    same = TO.gru_packed_episodes(flat, F, H, A, [eps[0], eps[0][:3]])
    np.testing.assert_allclose(same[0][:3], same[1], rtol=0, atol=1e-6)


def test_ucb1_oracle_restates_the_reference_rules():
    """ucb.rs: initial tables (:125-128), reward scaling (:118-123,:145), incremental mean (:157-159), the confidence bound
    (:219-231) and argmax_by's last-maximum tie rule (utils/iter/cmp.rs:58-76)."""
    import math

    o = O.Ucb1Oracle(3, 4, (-1.0, 3.0), 0.2)
    assert (o.mean == 0.5).all() and (o.count == 2).all() and (o.visits == 8).all()
    assert o.act(0, training=True) == 3 and o.act(0, training=False) == 3  # all equal: the LAST maximal element
    o.step_update(1, 2, 3.0)  # scaled reward (3 - (-1)) / 4 = 1
    assert o.visits[1] == 9 and o.count[1, 2] == 3 and o.mean[1, 2] == 0.5 + (1.0 - 0.5) / 3.0
    want = math.sqrt(2.0 * math.log(9.0) / 3.0) * 0.2 + o.mean[1, 2]
    assert o.ucb(1)[2] == want and o.act(1) == 2
    o.step_update(1, 0, -1.0)  # scaled 0
    assert o.mean[1, 0] == 0.5 + (0.0 - 0.5) / 3.0
    assert o.act(1, training=False) == 2  # counts [3, 2, 3, 2]: last maximum
    with pytest.raises(ValueError):
        O.Ucb1Oracle(2, 2, (0.0, float("inf")))


def test_deep_mlp_restatement_and_parameter_order():
    """Mlp::forward with several hidden layers (mlp.rs:139-151: activation between Linear layers, none on the output) and
    Module::variables() order ([W, b] per Linear, layers in order: mlp.rs:126-128, linear.rs:108-110): the torch restatement,
    the numpy one used by the GPU tests and a hand-written loop agree; the host-side parameter count matches."""
    torch = pytest.importorskip("torch")
    from oracle import tensor_oracle as TO
    import relearn_b200.modules as M
    from tests import parity as P

    rng = np.random.default_rng(3)
    F, hidden, A = 6, [9, 4, 7], 3
    flat = M.init_params(rng, F, hidden, A)
    assert flat.size == M.num_params(F, hidden, A) == 6 * 9 + 9 + 9 * 4 + 4 + 4 * 7 + 7 + 7 * 3 + 3
    assert M.num_params(5, 128, 2) == M.num_params(5, [128], 2) == 5 * 128 + 128 + 128 * 2 + 2
    x = rng.normal(size=(11, F)).astype(np.float32)
    params = TO.unflatten_mlp(torch.tensor(flat, dtype=torch.float64), F, hidden, A)
    assert [tuple(p.shape) for p in params] == [(9, 6), (9,), (4, 9), (4,), (7, 4), (7,), (3, 7), (3,)]
    with TO.mlp_activation("tanh"):
        got = TO.mlp_forward(params, torch.tensor(x, dtype=torch.float64)).numpy()
    h = x.astype(np.float64)
    o = 0
    dims = [F] + hidden + [A]
    for li in range(len(dims) - 1):
        w = flat[o:o + dims[li + 1] * dims[li]].astype(np.float64).reshape(dims[li + 1], dims[li]); o += w.size
        b = flat[o:o + dims[li + 1]].astype(np.float64); o += b.size
        h = np.stack([[sum(w[j, f] * row[f] for f in range(dims[li])) + b[j] for j in range(dims[li + 1])] for row in h])
        if li < len(dims) - 2:
            h = np.tanh(h)
    np.testing.assert_allclose(got, h, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(P.mlp_forward_any(flat, F, hidden, A, x, "tanh"), h.astype(np.float32), rtol=1e-6, atol=1e-7)
