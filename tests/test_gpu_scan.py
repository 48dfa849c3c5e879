"""GPU parity: discounted cumsum / GAE / reward-to-go scans, MLP forward, tabular Q."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
import relearn_b200 as R
from relearn_b200 import _lib as L
from tests import parity as P

pytestmark = pytest.mark.gpu


def _random_succ(rng, T, E, p_done=0.1, pad_tail=True):
    succ = np.where(rng.random((T, E)) < p_done, rng.integers(1, 3, (T, E)), 0).astype(np.uint8)
    lens = rng.integers(1, T + 1, E) if pad_tail else np.full(E, T)
    for e in range(E):
        succ[lens[e]:, e] = L.RL_PAD
        if succ[lens[e] - 1, e] == 0:
            succ[lens[e] - 1, e] = L.RL_INTERRUPT  # stored lanes always end an episode (finalize_last_episode)
    return succ, lens


def test_discounted_cumsum_packed_reference_kat(ctx):
    """torch/packed.rs:979-1008: sequences [1,2,3,4],[5,6],[7,8], discount 0.1."""
    x = np.array([1, 5, 7, 2, 6, 8, 3, 4], np.float32)
    bs = np.array([3, 3, 1, 1], np.uint64)
    L.check(ctx._lib.rl_discounted_cumsum_packed(ctx.handle, x.ctypes.data_as(C.c_void_p), x.size,
                                                 bs.ctypes.data_as(C.POINTER(C.c_uint64)), bs.size, 0.1), ctx.handle)
    np.testing.assert_allclose(x, [1.234, 5.6, 7.8, 2.34, 6.0, 8.0, 3.4, 4.0], rtol=1e-6)
    # bit-exact against the oracle's f32 restatement of the same loop
    y = np.array([1, 5, 7, 2, 6, 8, 3, 4], np.float32)
    bs_sz = np.array([3, 3, 1, 1], np.uintp)
    O.lib().ro_discounted_cumsum_packed_f32(y.ctypes.data_as(C.POINTER(C.c_float)), y.size,
                                            bs_sz.ctypes.data_as(C.POINTER(C.c_size_t)), 4, np.float32(0.1))
    np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("T,E", [(1, 1), (7, 33), (64, 200), (257, 1000)])
def test_discounted_cumsum_bit_exact(ctx, T, E):
    rng = np.random.default_rng(T * 1000 + E)
    x = rng.normal(size=(T, E)).astype(np.float32)
    succ, lens = _random_succ(rng, T, E)
    d = np.float32(0.97)
    xd, sd = ctx.to_device(x), ctx.to_device(succ)
    yd = ctx.alloc(x.nbytes)
    L.check(ctx._lib.rl_discounted_cumsum(ctx.handle, xd.c, sd.c, T, E, d, yd.c), ctx.handle)
    y = yd.download((T, E), np.float32)
    for e in range(E):
        n = int(lens[e])
        ref = O.discounted_cumsum_lane(x[:n, e], succ[:n, e], d)
        np.testing.assert_array_equal(y[:n, e], ref)
        assert (y[n:, e] == 0).all()


def _traj_from_arrays(ctx, cfg, obs, action, reward, succ, next_obs):
    T, E, F = obs.shape
    env = R.build_env(ctx, cfg, E, seed=0)
    traj = R.Trajectory(env, T)
    traj.load(obs, action, reward, succ, next_obs)
    return env, traj


@pytest.mark.parametrize("T,E", [(5, 3), (64, 100), (200, 333)])
def test_reward_to_go_and_gae_without_value_fn_bit_exact(ctx, T, E):
    rng = np.random.default_rng(7 * T + E)
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    obs = rng.normal(size=(T, E, 5)).astype(np.float32)
    nobs = rng.normal(size=(T, E, 5)).astype(np.float32)
    succ, lens = _random_succ(rng, T, E)
    reward = rng.normal(size=(T, E)).astype(np.float32)
    action = rng.integers(0, 2, (T, E)).astype(np.uint8)
    env, traj = _traj_from_arrays(ctx, cfg, obs, action, reward, succ, nobs)
    assert traj.view().num_steps == int(lens.sum())
    adv, rtg = ctx.alloc(T * E * 4), ctx.alloc(T * E * 4)
    gamma, lam = np.float32(0.99), np.float32(0.95)
    L.check(ctx._lib.rl_gae(traj.handle, None, gamma, lam, adv.c, rtg.c), ctx.handle)
    a, r = adv.download((T, E), np.float32), rtg.download((T, E), np.float32)
    zeros = np.zeros(T, np.float32)
    for e in range(E):
        n = int(lens[e])
        np.testing.assert_array_equal(r[:n, e], O.discounted_cumsum_lane(reward[:n, e], succ[:n, e], gamma))
        np.testing.assert_array_equal(a[:n, e], O.gae_lane(reward[:n, e], zeros[:n], zeros[:n], succ[:n, e], gamma, lam))


def test_gae_with_value_fn(ctx):
    """GAE with a critic MLP: the scan is bit-exact given V; V itself is an f32 MLP whose summation order differs
    from the oracle's, so advantages are compared at rtol 2e-5 / atol 2e-5."""
    rng = np.random.default_rng(77)
    T, E, F, H = 96, 150, 5, 128
    cfg = R.CartPoleConfig().wrap(R.VisibleStepLimit(500))
    obs = rng.normal(size=(T, E, F)).astype(np.float32)
    nobs = rng.normal(size=(T, E, F)).astype(np.float32)
    succ, lens = _random_succ(rng, T, E, p_done=0.05)
    reward = np.ones((T, E), np.float32)
    action = rng.integers(0, 2, (T, E)).astype(np.uint8)
    env, traj = _traj_from_arrays(ctx, cfg, obs, action, reward, succ, nobs)
    params = R.init_params(rng, F, H, 1)
    vf = R.Mlp(ctx, F, [H], 1)
    vf.set_weights(params)
    adv, rtg = ctx.alloc(T * E * 4), ctx.alloc(T * E * 4)
    gamma, lam = np.float32(0.99), np.float32(0.95)
    L.check(ctx._lib.rl_gae(traj.handle, vf.handle, gamma, lam, adv.c, rtg.c), ctx.handle)
    a = adv.download((T, E), np.float32)
    v = O.mlp_forward(params, F, H, 1, obs.reshape(-1, F)).reshape(T, E)
    vn = O.mlp_forward(params, F, H, 1, nobs.reshape(-1, F)).reshape(T, E)
    # Mlp.forward on the GPU agrees with the oracle MLP
    np.testing.assert_allclose(vf.forward(obs.reshape(-1, F)).reshape(T, E), v, rtol=2e-5, atol=2e-6)
    for e in range(E):
        n = int(lens[e])
        ref = O.gae_lane(reward[:n, e], v[:n, e], vn[:n, e], succ[:n, e], gamma, lam)
        np.testing.assert_allclose(a[:n, e], ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("shape", [(5, 16, 2), (36, 64, 32), (9, 128, 9)])
def test_mlp_forward(ctx, shape):
    F, H, A = shape
    rng = np.random.default_rng(F * H)
    params = R.init_params(rng, F, H, A)
    net = R.Mlp(ctx, F, [H], A)
    net.set_weights(params)
    np.testing.assert_array_equal(net.get_weights(), params)
    x = rng.normal(size=(777, F)).astype(np.float32)
    np.testing.assert_allclose(net.forward(x), O.mlp_forward(params, F, H, A, x), rtol=2e-5, atol=2e-6)


def test_tabular_q_chain_bit_exact(ctx):
    """config[0] chain-tabular-q: per-replica sequential fold, f64 table and u64 counts bit-identical to the
    oracle fold of the same steps (tabular.rs:159-179)."""
    from relearn_b200.agents import TabularQ

    rng = np.random.default_rng(123)
    cfg = R.Chain()
    E, T, periods, eps = 40, 200, 3, 0.2
    env = R.build_env(ctx, cfg, E, seed=9)
    table = TabularQ(ctx, E, 5, 2, cfg.discount_factor)
    q_ref = np.zeros((E, 5, 2), np.float64)
    c_ref = np.zeros((E, 5, 2), np.uint64)
    traj = R.Trajectory(env, T)
    olib = O.lib()
    for period in range(periods):
        ewords = P.random_words(rng, E, 6 * T)
        awords = P.random_words(rng, E, 6 * T)
        env.set_noise_replay(ewords, awords)
        R.rollout(env, R.ActorSpec(kind=L.RL_ACTOR_TABULAR_EPS_GREEDY, table=table, exploration_rate=eps, training=True),
                  R.HistoryDataBound(T, 0), traj)
        ref = P.oracle_rollout(cfg, E, T, 0, actor_kind=O.ACTOR_TABULAR, env_words=ewords, actor_words=awords,
                               exploration_rate=eps, q_tables=q_ref, training=True)
        P.compare_traj(traj.to_host(), ref, what=f"tabular period {period}")
        table.update(traj)
        for e in range(E):
            n = int(ref["lane_len"][e])
            t = O.TabQ()
            t.n_obs, t.n_act, t.discount = 5, 2, cfg.discount_factor
            t.q = q_ref[e].ctypes.data_as(C.POINTER(C.c_double))
            t.counts = c_ref[e].ctypes.data_as(C.POINTER(C.c_uint64))
            obs_idx = ref["obs"][:n, e].argmax(axis=1).astype(np.uint32)
            nobs_idx = ref["next_obs"][:n, e].argmax(axis=1).astype(np.uint32)
            olib.ro_tabq_update_buffer(C.byref(t), obs_idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       np.ascontiguousarray(ref["action"][:n, e]).ctypes.data_as(C.POINTER(C.c_uint8)),
                                       np.ascontiguousarray(ref["reward"][:n, e]).ctypes.data_as(C.POINTER(C.c_float)),
                                       np.ascontiguousarray(ref["succ"][:n, e]).ctypes.data_as(C.POINTER(C.c_uint8)),
                                       nobs_idx.ctypes.data_as(C.POINTER(C.c_uint32)), n)
        q, c = table.get_table()
        np.testing.assert_array_equal(c, c_ref)
        np.testing.assert_array_equal(q, q_ref)
    assert c_ref.sum() == sum(int(x) for x in [ref["lane_len"].sum()]) * periods


def test_feature_encoders_reference_kats(ctx):
    """spaces/index.rs:310-316, option.rs:239-287, interval.rs:342-389, boolean.rs tests: exact feature rows."""
    def enc(kind, size, elems, dtype, width):
        e = ctx.to_device(np.asarray(elems, dtype))
        out = ctx.alloc(len(elems) * width * 4)
        L.check(ctx._lib.rl_encode_features(ctx.handle, kind, size, e.c, len(elems), out.c), ctx.handle)
        return out.download((len(elems), width), np.float32)
    np.testing.assert_array_equal(enc(L.RL_SPACE_INDEX, 3, [2, 0, 1], np.int64, 3), [[0, 0, 1], [1, 0, 0], [0, 1, 0]])
    np.testing.assert_array_equal(enc(L.RL_SPACE_OPTION_INDEX, 3, [1, -1], np.int64, 4), [[0, 0, 1, 0], [1, 0, 0, 0]])
    np.testing.assert_array_equal(enc(L.RL_SPACE_BOOLEAN, 1, [1, 0], np.int64, 1), [[1], [0]])
    np.testing.assert_array_equal(enc(L.RL_SPACE_INTERVAL, 1, [0.5, -2.25, 1e-3], np.float64, 1),
                                  np.array([[0.5], [-2.25], [1e-3]], np.float32))
