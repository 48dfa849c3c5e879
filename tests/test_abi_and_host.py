"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every symbol that
include/relearn_b200.h declares, compute entry points fail loudly without a GPU (no CPU fallback), and the
host-side mirror of the reference's bookkeeping (HistoryDataBound, schedules, lane sharding, summary merge)
matches the reference's rules."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    if not os.path.exists(g.LIB):
        g.build()
    from relearn_b200 import _lib

    return _lib.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "relearn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in relearn_b200.h but not exported: {missing}"


def test_binding_covers_header():
    from relearn_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_no_oracle_in_product():
    """The product package must not import, link or dlopen the oracle."""
    pkg = os.path.join(ROOT, "relearn_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "relearn_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_version_and_status_strings(lib):
    assert lib.rl_version() == (0 << 16 | 1) or lib.rl_version() > 0
    assert lib.rl_status_str(0)
    for code in (1, 2, 3, 4, 5, 6, 16, 17, 18, 19):
        assert lib.rl_status_str(code)


def test_compute_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from relearn_b200 import _lib as L

    h = C.c_void_p()
    st = lib.rl_ctx_create(0, None, C.byref(h))
    assert st == L.RL_ERR_CUDA
    with pytest.raises(L.RelearnB200Error):
        import relearn_b200 as R

        R.Context(0)


def test_config_defaults_match_reference(lib):
    from relearn_b200 import _lib as L

    cp = L.CartPoleCfg()
    lib.rl_cartpole_cfg_default(C.byref(cp), 500)
    # cartpole.rs:178-216
    assert (cp.gravity, cp.mass_cart, cp.mass_pole, cp.length_half_pole) == (9.8, 1.0, 0.1, 0.5)
    assert (cp.friction_cart, cp.friction_pole, cp.time_step) == (0.01, 0.01, 0.02)
    assert (cp.action_force, cp.max_pos, cp.discount_factor) == (10.0, 2.4, 0.99)
    assert cp.max_angle == np.deg2rad(12.0) and cp.max_steps_per_episode == 500
    ch = L.ChainCfg()
    lib.rl_chain_cfg_default(C.byref(ch))
    assert (ch.size, ch.discount_factor) == (5, 0.95)  # chain.rs:38-45
    t = L.TrpoCfg()
    lib.rl_trpo_cfg_default(C.byref(t))
    # trpo.rs:29-38, conjugate_gradient.rs:55-64
    assert (t.max_policy_step_kl, t.cg_iterations, t.max_backtracks, t.backtrack_ratio, t.hpv_reg_coeff,
            t.accept_violation) == (0.01, 10, 15, 0.8, 1e-5, 0)
    a = L.AdamCfg()
    lib.rl_adam_cfg_default(C.byref(a))
    assert (a.learning_rate, a.beta1, a.beta2, a.weight_decay, a.eps) == (1e-3, 0.9, 0.999, 0.0, 1e-8)


def test_exploration_rate_schedule(lib):
    # schedules.rs:35-45 with DqnConfig defaults 1.0 -> 0.1 over 10M steps (dqn.rs:57-72)
    f = lib.rl_exploration_rate
    assert f(1.0, 0.1, 10_000_000, 0, 1) == 1.0
    assert abs(f(1.0, 0.1, 10_000_000, 5_000_000, 1) - 0.55) < 1e-12
    assert abs(f(1.0, 0.1, 10_000_000, 10_000_000, 1) - 0.1) < 1e-12
    assert abs(f(1.0, 0.1, 10_000_000, 50_000_000, 1) - 0.1) < 1e-12
    assert f(1.0, 0.1, 10_000_000, 123, 0) == 0.0  # evaluation mode never explores


def test_philox_slot_host_matches_oracle(lib):
    import oracle as O

    for seed, lane, t, stream, draw in [(0, 0, 0, 0, 0), (1234, 4095, 255, 2, 0), (2 ** 63 + 5, 2 ** 33, 7, 1, 3)]:
        assert lib.rl_philox_slot(seed, lane, t, stream, draw) == O.lib().ro_philox_slot(seed, lane, t, stream, draw)


def test_history_data_bound():
    from relearn_b200 import HistoryDataBound as B

    # buffers/mod.rs:57-63 default slack = clamp(min_steps / 100, 5, 1000)
    assert B.with_default_slack(100).slack_steps == 5
    assert B.with_default_slack(100_000).slack_steps == 1000
    assert B.with_default_slack(1_000_000).slack_steps == 1000
    assert B.with_default_slack(10_000).slack_steps == 100
    # :68-86 max / divide (min_steps divided rounding up, slack unchanged)
    assert B(10_000, 100).divide(3) == B(3334, 100)
    assert B(10_000, 100).divide(16).max(B(10_000, 0)) == B(10_000, 100)  # train.rs:111-118 worker sizing
    import oracle as O

    for n in (1, 99, 100, 499, 500, 10_000, 99_999, 100_000, 10 ** 7):
        assert B.with_default_slack(n).slack_steps == O.lib().ro_default_slack(n)


def test_shard_lanes_partition():
    from relearn_b200.parallel import shard_lanes

    for E in (1, 7, 4096, 65_536, 1_000_003):
        for G in (1, 2, 3, 4, 8):
            parts = [shard_lanes(E, r, G) for r in range(G)]
            assert sum(c for c, _ in parts) == E
            off = 0
            for c, o in parts:
                assert o == off
                off += c
            assert max(c for c, _ in parts) - min(c for c, _ in parts) <= 1
    with pytest.raises(ValueError):
        shard_lanes(8, 2, 2)


def test_mean_var_merge_matches_oracle():
    import oracle as O
    from relearn_b200.parallel import MeanVar, merge_summaries

    rng = np.random.default_rng(0)
    chunks = [rng.normal(size=n) for n in (5, 1, 0, 17)]
    parts = []
    for c in chunks:
        s = O.Omv()
        for v in c:
            O.lib().ro_omv_push(C.byref(s), float(v))
        parts.append([MeanVar(s.mean, s.m2, s.count)] * 3)
    merged = merge_summaries(parts)[0]
    allv = np.concatenate(chunks)
    assert merged.count == allv.size
    assert abs(merged.mean - allv.mean()) < 1e-12
    assert abs(merged.variance() - allv.var()) < 1e-12
    # stats.rs:236-241: [1,2] + [3,4] == collect([1,2,3,4]) exactly
    a = MeanVar(1.5, 0.5, 2).merge(MeanVar(3.5, 0.5, 2))
    assert (a.mean, a.squared_residual_sum, a.count) == (2.5, 5.0, 4)


def test_init_params_shape_and_limits():
    from relearn_b200 import init_params

    p = init_params(np.random.default_rng(0), 5, 128, 2)
    assert p.dtype == np.float32 and p.size == 1026  # P = 128*5 + 128 + 2*128 + 2
    lim1, lim2 = np.sqrt(6.0 / (5 + 1 + 128)), np.sqrt(6.0 / (128 + 1 + 2))  # initializers.rs:159-163, linear.rs:56
    assert np.all(np.abs(p[:768]) <= lim1) and np.all(np.abs(p[768:]) <= lim2)
