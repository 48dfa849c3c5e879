"""The C ABI from a compiled host (what a Rust `-sys` crate sees), without Python in the data path.

* CPU: `include/relearn_b200.h` is valid strict C99; every ctypes mirror in `relearn_b200/_lib.py` has the size and
  field offsets the C compiler gives the header's struct (the probe is generated from the ctypes field lists, so a
  renamed or reordered field fails to compile or compare); a C program linked against the library fails loudly
  without a GPU.
* GPU: `tests/c_host/cartpole_trpo.c` -- the reference's `cartpole-trpo` example flow (examples/cartpole-trpo.rs:14-66)
  in plain C over the header -- prints bit-identical summaries, TRPO / critic statistics and final weights to the
  Python host mirror (`ActorCriticAgent`) driving the same library on the same seed.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import relearn_b200 as R
from relearn_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "relearn_b200")
HOST_SRC = os.path.join(ROOT, "tests", "c_host", "cartpole_trpo.c")

# ctypes mirror -> struct name in the header
STRUCTS = {
    "CartPoleCfg": "rl_cartpole_cfg", "ChainCfg": "rl_chain_cfg", "MemoryCfg": "rl_memory_cfg",
    "BanditMetaCfg": "rl_bandit_meta_cfg", "EnvStructure": "rl_env_structure", "StepOut": "rl_step_out",
    "ActorCfg": "rl_actor_cfg", "Bound": "rl_bound", "MeanVar": "rl_mean_var", "StepsSummary": "rl_steps_summary",
    "TrajView": "rl_traj_view", "TrpoCfg": "rl_trpo_cfg", "TrpoStats": "rl_trpo_stats", "AdamCfg": "rl_adam_cfg",
    "OptStats": "rl_opt_stats", "PpoCfg": "rl_ppo_cfg", "PolicyOptStats": "rl_policy_opt_stats",
    "ReplayStats": "rl_replay_stats", "DqnCfg": "rl_dqn_cfg", "MinibatchView": "rl_minibatch_view",
    "PackedInfo": "rl_packed_info",
}


def _cc(src, out, link=False):
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O1", f"-I{INCLUDE}", src, "-o", out]
    if link:
        cmd += [f"-L{LIBDIR}", "-lrelearn_b200", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_every_ctypes_mirror_is_listed():
    mirrors = {n for n, v in vars(L).items() if isinstance(v, type) and issubclass(v, C.Structure) and v is not C.Structure}
    assert mirrors == set(STRUCTS)


def test_struct_layouts_match_the_header(tmp_path):
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "relearn_b200.h"', "int main(void) {"]
    for py, cname in STRUCTS.items():
        lines.append(f'    printf("{py} %zu", sizeof({cname}));')
        for field, _ in getattr(L, py)._fields_:
            lines.append(f'    printf(" {field}:%zu", offsetof({cname}, {field}));')
        lines.append('    printf("\\n");')
    lines += ["    return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = _cc(str(src), str(tmp_path / "layout"))
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(out) == len(STRUCTS)
    for line in out:
        name, size, *fields = line.split()
        cls = getattr(L, name)
        assert int(size) == C.sizeof(cls), f"{name}: sizeof {size} (C) != {C.sizeof(cls)} (ctypes)"
        assert len(fields) == len(cls._fields_)
        for item in fields:
            field, off = item.split(":")
            assert int(off) == getattr(cls, field).offset, f"{name}.{field}: offset {off} (C) != {getattr(cls, field).offset}"


def _weights(seed):
    rng = np.random.default_rng(seed)
    return R.init_params(rng, 5, 128, 2), R.init_params(rng, 5, 128, 1)


def _run_host(tmp_path, E, T, periods, seed):
    exe = _cc(HOST_SRC, str(tmp_path / "cartpole_trpo"), link=True)
    wp, wc = _weights(seed)
    wfile = tmp_path / "weights.bin"
    np.concatenate([wp, wc]).astype(np.float32).tofile(wfile)
    return subprocess.run([exe, str(wfile), str(E), str(T), str(periods), str(seed)], capture_output=True, text=True,
                          timeout=300)


@pytest.mark.skipif(L.lib().rl_device_count() > 0, reason="checks the no-GPU failure mode")
def test_c_host_fails_loudly_without_gpu(tmp_path):
    r = _run_host(tmp_path, 64, 16, 1, 5)
    assert r.returncode == 10 + L.RL_ERR_CUDA, (r.returncode, r.stdout, r.stderr)
    assert "rl_ctx_create" in r.stderr and "version" in r.stdout


def _fnv1a(data: bytes) -> int:
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
def test_c_host_matches_python_host_bit_for_bit(tmp_path):
    E, T, periods, seed = 512, 96, 3, 11
    r = _run_host(tmp_path, E, T, periods, seed)
    assert r.returncode == 0, (r.stdout, r.stderr)
    rows = [ln.split() for ln in r.stdout.splitlines()]
    by_tag = {}
    for row in rows:
        by_tag.setdefault(row[0], []).append(row[1:])

    ctx = R.Context(0)
    env = R.build_env(ctx, R.CartPoleConfig().wrap(R.VisibleStepLimit(500)), E, seed=seed)
    agent = R.ActorCriticConfig(min_batch_size=R.HistoryDataBound(T, 0)).build_agent(env)
    wp, wc = _weights(seed)
    agent.policy.policy_fn.set_weights(wp)
    agent.critic.state_value_fn.set_weights(wc)
    traj = agent.buffer(agent.min_update_size())
    hx = float.fromhex
    assert by_tag["structure"][0][:2] == [str(env.num_features), str(env.num_actions)]
    for p in range(periods):
        summ = R.rollout(env, agent.actor(), agent.min_update_size(), traj)
        log = {}
        status = agent.batch_update(traj, log)
        row = by_tag["period"][p]
        assert [int(row[0]), int(row[2]), int(row[4])] == [p, summ.num_stored_steps, summ.num_stored_episodes]
        assert [hx(row[6]), hx(row[7]), int(row[8])] == [summ.step_reward.mean, summ.step_reward.squared_residual_sum,
                                                         summ.step_reward.count]
        assert [hx(row[10]), hx(row[11]), int(row[12])] == [summ.episode_length.mean, summ.episode_length.squared_residual_sum,
                                                            summ.episode_length.count]
        t = by_tag["trpo"][p]
        kv = dict(zip(t[1::2], t[2::2]))
        assert int(kv["status"]) == status
        for key, name in (("entropy", "entropy"), ("step_size", "step_size"), ("loss_initial", "loss_initial"),
                          ("loss_final", "loss_final"), ("kl", "constraint_val_final"), ("step_scale", "step_scale")):
            assert hx(kv[key]) == log[name] or (np.isnan(hx(kv[key])) and np.isnan(log[name])), (p, key, kv[key], log[name])
        assert [int(kv["backtracks"]), int(kv["cg"]), int(kv["n"])] == [log["num_backtracks"], log["cg_iterations"], log["num_steps"]]
        c = by_tag["critic"][p]
        ckv = dict(zip(c[1::2], c[2::2]))
        assert hx(ckv["loss_last"]) == log["critic/loss"] and hx(ckv["loss_first"]) == log["critic/loss_first"]
        assert int(ckv["opt_steps"]) == 80
    w = by_tag["weights"][0]
    assert int(w[1], 16) == _fnv1a(agent.policy.policy_fn.get_weights().tobytes())
    assert int(w[3], 16) == _fnv1a(agent.critic.state_value_fn.get_weights().tobytes())
    assert int(by_tag["launches"][0][0]) > 0
    # and the agent learned something in three periods: the critic's loss fell within each update
    assert all(float.fromhex(dict(zip(c[1::2], c[2::2]))["loss_last"]) < float.fromhex(dict(zip(c[1::2], c[2::2]))["loss_first"])
               for c in by_tag["critic"])
