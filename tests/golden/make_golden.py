"""Generate the committed golden trajectories under tests/golden/ from the CPU oracle.

The reference (Rust) cannot be built or imported in this image, so these fixtures are outputs of the ORACLE
(oracle/relearn_oracle.c), which is itself pinned against the reference's own known-answer tests in
tests/test_oracle_golden.py.  They serve two purposes: (i) the CPU suite checks that the oracle still reproduces
them (guards the checker against drift), (ii) the GPU suite checks the CUDA rollout against files, not only
against a freshly computed oracle run.  Every fixture stores its own noise words and actions, so nothing depends
on a random generator's stream.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle as O  # noqa: E402
import relearn_b200 as R  # noqa: E402  (config dataclasses only; no GPU needed)
from tests import parity as P  # noqa: E402

CASES = {
    "cartpole_limit20": (R.CartPoleConfig().wrap(R.VisibleStepLimit(20)), 6, 48, 3),
    "chain5": (R.Chain(), 6, 40, 0),
    "memory_3_2": (R.MemoryGame(3, 2), 6, 30, 2),
    "bandit_meta_3x4": (R.MetaEnv(R.UniformBernoulliBandits(3), 4), 6, 30, 0),
}


if __name__ == "__main__":
    import ctypes as C
    import zlib

    for name in CASES:
        cfg, E, T, slack = CASES[name]
        rng = np.random.default_rng(zlib.crc32(name.encode()))  # only used to MAKE the words; they are stored
        words = P.random_words(rng, E, 24 * (T + slack) + 64)
        A = O.lib().ro_env_num_actions(C.byref(O.make_env(P.oracle_cfg_for(cfg))))
        actions = rng.integers(0, A, size=(T + slack, E), dtype=np.uint8)
        ref = P.oracle_rollout(cfg, E, T, slack, actor_kind=O.ACTOR_REPLAY, actions=actions, env_words=words)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), words=words, actions=actions, slack=np.int64(slack),
                            min_steps=np.int64(T), obs=ref["obs"], next_obs=ref["next_obs"], action=ref["action"],
                            reward=ref["reward"], succ=ref["succ"], lane_len=ref["lane_len"])
        print(name, "steps", int(ref["lane_len"].sum()))
