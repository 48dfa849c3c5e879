"""Data-parallel host logic: how lanes, summaries and update sums are split over ranks.

relearn has no multi-process code; its only parallelism is `train_parallel`'s thread fan-out
(src/simulation/train.rs:98-158): every worker owns a buffer and forked RNGs, worker summaries are merged
with Chan's parallel variance (src/utils/stats.rs:184-209) and the update runs once over all buffers.
Across GPUs the same structure holds with ranks in place of threads:

* lanes shard by contiguous global index range and need no exchange (Philox noise is keyed by the
  *global* lane id, so a lane's trajectory does not depend on the rank that runs it);
* per-rank `StepsSummary` partials merge exactly like worker summaries;
* every full-batch reduction of the update is a *sum* on each rank, all-reduced, then divided by the
  global step count (ranks may hold different numbers of valid steps).

These helpers are pure host code (numpy only) so that the N>1 logic is testable with `gloo` on CPU; on
GPUs the all-reduce itself is `rl_ctx_allreduce_f64` (NCCL) inside the update kernels' launch sequence.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def shard_lanes(num_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """(count, global offset) of the contiguous lane range owned by `rank`; remainders go to low ranks."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(num_envs, world_size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return count, offset


@dataclass
class MeanVar:
    """OnlineMeanVariance (src/utils/stats.rs:121-213) as plain numbers."""

    mean: float = 0.0
    squared_residual_sum: float = 0.0
    count: int = 0

    def variance(self):
        return self.squared_residual_sum / self.count if self.count else None

    def merge(self, other: "MeanVar") -> "MeanVar":
        """`impl Add for OnlineMeanVariance` (stats.rs:184-209), Chan et al. pairwise update."""
        if other.count == 0:
            return MeanVar(self.mean, self.squared_residual_sum, self.count)
        if self.count == 0:
            return MeanVar(other.mean, other.squared_residual_sum, other.count)
        n = self.count + other.count
        delta = other.mean - self.mean
        mean = self.mean + delta * (other.count / n)
        m2 = self.squared_residual_sum + other.squared_residual_sum + delta * delta * (self.count * other.count / n)
        return MeanVar(mean, m2, n)

    def to_array(self) -> np.ndarray:
        return np.array([self.mean, self.squared_residual_sum, float(self.count)], np.float64)

    @staticmethod
    def from_array(a) -> "MeanVar":
        return MeanVar(float(a[0]), float(a[1]), int(a[2]))


def merge_summaries(parts) -> list:
    """Merge per-rank [step_reward, episode_reward, episode_length] MeanVar triples (train.rs:153-156)."""
    out = [MeanVar(), MeanVar(), MeanVar()]
    for p in parts:
        out = [a.merge(b) for a, b in zip(out, p)]
    return out


def global_mean_from_sums(local_sums: np.ndarray, local_count: float, all_reduce_sum) -> np.ndarray:
    """mean over the global batch from per-rank sums: all-reduce [sums..., count] once, divide after.

    `all_reduce_sum(array) -> array` is the collective (gloo in tests, NCCL in the library).  This is the
    contract of the update kernels: partial rows hold sums, `sums[P + SC_COUNT]` holds the valid-step count.
    """
    packed = np.concatenate([np.asarray(local_sums, np.float64).ravel(), [float(local_count)]])
    total = np.asarray(all_reduce_sum(packed), np.float64)
    if total[-1] == 0:
        raise ZeroDivisionError("empty global batch")
    return total[:-1] / total[-1]
