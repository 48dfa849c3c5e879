"""Host logic of the logging bridge and the training-loop configuration (no GPU)."""
import pytest

from relearn_b200.logging import DisplayLogger, HistoryLogger, NullLogger
from relearn_b200.simulation import HistoryDataBound, TrainParallelConfig


def test_scoped_ids_follow_the_reference_layout():
    log = HistoryLogger()
    sim = log.with_scope("sim")
    sim.with_scope("ep").with_scope("fbk").with_scope("reward").log_scalar("mean", 1.5)
    sim.with_scope("step").log_counter_increment("count", 7)
    sim.with_scope("step").log_counter_increment("count", 3)
    sim.log_duration("time", 0.25)
    assert log.scalars == {"sim/ep/fbk/reward/mean": [1.5]}
    assert log.counters == {"sim/step/count": 10}
    assert log.durations == {"sim/time": 0.25}
    NullLogger().with_scope("x").log_scalar("y", 1.0)  # swallows everything
    with pytest.raises(ValueError):
        log.log("a", "histogram", 1)


def test_display_logger_chunks_by_counter():
    lines = []
    log = DisplayLogger("agent_update/count", every=2, out=lines.append)
    for i in range(4):
        log.log_scalar("sim/ep/length_mean", 10.0 * (i + 1))
        log.with_scope("agent_update").log_counter_increment("count", 1)
    heads = [ln for ln in lines if ln.startswith("====")]
    assert len(heads) == 2 and "agent_update/count = 2" in heads[0] and "agent_update/count = 4" in heads[1]
    means = [ln for ln in lines if ln.startswith("sim/ep/length_mean")]
    assert means[0].startswith("sim/ep/length_mean: 15") and means[1].startswith("sim/ep/length_mean: 35")


def test_worker_update_size_follows_train_parallel():
    # train.rs:111-118: min_update_size().divide(num_threads).max(min_worker_steps)
    cfg = TrainParallelConfig(num_periods=10, num_threads=8, min_worker_steps=10_000)
    assert HistoryDataBound(1, 0).divide(cfg.num_threads).max(HistoryDataBound(cfg.min_worker_steps, 0)) == HistoryDataBound(10_000, 0)
    assert HistoryDataBound(10_000, 100).divide(3).max(HistoryDataBound(0, 0)) == HistoryDataBound(3334, 100)
