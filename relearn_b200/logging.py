"""Statistics loggers for the training loop (src/logging/mod.rs:20-135, display.rs): the ids `train_device` writes are
the reference's own (`sim/ep/fbk/reward/mean`, `sim/step/count`, `agent_update/count`, ...), so a consumer of the
reference's logs reads these unchanged.  Three value kinds as in `LogValue`: scalars, counter increments, durations."""
from __future__ import annotations

import time

SCALAR, COUNTER, DURATION = "scalar", "counter", "duration"


class StatsLogger:
    """`StatsLogger` (logging/mod.rs:20-135): `log(id, kind, value)` plus the convenience methods of the trait."""

    def log(self, name: str, kind: str, value) -> None:  # pragma: no cover - interface
        raise NotImplementedError

    def flush(self) -> None:
        pass

    def log_scalar(self, name: str, value: float) -> None:
        self.log(name, SCALAR, float(value))

    def log_counter_increment(self, name: str, increment: int) -> None:
        self.log(name, COUNTER, int(increment))

    def log_duration(self, name: str, seconds: float) -> None:
        self.log(name, DURATION, float(seconds))

    def with_scope(self, scope: str) -> "ScopedLogger":
        return ScopedLogger(scope, self)


class ScopedLogger(StatsLogger):
    """`ScopedLogger` (logging/mod.rs:75-80): prefixes every id with `scope/`."""

    def __init__(self, scope: str, inner: StatsLogger):
        self.scope, self.inner = scope, inner

    def log(self, name, kind, value):
        self.inner.log(f"{self.scope}/{name}", kind, value)

    def flush(self):
        self.inner.flush()


class NullLogger(StatsLogger):
    """`()` as a logger (logging/mod.rs:339-356)."""

    def log(self, name, kind, value):
        pass


class HistoryLogger(StatsLogger):
    """Keeps everything: `scalars[id]` = list of values in log order, `counters[id]` = running total,
    `durations[id]` = total seconds.  What the tests and `bench.py` read."""

    def __init__(self):
        self.scalars: dict[str, list[float]] = {}
        self.counters: dict[str, int] = {}
        self.durations: dict[str, float] = {}

    def log(self, name, kind, value):
        if kind == SCALAR:
            self.scalars.setdefault(name, []).append(value)
        elif kind == COUNTER:
            self.counters[name] = self.counters.get(name, 0) + value
        elif kind == DURATION:
            self.durations[name] = self.durations.get(name, 0.0) + value
        else:
            raise ValueError(f"unknown log value kind {kind!r}")


class DisplayLogger(HistoryLogger):
    """`DisplayLogger::new(ByCounter::of_path(path, n))` (logging/display.rs, chunk.rs): prints a summary of everything
    logged since the last one every `n` increments of the counter `path` (e.g. `agent_update/count`)."""

    def __init__(self, counter_path: str = "agent_update/count", every: int = 10, out=print):
        super().__init__()
        self.counter_path, self.every, self.out = counter_path, every, out
        self._since: dict[str, list[float]] = {}
        self._start = time.perf_counter()

    def log(self, name, kind, value):
        super().log(name, kind, value)
        if kind == SCALAR:
            self._since.setdefault(name, []).append(value)
        if kind == COUNTER and name == self.counter_path and self.counters[name] % self.every == 0:
            self.flush()

    def flush(self):
        if not self._since and not self.counters:
            return
        self.out(f"==== {self.counter_path} = {self.counters.get(self.counter_path, 0)} "
                 f"({time.perf_counter() - self._start:.1f} s) ====")
        for name in sorted(self._since):
            vals = self._since[name]
            self.out(f"{name}: {sum(vals) / len(vals):.6g}  (n = {len(vals)})")
        for name in sorted(self.counters):
            self.out(f"{name}: {self.counters[name]}")
        self._since = {}
