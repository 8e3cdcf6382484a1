"""perf — the reference's PerfStats surface for the local engine.

Mirrors reference pkg/utils/perf.go: StartTimer/StopTimer accumulate per-operation durations and call counts
(perf.go:64-121), TraceFunc returns the closer (perf.go:288), GetStats returns {"timers", "callCounts", "lastResetTime"}
(perf.go:296-320) and Reset clears them (perf.go:323-335).  Two deliberate differences, both called out in SURVEY.md §5.1:
timers are keyed per (thread, operation) while running, so concurrent requests do not overwrite each other's start time
(the reference keeps ONE start time per operation name: perf.go:64-80), and durations are integer nanoseconds — what Go's
time.Duration marshals to in the /api/perf/stats JSON (pkg/handlers/perf.go:12-25)."""
from __future__ import annotations

import datetime
import threading
import time


class PerfStats:
    def __init__(self, clock=time.perf_counter_ns):
        self._mu = threading.Lock()
        self._clock = clock
        self._start: dict[tuple[int, str], int] = {}
        self.timers: dict[str, int] = {}
        self.callCounts: dict[str, int] = {}
        self.lastResetTime = datetime.datetime.now(datetime.timezone.utc)

    def StartTimer(self, operation: str) -> None:
        with self._mu:
            self._start[(threading.get_ident(), operation)] = self._clock()

    def StopTimer(self, operation: str) -> int:
        """-> elapsed ns (0 if the timer was never started, as perf.go:87-93 does)"""
        now = self._clock()
        with self._mu:
            t0 = self._start.pop((threading.get_ident(), operation), None)
            if t0 is None:
                return 0
            d = now - t0
            self.timers[operation] = self.timers.get(operation, 0) + d
            self.callCounts[operation] = self.callCounts.get(operation, 0) + 1
            return d

    def RecordMetric(self, operation: str, duration_ns: int) -> None:
        with self._mu:
            self.timers[operation] = self.timers.get(operation, 0) + int(duration_ns)
            self.callCounts[operation] = self.callCounts.get(operation, 0) + 1

    def TraceFunc(self, operation: str):
        self.StartTimer(operation)
        return lambda: self.StopTimer(operation)

    def GetStats(self) -> dict:
        with self._mu:
            return {"timers": dict(self.timers), "callCounts": dict(self.callCounts),
                    "lastResetTime": self.lastResetTime.isoformat().replace("+00:00", "Z")}

    def Reset(self) -> None:
        with self._mu:
            self.timers.clear(); self.callCounts.clear(); self._start.clear()
            self.lastResetTime = datetime.datetime.now(datetime.timezone.utc)


_global = PerfStats()


def GetPerfStats() -> PerfStats:
    """process-wide instance, as utils.GetPerfStats() (perf.go:38-47)"""
    return _global
