"""Named wall-clock timers with NVTX ranges (reference: nvblox_torch/timer.py)."""
from __future__ import annotations

import time
from typing import Any, Literal

try:    # NVTX through torch when a CUDA build is present; otherwise timing only
    import torch.cuda.nvtx as _nvtx
except Exception:    # pragma: no cover
    _nvtx = None

_timers = {}


class Accumulator:
    """Running sum / count / last value of a named timer."""

    def __init__(self) -> None:
        self._n = 0
        self._sum = 0.0
        self._last = 0.0

    def mean(self) -> float:
        return self._sum / self._n if self._n else -1

    def last(self) -> float | None:
        return self._last

    def num_samples(self) -> int:
        return self._n

    def sum(self) -> float:
        return self._sum

    def accumulate(self, value: float) -> None:
        self._sum += value
        self._n += 1
        self._last = value


class Timer:
    """`with Timer('name'):` or `t = Timer('name'); ...; t.stop()`."""

    def __init__(self, name: str):
        self._name = name
        self._pushed = False
        if _nvtx is not None:
            try:
                _nvtx.range_push(name)
                self._pushed = True
            except Exception:
                self._pushed = False
        self._t0 = time.perf_counter()

    def __enter__(self) -> Timer:
        return self

    def stop(self) -> float:
        dt = time.perf_counter() - self._t0
        if self._pushed:
            _nvtx.range_pop()
            self._pushed = False
        _timers.setdefault(self._name, Accumulator()).accumulate(dt)
        return dt

    def __exit__(self, exc_type: Any, exc: Any, tb: Any) -> Literal[False]:
        self.stop()
        return False


def get_last_time(timer_name: str) -> float | None:
    return _timers[timer_name].last() if timer_name in _timers else 0.0


def get_mean_time(timer_name: str) -> float:
    return _timers[timer_name].mean() if timer_name in _timers else 0


def timer_status_string() -> str:
    if not _timers:
        return ''
    width = max(len(n) for n in _timers) + 2
    lines = ['', f"{'Timer name':<{width}}{'Mean[ms]':<20}{'Total[s]':<20}{'Num':<20}", '-' * 80]
    for name, acc in sorted(_timers.items()):
        lines.append(f'{name:<{width}}{1000 * acc.mean():<20.3}{acc.sum():<20.3}{acc.num_samples():<20}')
    return '\n'.join(lines) + '\n'


def print_timers() -> None:
    print(timer_status_string())
