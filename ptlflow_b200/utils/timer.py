"""Elapsed-time counter with the interface the reference's scripts use (ptlflow/utils/timer.py:25-128):
``tic`` / ``toc`` bracket a region with ``torch.cuda.synchronize()`` on both sides and add it to the total;
``reset`` zeroes the total; ``total()`` is in seconds, ``mean()`` divides by the number of closed intervals (as the
reference does, by ``max(1, tocs - 1)``: its first interval is treated as warm-up)."""
from __future__ import annotations

import time

import torch


class Timer:
    def __init__(self, name: str, indent_level: int = 0) -> None:
        self.name, self.indent_level = name, indent_level
        self.num_tocs = 0
        self.num_global_tocs = 0
        self.has_tic = False
        self.reset()

    def reset(self) -> None:
        self.total_time = 0.0

    @staticmethod
    def _sync() -> None:
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def tic(self) -> None:
        self.has_tic = True
        self._sync()
        self.start = time.perf_counter()

    def toc(self) -> None:
        self._sync()
        self.end = time.perf_counter()
        assert self.has_tic, "toc called without tic"
        self.total_time += self.end - self.start
        self.has_tic = False
        self.num_tocs += 1

    def total(self) -> float:
        return self.total_time

    def mean(self) -> float:
        n = self.num_global_tocs if self.num_global_tocs > 0 else self.num_tocs
        return self.total() / max(1, n - 1)

    def __repr__(self) -> str:
        return f'{"  " * self.indent_level}{self.name}: {1000 * self.total():.1f} ({1000 * self.mean():.1f}) ms'
