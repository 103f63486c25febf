"""Minimal stand-in for ``gym.spaces.Box`` (gym is not installed in the target image).

Only the attributes the reference's callers read are provided: ``shape``, ``dtype``, ``low``, ``high``
(``rl_playground.py:68-69,121-122``), plus ``sample``/``contains`` for convenience.
"""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape, dtype=np.float64, seed=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low.real), self.low.real, -1.0)
        hi = np.where(np.isfinite(self.high.real), self.high.real, 1.0)
        x = self._rng.uniform(lo, hi)
        if np.issubdtype(self.dtype, np.complexfloating):
            x = x + 1j * self._rng.uniform(lo, hi)
        return x.astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        if x.shape != self.shape:
            return False
        if np.issubdtype(self.dtype, np.complexfloating):
            return True
        return bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.flat[0]}, {self.high.flat[0]}, {self.shape}, {self.dtype})"
