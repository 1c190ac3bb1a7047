"""Deterministic sampling of a flat gradient (shared by make_golden_grads.py and the tests)."""
import numpy as np

MAX_SAMPLE = 4096


def sample_idx(n: int) -> np.ndarray:
    if n <= MAX_SAMPLE:
        return np.arange(n)
    return (np.arange(MAX_SAMPLE, dtype=np.int64) * n) // MAX_SAMPLE
