"""Synthetic instruction-pair batches — the input generator shared by bench.py, the parity tests and the
oracle.  Pure numpy, no model math: ids ~ U{3..V-1}, BOS first, EOS last, the prompt part masked to -100 in the labels,
seeded per (step, rank) so every rank and both sides of a parity test draw the same data."""
from __future__ import annotations

from typing import Tuple

import numpy as np

IGNORE_INDEX = -100


def synthetic_batch(step: int, rank: int, batch: int, seq_len: int, vocab: int) -> Tuple[np.ndarray, np.ndarray]:
    """Synthetic instruction pairs: ids ~ U{3..V-1}, BOS first, EOS last, prompt masked to -100."""
    rng = np.random.default_rng(2024 + rank * 1_000_003 + step)
    ids = rng.integers(3, vocab, size=(batch, seq_len), dtype=np.int64)
    ids[:, 0] = 1
    ids[:, -1] = 2
    lo, hi = max(1, seq_len // 32), max(2, seq_len // 2)
    plen = rng.integers(lo, hi + 1, size=(batch,))
    labels = ids.copy()
    for b in range(batch):
        labels[b, : plen[b]] = IGNORE_INDEX
    return ids.astype(np.int32), labels.astype(np.int32)
