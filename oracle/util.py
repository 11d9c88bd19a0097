"""Oracle restatement of /root/reference/vbsky/util.py:13-71 (test infrastructure)."""
from typing import NamedTuple

import numpy as np


class TipData(NamedTuple):
    """util.py:13-15."""

    partials: np.ndarray
    counts: np.ndarray


def order_events(times, node_heights, node_is_sample):
    """util.py:25-71.  jnp.argsort is a stable sort (lax.sort is_stable=True), hence kind='stable'."""
    times = np.asarray(times, dtype=np.float64)
    node_heights = np.asarray(node_heights, dtype=np.float64)
    node_is_sample = np.asarray(node_is_sample)
    assert len(node_heights) == len(node_is_sample)
    i = node_heights.argsort(kind="stable")
    node_heights = node_heights[i]
    node_is_sample = node_is_sample[i]
    merged = np.concatenate([times, node_heights])
    j = merged.argsort(kind="stable")
    delta = np.concatenate([np.zeros_like(times), 2 * node_is_sample.astype(np.int64) - 1])
    return np.array([merged[j], np.cumsum(delta[j])]).T
