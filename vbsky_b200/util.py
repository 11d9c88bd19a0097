"""Small containers shared by the host-side API (mirror of /root/reference/vbsky/util.py:13-15)."""
from typing import NamedTuple

import numpy as np


class TipData(NamedTuple):
    """partials [S, n, 4] 0/1 tip partials per site pattern, counts [S] pattern multiplicities."""

    partials: np.ndarray
    counts: np.ndarray
