"""Compactness scores (the reference's metrics/compactness.py:1-41): host arithmetic on the
``area`` and ``perimeter`` updaters of a partition."""

from __future__ import annotations

import math
from typing import Dict


def compute_polsby_popper(area: float, perimeter: float) -> float:
    """``4 pi area / perimeter^2``; nan for a district without perimeter."""
    try:
        return 4 * math.pi * area / perimeter**2
    except ZeroDivisionError:
        return math.nan


def polsby_popper(partition) -> Dict:
    """Polsby-Popper score of every district."""
    return {part: compute_polsby_popper(partition["area"][part], partition["perimeter"][part])
            for part in partition.parts}
