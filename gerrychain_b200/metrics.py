"""Scores users compute from a state (the reference's metrics/): Polsby-Popper compactness
(metrics/compactness.py:1-41) on the ``area`` and ``perimeter`` updaters, and the partisan
scores of an ``ElectionResults`` (metrics/partisan.py).  Host arithmetic on values that
come from the device tallies."""

from __future__ import annotations

import math
from typing import Dict

import numpy as np


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


# ---- partisan scores (the reference's metrics/partisan.py) ----------------------------------
# Host arithmetic on an ElectionResults (per-district vote totals, which come from the device
# tallies).  A positive value favours the first party of the election in every score.

def _first_party_shares(election_results):
    return np.asarray(election_results.percents(election_results.election.parties[0]), dtype=np.float64)


def mean_median(election_results) -> float:
    """median - mean of the first party's district vote shares (partisan.py:14-29)."""
    shares = _first_party_shares(election_results)
    return float(np.median(shares) - np.mean(shares))


def mean_thirdian(election_results) -> float:
    """Share at one third of the sorted districts minus the mean share (partisan.py:32-54)."""
    shares = _first_party_shares(election_results)
    return float(np.sort(shares)[round(len(shares) / 3)] - np.mean(shares))


def wasted_votes(party1_votes, party2_votes):
    """(wasted by party 1, wasted by party 2) in one two-party race: the winner wastes what
    exceeds half the votes, the loser everything; a tie counts as a win of party 2
    (partisan.py:81-103)."""
    half = (party1_votes + party2_votes) / 2
    if party1_votes > party2_votes:
        return party1_votes - half, party2_votes
    return party1_votes, party2_votes - half


def efficiency_gap(election_results) -> float:
    """(votes wasted by party 2 - votes wasted by party 1) / all votes (partisan.py:57-78)."""
    first, second = (election_results.counts(party) for party in election_results.election.parties)
    gap = 0.0
    for a, b in zip(first, second):
        waste_a, waste_b = wasted_votes(a, b)
        gap += waste_b - waste_a
    return float(gap / election_results.total_votes())


def partisan_bias(election_results) -> float:
    """Fraction of districts above the first party's mean share, minus one half
    (partisan.py:106-125)."""
    shares = _first_party_shares(election_results)
    return float(np.count_nonzero(shares > shares.mean()) / len(shares) - 0.5)


def partisan_gini(election_results) -> float:
    """Area between the seats-votes curve (uniform swing) and its reflection about
    (1/2, 1/2), per seat (partisan.py:128-164)."""
    party = election_results.election.parties[0]
    overall = election_results.percent(party)
    curve = [overall - share + 0.5 for share in sorted(election_results.percents(party), reverse=True)]
    mirrored = [1 - s for s in reversed(curve)]
    return float(sum(abs(s - r) for s, r in zip(curve, mirrored)) / len(curve))
