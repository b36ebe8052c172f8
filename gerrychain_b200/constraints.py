"""Population constraint of the ReCom path.

Reference: /root/reference/gerrychain/constraints/validity.py:8-60 (Validator), :63-93
(within_percent_of_ideal_population), :96-122 (deviation_from_ideal) and
constraints/bounds.py:5-35 (Bounds).  Same names, arguments and error behaviour; the
``Bounds`` returned by ``within_percent_of_ideal_population`` additionally remembers
``pop_key`` so that the batched runner can hand the very same bounds to the device
(RcConfig.pop_lower / pop_upper).
"""

from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import numpy


class Bounds:
    """``lower <= min(func(partition))`` and ``max(func(partition)) <= upper``."""

    def __init__(self, func: Callable, bounds: Tuple[float, float]) -> None:
        self.func = func
        self.bounds = bounds
        self.pop_key = None  # set when built by within_percent_of_ideal_population

    def __call__(self, *args, **kwargs) -> bool:
        lower, upper = self.bounds
        values = self.func(*args, **kwargs)
        return lower <= min(values) and max(values) <= upper

    @property
    def __name__(self) -> str:
        return "Bounds({},{})".format(self.func.__name__, str(self.bounds))

    def __repr__(self) -> str:
        return "<{}>".format(self.__name__)


class Validator:
    """AND of boolean constraints; insists on real booleans (validity.py:32-56)."""

    def __init__(self, constraints: List[Callable]) -> None:
        self.constraints = constraints

    def __call__(self, partition) -> bool:
        for constraint in self.constraints:
            is_valid = constraint(partition)
            if isinstance(is_valid, numpy.bool_):
                is_valid = bool(is_valid)
            if is_valid is False:
                return False
            elif is_valid is True:
                pass
            else:
                raise TypeError("Constraint {} returned a non-boolean.".format(repr(constraint)))
        return True

    def __repr__(self) -> str:
        names = [constraint.__name__ for constraint in self.constraints]
        return f"Validator(constraints={names})"


def within_percent_of_ideal_population(initial_partition, percent: float = 0.01,
                                       pop_key: str = "population") -> Bounds:
    """Bounds fixed from the INITIAL partition: ``((1-p)*ideal, (1+p)*ideal)`` with
    ``ideal = total / number_of_districts`` (validity.py:63-93)."""

    def population(partition):
        return partition[pop_key].values()

    number_of_districts = len(initial_partition[pop_key].keys())
    total_population = sum(initial_partition[pop_key].values())
    ideal_population = total_population / number_of_districts
    bounds = ((1 - percent) * ideal_population, (1 + percent) * ideal_population)
    b = Bounds(population, bounds=bounds)
    b.pop_key = pop_key
    return b


def deviation_from_ideal(partition, attribute: str = "population") -> Dict:
    """``(value - ideal) / ideal`` per district (validity.py:96-122)."""
    number_of_districts = len(partition[attribute].keys())
    total = sum(partition[attribute].values())
    ideal = total / number_of_districts
    return {part: (value - ideal) / ideal for part, value in partition[attribute].items()}


def contiguous(partition) -> bool:
    """Are the districts connected?  (constraints/contiguity.py:169-181.)  Like the reference
    only the districts the last move touched are examined (contiguity.py:141-166
    ``affected_parts``); the pieces of every district are counted on the device
    (``rc_engine_contiguity``: union-find over the edges that are not cut)."""
    pieces = partition._district_pieces()
    flips, parent = partition.flips, partition.parent
    if flips is None:
        return bool((pieces <= 1).all())
    touched = set(flips.values())
    if parent is not None:
        touched.update(parent.assignment.mapping[node] for node in flips)
    rank = partition._rank
    return all(pieces[rank[part]] <= 1 for part in touched)


contiguous_bfs = contiguous


def number_of_contiguous_parts(partition) -> int:
    """How many districts are connected (constraints/contiguity.py:207-216)."""
    return int((partition._district_pieces() == 1).sum())
