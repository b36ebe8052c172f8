"""Constraints of the ReCom path.

Reference: /root/reference/gerrychain/constraints/validity.py:8-60 (Validator), :63-93
(within_percent_of_ideal_population), :96-122 (deviation_from_ideal) and
constraints/bounds.py:5-35 (Bounds).  Same names, arguments and error behaviour; the
``Bounds`` returned by ``within_percent_of_ideal_population`` additionally remembers
``pop_key`` so that the batched runner can hand the very same bounds to the device
(RcConfig.pop_lower / pop_upper).

The scalar bound wrappers users put next to it in a ReCom chain - ``UpperBound`` on the
number of cut edges in the reference's own ReCom walkthrough (docs/user/recom.rst), the
self-configuring ones, the Polsby-Popper norms - are plain host callables evaluated by
``MarkovChain`` on every proposed state (constraints/bounds.py:36-215,
constraints/compactness.py, validity.py:125-186).
"""

from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

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


class _ScalarBound:
    """Shared plumbing of the single-number wrappers below: a callable score plus the
    ``__name__`` the reference prints in ``MarkovChain``'s error message."""

    _label = "{cls}({func})"

    def __init__(self, func: Callable) -> None:
        self.func = func

    @property
    def __name__(self) -> str:
        return self._label.format(cls=type(self).__name__, func=self.func.__name__,
                                  bound=getattr(self, "bound", None))

    def __repr__(self) -> str:
        return "<{}>".format(self.__name__)


class UpperBound(_ScalarBound):
    """``func(partition) <= bound`` (bounds.py:36-71; the name string keeps the reference's
    ``>=``)."""

    _label = "{cls}({func} >= {bound})"

    def __init__(self, func: Callable, bound: float) -> None:
        super().__init__(func)
        self.bound = bound

    def __call__(self, *args, **kwargs) -> bool:
        return self.func(*args, **kwargs) <= self.bound


class LowerBound(_ScalarBound):
    """``func(partition) >= bound`` (bounds.py:74-109)."""

    _label = "{cls}({func} <= {bound})"

    def __init__(self, func: Callable, bound: float) -> None:
        super().__init__(func)
        self.bound = bound

    def __call__(self, *args, **kwargs) -> bool:
        return self.func(*args, **kwargs) >= self.bound


class SelfConfiguringUpperBound(_ScalarBound):
    """Upper bound fixed by the first partition it sees (bounds.py:112-147).  As in the
    reference a falsy bound (``None`` or 0) is configured again on the next call."""

    def __init__(self, func: Callable) -> None:
        super().__init__(func)
        self.bound: Optional[float] = None

    def __call__(self, partition) -> bool:
        if not self.bound:
            self.bound = self.func(partition)
        return self.func(partition) <= self.bound


class SelfConfiguringLowerBound(_ScalarBound):
    """Lower bound = first value seen minus ``epsilon`` (bounds.py:150-190)."""

    def __init__(self, func: Callable, epsilon: float = 0.05) -> None:
        super().__init__(func)
        self.bound: Optional[float] = None
        self.epsilon = epsilon

    def __call__(self, partition) -> bool:
        if not self.bound:
            self.bound = self.func(partition) - self.epsilon
        return self.func(partition) >= self.bound


class WithinPercentRangeOfBounds(_ScalarBound):
    """Stay within ``percent`` (0-100) of the first value seen; the configuring call itself
    answers True (bounds.py:193-236)."""

    def __init__(self, func: Callable, percent: float) -> None:
        super().__init__(func)
        self.percent = float(percent) / 100.0
        self.lbound: Optional[float] = None
        self.ubound: Optional[float] = None

    def __call__(self, partition) -> bool:
        if self.lbound and self.ubound:
            return self.lbound <= self.func(partition) <= self.ubound
        first = self.func(partition)
        self.lbound, self.ubound = first * (1.0 - self.percent), self.func(partition) * (1.0 + self.percent)
        return True


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


def districts_within_tolerance(partition, attribute_name: str = "population", percentage: float = 0.1) -> bool:
    """``max - min <= percentage * min`` over the districts' values; a percentage of 1 or more
    is read as percent (validity.py:125-153)."""
    if percentage >= 1:
        percentage *= 0.01
    values = partition[attribute_name].values()
    smallest = min(values)
    return max(values) - smallest <= percentage * smallest


def refuse_new_splits(partition_county_field: str) -> Callable:
    """Reject a state that splits a county which was whole before; ``partition_county_field``
    names a ``county_splits`` updater (validity.py:156-173)."""
    from .updaters import CountySplit

    def _refuse_new_splits(partition) -> bool:
        return all(info.split != CountySplit.NEW_SPLIT for info in partition[partition_county_field].values())

    return _refuse_new_splits


def no_vanishing_districts(partition) -> bool:
    """No district may be emptied by a move (validity.py:176-186)."""
    if not partition.parent:
        return True
    return all(len(members) > 0 for members in partition.assignment.parts.values())


# ---- Polsby-Popper norms (constraints/compactness.py) --------------------------------------
def L1_reciprocal_polsby_popper(partition) -> float:
    return sum(1 / score for score in partition["polsby_popper"].values())


def L1_polsby_popper(partition) -> float:
    return sum(score for score in partition["polsby_popper"].values())


def L2_polsby_popper(partition) -> float:
    return math.sqrt(sum(score**2 for score in partition["polsby_popper"].values()))


def L_minus_1_polsby_popper(partition) -> float:
    return len(partition.parts) / sum(1 / score for score in partition["polsby_popper"].values())


no_worse_L_minus_1_polsby_popper = SelfConfiguringLowerBound(L_minus_1_polsby_popper)
no_worse_L1_reciprocal_polsby_popper = SelfConfiguringUpperBound(L1_reciprocal_polsby_popper)


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
# constraints/contiguity.py:184-204 answers the same question by a local search around the
# flipped nodes, valid when the parent was contiguous; the device count needs no such premise
single_flip_contiguous = contiguous


def number_of_contiguous_parts(partition) -> int:
    """How many districts are connected (constraints/contiguity.py:207-216)."""
    return int((partition._district_pieces() == 1).sum())


# constraints/contiguity.py:220-223: never fewer connected districts than at the start
no_more_discontiguous = SelfConfiguringLowerBound(number_of_contiguous_parts)
