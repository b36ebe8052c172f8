"""``MarkovChain`` with the reference's iterator contract, plus the batched runner.

Reference: /root/reference/gerrychain/chain.py:33-221.

* ``MarkovChain`` is the reference's loop, statement for statement in behaviour
  (chain.py:166-196): the first yield is the initial state and counts as a step; a proposal
  that fails ``is_valid`` is discarded without counting; a valid one counts whether or not
  ``accept`` takes it.  Any Python ``proposal`` / constraint / ``accept`` works; with
  ``gerrychain_b200.recom`` as the proposal every step is one kernel launch.
* ``BatchedMarkovChain`` is what chain.py gains: ``n_chains`` independent chains resident
  on one GPU, advanced together by the persistent chain-step kernel.  The proposal must be
  ``recom`` (a ``functools.partial`` of it or a ``ReCom``), the constraints population
  ``Bounds`` from ``within_percent_of_ideal_population`` and ``accept`` ``always_accept`` -
  those are the pieces that exist on the device; anything else raises ``TypeError`` rather
  than silently running on the host.
"""

from __future__ import annotations

import functools
import random
import warnings
from typing import Callable, Iterable, Optional, Union

import numpy as np

from . import _abi
from .accept import always_accept
from .config import EngineConfig
from .constraints import Bounds, Validator
from .partition import Partition
from .proposals import (ReCom, _WARNING_TEXT, engine_kwargs, raise_for_status, recom, reversible_engine_kwargs,
                        reversible_recom)
from .tree import BipartitionWarning, bipartition_tree


class MarkovChain:
    """chain.py:33."""

    def __init__(self, proposal: Callable, constraints: Union[Iterable[Callable], Validator, Callable],
                 accept: Callable, initial_state: Partition, total_steps: int) -> None:
        is_valid = Validator([constraints]) if callable(constraints) else Validator(constraints)
        if not is_valid(initial_state):
            failed = [c for c in is_valid.constraints if not c(initial_state)]
            raise ValueError("The given initial_state is not valid according is_valid. "
                             "The failed constraints were: " + ",".join([f.__name__ for f in failed]))
        self.proposal = proposal
        self.is_valid = is_valid
        self.accept = accept
        self.total_steps = total_steps
        self.initial_state = initial_state
        self.state = initial_state

    @property
    def constraints(self) -> Validator:
        return self.is_valid

    @constraints.setter
    def constraints(self, constraints) -> None:
        is_valid = Validator([constraints]) if callable(constraints) else Validator(constraints)
        if not is_valid(self.initial_state):
            failed = [c for c in is_valid.constraints if not c(self.initial_state)]
            raise ValueError("The given initial_state is not valid according to the new constraints. "
                             "The failed constraints were: " + ",".join([f.__name__ for f in failed]))
        self.is_valid = is_valid

    def __iter__(self) -> "MarkovChain":
        self.counter = 0
        self.state = self.initial_state
        return self

    def __next__(self) -> Optional[Partition]:
        if self.counter == 0:
            self.counter += 1
            return self.state
        while self.counter < self.total_steps:
            proposed_next_state = self.proposal(self.state)
            if self.state is not None:
                self.state.parent = None  # chain.py:188-189: at most two generations alive
            if self.is_valid(proposed_next_state):
                if self.accept(proposed_next_state):
                    self.state = proposed_next_state
                self.counter += 1
                return self.state
        raise StopIteration

    def __len__(self) -> int:
        return self.total_steps

    def __repr__(self) -> str:
        return "<MarkovChain [{} steps]>".format(len(self))

    def with_progress_bar(self):
        from tqdm.auto import tqdm

        return tqdm(self)


# ------------------------------------------------------------------------------------------
def _recom_arguments(proposal):
    """(pop_col, pop_target, epsilon, node_repeats, region_surcharge, method) of a recom
    proposal object, or TypeError."""
    if isinstance(proposal, ReCom):
        return proposal.pop_col, proposal.ideal_pop, proposal.epsilon, 1, None, proposal.method
    if isinstance(proposal, functools.partial) and proposal.func is reversible_recom and not proposal.args:
        kw = proposal.keywords
        try:
            return (kw["pop_col"], kw["pop_target"], kw["epsilon"], 1, None,
                    ("reversible", kw.get("M", 1), kw.get("repeat_until_valid", False)))
        except KeyError as e:
            raise TypeError(f"reversible_recom partial is missing the keyword {e}") from None
    if isinstance(proposal, functools.partial) and proposal.func is recom and not proposal.args:
        kw = proposal.keywords
        try:
            return (kw["pop_col"], kw["pop_target"], kw["epsilon"], kw.get("node_repeats", 1),
                    kw.get("region_surcharge"), kw.get("method", bipartition_tree))
        except KeyError as e:
            raise TypeError(f"recom partial is missing the keyword {e}") from None
    raise TypeError("BatchedMarkovChain runs gerrychain_b200.recom on the device: pass "
                    "functools.partial(recom, pop_col=..., pop_target=..., epsilon=...) or a ReCom")


def _device_bounds(constraints, pop_col):
    """(constraints, lower, upper): the population bounds the device enforces.  ``contiguous``
    (the constraint of the reference's ReCom walkthrough) is accepted as well: both sides of
    a tree cut are connected, so every ReCom child of a contiguous plan is contiguous and the
    constraint only has to hold - and is checked - for the initial state."""
    from .constraints import contiguous

    cons = constraints.constraints if isinstance(constraints, Validator) else (
        [constraints] if callable(constraints) else list(constraints))
    lower, upper = -np.inf, np.inf
    for c in cons:
        if c is contiguous:  # == contiguous_bfs == single_flip_contiguous
            continue
        if not isinstance(c, Bounds) or c.pop_key is None:
            raise TypeError(f"constraint {c!r} has no device implementation: BatchedMarkovChain takes "
                            "within_percent_of_ideal_population(...) bounds and contiguous only")
        lower, upper = max(lower, c.bounds[0]), min(upper, c.bounds[1])
    return cons, lower, upper


class ChainBatch:
    """State of all resident chains after a batched step: a VIEW of the engine's device
    buffers, valid until the chain advances again (call ``.partition(i)`` or ``.cpu()``
    values to keep something)."""

    def __init__(self, runner: "BatchedMarkovChain", step: int):
        self._r = runner
        self.step = step

    @property
    def assignment(self):
        """[n_chains, N] district index per node, on the device (torch.uint8; torch.int16 for
        plans with more than 255 districts)."""
        return self._r.engine.assignment[:, : self._r.engine.csr.n_nodes]

    def cut_edge_counts(self) -> np.ndarray:
        """``len(partition["cut_edges"])`` for every chain."""
        return self._r.engine.cut_count.cpu().numpy().copy()

    def population(self) -> np.ndarray:
        """int64 [n_chains, k]: ``partition["population"]`` values, columns in sorted district order."""
        return self._r.engine.tally.cpu().numpy().copy()

    def tally(self, field) -> np.ndarray:
        """int64 [n_chains, k]: ``Tally(field)`` for every chain (any integer node attribute
        or list of attributes, e.g. vote counts), computed on the device from the resident plans."""
        col = self._r.initial_state._ctx.column(field)
        return self._r.engine.tally_column(col).cpu().numpy()

    def region_splits(self, reg_attr: str) -> np.ndarray:
        """int32 [n_chains, 3]: per chain the number of ``reg_attr`` regions with a cut edge
        inside them (``tally_region_splits``), the number spread over more than one district,
        and the total (region, district) pieces (``rc_engine_region_splits``)."""
        col, n_regions = self._r.initial_state._ctx.region_codes(reg_attr)
        return self._r.engine.region_splits(col, n_regions).cpu().numpy()

    def district_pieces(self) -> np.ndarray:
        """int32 [n_chains, k]: connected pieces of every district (``rc_engine_contiguity``)."""
        return self._r.engine.contiguity().cpu().numpy()

    def contiguous(self) -> np.ndarray:
        """bool [n_chains]: ``constraints.contiguous`` of every chain's plan."""
        return (self.district_pieces() <= 1).all(axis=1)

    def election(self, election) -> np.ndarray:
        """int64 [n_chains, parties, k]: the vote totals of an ``Election`` for every chain."""
        return np.stack([self.tally(col) for col in election.columns], axis=1)

    @property
    def district_labels(self):
        return list(self._r.initial_state._district_labels)

    def partition(self, i: int) -> Partition:
        """Host ``Partition`` snapshot of chain ``i``."""
        e, init = self._r.engine, self._r.initial_state
        return init.__class__._from_device(
            init, e._labels_to_numpy(e.assignment[i, : e.csr.n_nodes]), (self._r.pop_col,),
            e.tally[i].cpu().numpy(), int(e.cut_count[i].item()),
            e.cut_mask[i].cpu().numpy().view(np.uint32).copy(), e.last_pair[i].cpu().numpy())

    def histograms(self, n_cut_bins=4096, n_dev_bins=64, dev_bin_width=0.001, all_reduce=True):
        """Ensemble histograms of cut-edge counts and max |population deviation|; summed over
        ranks with one NCCL all-reduce when torch.distributed is initialised."""
        cut, dev = self._r.engine.histograms(n_cut_bins, n_dev_bins, dev_bin_width)
        if all_reduce:
            from .distributed import all_reduce_histograms

            all_reduce_histograms(cut, dev)
        return cut, dev


class BatchedMarkovChain:
    """``n_chains`` ReCom chains advanced together on one GPU (this rank's shard)."""

    def __init__(self, proposal, constraints, accept, initial_state: Partition, total_steps: int,
                 n_chains: int, seed: Optional[int] = None, chain_offset: int = 0, device=None,
                 yield_every: int = 1, initial_assignments=None, **engine_options):
        from .accept import cut_edge_accept

        if accept is always_accept:
            accept_rule = _abi.RC_ACCEPT_ALWAYS
        elif accept is cut_edge_accept:
            accept_rule = _abi.RC_ACCEPT_CUT_EDGE
        else:
            raise TypeError("BatchedMarkovChain implements accept=always_accept and accept=cut_edge_accept "
                            "on the device; other acceptance functions need the host loop (MarkovChain)")
        (self.pop_col, pop_target, epsilon, node_repeats, self.region_surcharge,
         method) = _recom_arguments(proposal)
        cons, lower, upper = _device_bounds(constraints, self.pop_col)
        for c in cons:
            key = getattr(c, "pop_key", None)
            if key is not None and initial_state.updaters.get(key) is None:
                raise KeyError(key)
        if initial_assignments is not None and any(not isinstance(c, Bounds) for c in cons):
            raise TypeError("contiguous is checked on initial_state only: with initial_assignments, verify the "
                            "plans yourself (ChainBatch.contiguous()) and pass the population bounds alone")
        if not Validator(cons)(initial_state):
            failed = [c for c in cons if not c(initial_state)]
            raise ValueError("The given initial_state is not valid according is_valid. "
                             "The failed constraints were: " + ",".join([f.__name__ for f in failed]))
        self.initial_state = initial_state
        self.total_steps = int(total_steps)
        self.n_chains = int(n_chains)
        self.yield_every = max(1, int(yield_every))
        if isinstance(method, tuple) and method[0] == "reversible":
            self._ekw = reversible_engine_kwargs(pop_target, epsilon, method[1], method[2])
        else:
            self._ekw = engine_kwargs(pop_target, epsilon, node_repeats, self.region_surcharge, method)
        self._ekw.update(engine_options)
        if np.isfinite(lower) and np.isfinite(upper):
            self._ekw.update(pop_lower=float(lower), pop_upper=float(upper))
        self._ekw["seed"] = random.getrandbits(64) if seed is None else int(seed)
        self._ekw["chain_offset"] = int(chain_offset)
        self._ekw["accept_rule"] = accept_rule
        self._device = device
        self._plans = initial_assignments
        self.engine = None

    def _build(self):
        from .engine import Engine

        init = self.initial_state
        csr = init._ctx.csr(self.pop_col, self.region_surcharge)
        cfg = EngineConfig(n_districts=len(init), n_chains=self.n_chains, **self._ekw)
        plans = init._vec if self._plans is None else np.asarray(self._plans, dtype=init._vec.dtype)
        self.engine = Engine(csr, cfg, plans, device=self._device)
        self._warned = False

    def __iter__(self):
        if self.engine is None:
            self._build()
        else:
            # a second pass starts from the initial plans again but NOT from the same random
            # numbers: the chains keep their position in the Philox streams, as the reference's
            # chain keeps drawing from Python's global `random` when it is iterated again
            position = self.engine.proposal_no.clone()
            self.engine.set_assignment(self.initial_state._vec if self._plans is None else self._plans)
            self.engine.proposal_no.copy_(position)
        self.counter = 0
        return self

    def __len__(self):
        return self.total_steps

    def __repr__(self):
        return "<BatchedMarkovChain [{} chains x {} steps]>".format(self.n_chains, self.total_steps)

    def __next__(self) -> ChainBatch:
        if self.counter == 0:
            self.counter = 1
            return ChainBatch(self, 1)
        if self.counter >= self.total_steps:
            raise StopIteration
        n = min(self.yield_every, self.total_steps - self.counter)
        self.engine.run(n)
        self.counter += n
        self._check()
        return ChainBatch(self, self.counter)

    def run_to_end(self) -> ChainBatch:
        """Advance all chains to ``total_steps`` in a single launch."""
        if self.engine is None or not hasattr(self, "counter"):
            iter(self)
        if self.counter == 0:
            self.counter = 1
        n = self.total_steps - self.counter
        if n > 0:
            self.engine.run(n)
            self.counter += n
        self._check()
        return ChainBatch(self, self.counter)

    def _check(self):
        e = self.engine
        status = e.status.cpu().numpy()
        bad = np.nonzero(status)[0]
        if len(bad):
            k = len(self.initial_state)
            raise_for_status(int(status[bad[0]]), self._ekw["max_attempts"], k * (k - 1) / 2)
        if not self._warned and int((e.flags & _abi.RC_FLAG_WARNED).max().item()):
            self._warned = True
            warnings.warn(_WARNING_TEXT.format(self._ekw["warn_attempts"]), BipartitionWarning)
