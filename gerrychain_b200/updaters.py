"""Updaters of the ReCom path: ``Tally`` and ``cut_edges``, and the statistics users read
behind every state: ``Election`` vote tallies, ``tally_region_splits`` and the compactness
updaters of ``GeographicPartition`` (perimeter, boundaries).

Reference: /root/reference/gerrychain/updaters/tally.py:69-170 and
updaters/cut_edges.py:99-120.  The reference recomputes both incrementally on the host
from parent/child flows; here both are maintained by the chain-step kernel (per-district
sums and a cut-edge bitmask resident in HBM) and these callables only present the device
values in the reference's shapes (``{part: sum}`` and ``{(u, v), ...}``).
"""

from __future__ import annotations

import collections
import math
import warnings
from enum import Enum
from typing import Dict, List, Optional, Type, Union

import numpy as np


class Tally:
    """``Tally(fields, alias=None, dtype=int)`` - per-district sum of node attribute(s)."""

    __slots__ = ["fields", "alias", "dtype"]

    def __init__(self, fields: Union[str, List[str]], alias: Optional[str] = None, dtype: Type = int):
        if not isinstance(fields, list):
            fields = [fields]
        if not alias:
            alias = fields[0]
        self.fields = fields
        self.alias = alias
        self.dtype = dtype

    def __call__(self, partition):
        sums = partition._tally(tuple(self.fields))
        return {part: self.dtype(v) for part, v in zip(partition._district_labels, sums)}


class DataTally:
    """``DataTally(data, alias)`` - per-district sum of numbers that need not be node attributes
    nor integers: a ``{node: value}`` mapping, a pandas Series indexed by node, or the name
    of a node attribute (updaters/tally.py:12-66).  Summed on the host over the dense plan;
    NaNs are skipped with the reference's warning."""

    __slots__ = ["data", "alias", "_column"]

    def __init__(self, data, alias: str) -> None:
        self.data = data
        self.alias = alias
        self._column = None

    def __call__(self, partition, previous=None):
        ctx = partition._ctx
        if self._column is None or len(self._column) != ctx.n_nodes:
            if isinstance(self.data, str):
                nodes, attribute = partition.graph.nodes, self.data
                self.data = {node: nodes[node][attribute] for node in nodes}
            self._column = np.array([self.data[lab] for lab in ctx.labels], dtype=np.float64)
        col = self._column
        bad = np.isnan(col)
        for i in np.nonzero(bad)[0].tolist():
            warnings.warn("ignoring nan encountered at node '{}' for attribute '{}'".format(ctx.labels[i], self.alias))
        dl = partition._district_labels
        sums = np.bincount(partition._vec[~bad], weights=col[~bad], minlength=len(dl))
        integral = all(isinstance(self.data[lab], (int, np.integer)) for lab in ctx.labels)
        return {part: (int(round(v)) if integral else float(v)) for part, v in zip(dl, sums.tolist())}


def cut_edges(partition):
    """Set of edges crossing districts, each as ``tuple(sorted((u, v)))``
    (updaters/cut_edges.py:99-120)."""
    return partition._cut_edge_set()


def cut_edges_by_part(partition):
    """``{part: set of cut edges touching it}`` (updaters/cut_edges.py:76-96)."""
    out = {part: set() for part in partition.parts}
    mapping = partition.assignment.mapping
    for e in partition["cut_edges"]:
        out[mapping[e[0]]].add(e)
        out[mapping[e[1]]].add(e)
    return out


# ---- compactness (updaters/compactness.py) -------------------------------------------------
# Host-side statistics behind a state, computed with numpy from the dense district vector the
# engine works on (no flows to replay: a ReCom child rewrites two whole districts).  They are
# not on the accelerated path and touch neither the device nor the oracle.

def _per_part(partition, totals, integral):
    cast = (lambda x: int(round(x))) if integral else float
    return {part: cast(t) for part, t in zip(partition._district_labels, totals)}


def boundary_nodes(partition, alias: str = "boundary_nodes"):
    """Nodes on the boundary of the whole graph (``boundary_node`` attribute;
    updaters/compactness.py:8-25)."""
    on_boundary = partition._ctx.geometry()[0]
    labels = partition._ctx.labels
    return {labels[i] for i in np.nonzero(on_boundary)[0].tolist()}


def exterior_boundaries_as_a_set(partition):
    """``{part: boundary nodes in it}`` (updaters/compactness.py:28-69)."""
    on_boundary = partition._ctx.geometry()[0]
    labels, dl, vec = partition._ctx.labels, partition._district_labels, partition._vec
    out = {part: set() for part in dl}
    for i in np.nonzero(on_boundary)[0].tolist():
        out[dl[int(vec[i])]].add(labels[i])
    return out


def exterior_boundaries(partition):
    """``{part: total boundary_perim of its boundary nodes}`` (updaters/compactness.py:72-121).
    Every part has an entry (the reference's defaultdict answers 0 for the others)."""
    on_boundary, perim, _, integral = partition._ctx.geometry()
    k = len(partition._district_labels)
    totals = np.bincount(partition._vec[on_boundary], weights=perim[on_boundary], minlength=k)
    return _per_part(partition, totals, integral)


def interior_boundaries(partition):
    """``{part: total shared_perim of the cut edges touching it}``
    (updaters/compactness.py:124-175)."""
    _, _, shared, integral = partition._ctx.geometry()
    edges, vec = partition._ctx.edges, partition._vec
    k = len(partition._district_labels)
    if len(edges) == 0:
        return _per_part(partition, np.zeros(k), integral)
    a, b = vec[edges[:, 0]], vec[edges[:, 1]]
    cut = a != b
    totals = (np.bincount(a[cut], weights=shared[cut], minlength=k)
              + np.bincount(b[cut], weights=shared[cut], minlength=k))
    return _per_part(partition, totals, integral)


def perimeter_of_part(partition, part):
    """updaters/compactness.py:190-211."""
    return partition["exterior_boundaries"][part] + partition["interior_boundaries"][part]


def perimeter(partition):
    """``{part: exterior + interior boundary}`` (updaters/compactness.py:214-222)."""
    return {part: perimeter_of_part(partition, part) for part in partition.parts}


def flips(partition):
    """updaters/compactness.py:178-187."""
    return partition.flips


# ---- spanning-tree counts (updaters/spanning_trees.py) ---------------------------------------
def num_spanning_trees(partition):
    """``{part: number of spanning trees of the district's induced subgraph}`` by Kirchhoff's
    theorem: exp(log |det|) of the Laplacian with one row and one column struck (the
    reference strikes row 0 and column 1; every first minor has the same magnitude)."""
    import networkx

    out = {}
    for part, sub in partition.subgraphs.items():
        lap = networkx.laplacian_matrix(sub).toarray().astype(np.float64)
        minor = np.delete(np.delete(lap, 0, 0), 1, 1)
        out[part] = math.exp(np.linalg.slogdet(minor)[1])
    return out


# ---- locality split scores (updaters/locality_split_scores.py) -------------------------------
class LocalitySplits:
    """``LocalitySplits(name, col_id, pop_col, scores_to_compute, pent_alpha)``: scores of how
    a plan cuts localities (counties, municipalities).  All of them are functions of the
    locality x district contingency table of the dense plan (node counts; populations for the
    symmetric entropy), of the cut edges inside localities, or of the connected pieces left
    when the graph is cut along both locality and district lines.  Calling it returns - and
    stores in ``scores`` - ``{score name: value}`` for the names asked for, as in the
    reference."""

    KNOWN = ("num_parts", "num_pieces", "naked_boundary", "shannon_entropy", "power_entropy",
             "symmetric_entropy", "num_split_localities")

    def __init__(self, name: str, col_id: str, pop_col: str, scores_to_compute: List[str] = ["num_parts"],
                 pent_alpha: float = 0.05):
        self.name = name
        self.col_id = col_id
        self.pop_col = pop_col
        self.pent_alpha = pent_alpha
        self.localities = []
        self.allowed_pieces = {}
        self.scores = dict.fromkeys(scores_to_compute)

    # -- tables ------------------------------------------------------------------------
    def _table(self, partition, weights=None):
        codes, values, _ = partition._ctx.region_members(self.col_id)
        k = len(partition._district_labels)
        flat = np.bincount(codes.astype(np.int64) * k + partition._vec, weights=weights, minlength=len(values) * k)
        return flat.reshape(len(values), k)

    def __call__(self, partition):
        codes, values, _ = partition._ctx.region_members(self.col_id)
        if not self.localities:
            self.localities = set(values)
        if not self.allowed_pieces:  # districts a locality must touch at least, by population
            pop = partition._ctx.column(self.pop_col).astype(np.float64)
            per_locality = np.bincount(codes, weights=pop, minlength=len(values))
            ideal = pop.sum() / len(partition._district_labels)
            self.allowed_pieces = {v: math.ceil(t / ideal) for v, t in zip(values, per_locality.tolist())}
        for score in self.scores:
            if score in self.KNOWN:
                self.scores[score] = getattr(self, score)(partition)
        return self.scores

    # -- scores (locality_split_scores.py:200-431) ---------------------------------------
    def num_parts(self, partition) -> int:
        """Distinct (locality, district) pairs."""
        return int(np.count_nonzero(self._table(partition)))

    def num_split_localities(self, partition) -> int:
        """Localities touching two or more districts."""
        return int(np.count_nonzero(np.count_nonzero(self._table(partition), axis=1) > 1))

    def naked_boundary(self, partition) -> int:
        """Cut edges with both ends in one locality."""
        codes = partition._ctx.region_members(self.col_id)[0]
        edges, vec = partition._ctx.edges, partition._vec
        if len(edges) == 0:
            return 0
        u, v = edges[:, 0], edges[:, 1]
        return int(np.count_nonzero((vec[u] != vec[v]) & (codes[u] == codes[v])))

    def num_pieces(self, partition) -> int:
        """Connected pieces after cutting along locality AND district boundaries."""
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components

        codes = partition._ctx.region_members(self.col_id)[0]
        edges, vec = partition._ctx.edges, partition._vec
        n = len(vec)
        if len(edges):
            u, v = edges[:, 0], edges[:, 1]
            keep = (vec[u] == vec[v]) & (codes[u] == codes[v])
            u, v = u[keep], v[keep]
        else:
            u = v = np.zeros(0, dtype=np.int64)
        adj = coo_matrix((np.ones(len(u), dtype=np.int8), (u, v)), shape=(n, n))
        return int(connected_components(adj, directed=False)[0])

    def _entropy(self, partition, term):
        table = self._table(partition)
        total = table.sum()
        out = 0.0
        for row in table:
            size = row.sum()
            if size == 0:
                continue
            shares = row[row > 0] / size
            out += term(size / total, shares)
        return out

    def shannon_entropy(self, partition) -> float:
        """Sum over localities of (share of all nodes) x entropy of its split over districts."""
        return float(self._entropy(partition, lambda q, p: q * float(np.sum(p * np.log(1.0 / p)))))

    def power_entropy(self, partition) -> float:
        """Sum over localities of (1 / share of all nodes) x (sum of shares^(1 - alpha) - 1)."""
        a = self.pent_alpha
        return float(self._entropy(partition, lambda q, p: (float(np.sum(p ** (1 - a))) - 1.0) / q))

    def symmetric_entropy(self, partition) -> float:
        """Population-weighted: sum over districts of total x sum sqrt(share of each locality),
        plus the same with the roles of districts and localities exchanged."""
        pop = partition._ctx.column(self.pop_col).astype(np.float64)
        table = self._table(partition, weights=pop)
        score = 0.0
        for axis in (0, 1):  # columns = districts over localities, rows = localities over districts
            totals = table.sum(axis=axis)
            shares = table / (totals[None, :] if axis == 0 else totals[:, None])
            score += float(np.sum(totals * np.sqrt(shares).sum(axis=axis)))
        return score


# ---- county splits (updaters/county_splits.py:1-117) -----------------------------------------
CountyInfo = collections.namedtuple("CountyInfo", "split nodes contains")


class CountySplit(Enum):
    """Is a county in one district, split since the start, or split by this chain?"""

    NOT_SPLIT = 0
    NEW_SPLIT = 1
    OLD_SPLIT = 2


def county_splits(partition_name: str, county_field_name: str):
    """Updater ``{county: CountyInfo(split, nodes, contains)}``; ``partition_name`` is the
    key this updater is registered under (a child reads its parent's value through it)."""

    def _get_county_splits(partition):
        return compute_county_splits(partition, county_field_name, partition_name)

    return _get_county_splits


def compute_county_splits(partition, county_field: str, partition_field: str):
    """Districts present in every county, from the dense plan: the distinct (county,
    district) pairs.  Without a parent a county in several districts is an OLD_SPLIT; with one,
    it is OLD_SPLIT only if the parent said so and a NEW_SPLIT otherwise
    (updaters/county_splits.py:94-117)."""
    codes, values, members = partition._ctx.region_members(county_field)
    dl = partition._district_labels
    k = len(dl)
    pairs = np.unique(codes.astype(np.int64) * k + partition._vec)
    contains = [set() for _ in values]
    for pair in pairs.tolist():
        contains[pair // k].add(dl[pair % k])
    before = partition.parent[partition_field] if partition.parent else None
    out = {}
    for value, nodes, seen in zip(values, members, contains):
        if len(seen) <= 1:
            split = CountySplit.NOT_SPLIT
        elif before is None or before[value].split == CountySplit.OLD_SPLIT:
            split = CountySplit.OLD_SPLIT
        else:
            split = CountySplit.NEW_SPLIT
        out[value] = CountyInfo(split, nodes, seen)
    return out


def tally_region_splits(reg_attr_lst):
    """``{region attribute: number of its regions that are split}``
    (updaters/county_splits.py:120-161): a region counts as split when some cut edge has both
    ends in it.  Counted on the device over the plan's cut-edge bitmask
    (``rc_engine_region_splits``)."""

    def _get_splits(partition):
        if "cut_edges" not in partition.updaters:
            raise ValueError("The cut_edges updater must be attached to the partition")
        return {reg_attr: partition._region_stats(reg_attr)[0] for reg_attr in reg_attr_lst}

    return _get_splits


def regions_in_several_districts(reg_attr: str):
    """Number of regions of ``reg_attr`` whose nodes lie in more than one district - the
    counties ``county_splits`` (updaters/county_splits.py:58-117) marks other than NOT_SPLIT."""

    def _count(partition):
        return partition._region_stats(reg_attr)[1]

    return _count


class Election:
    """``Election(name, parties_to_columns, alias=None)`` (updaters/election.py:7-122): one
    vote column per party, tallied per district on the device (``rc_engine_tally``, or by the
    step kernel's incremental tally when the column is the balanced population).  Calling it
    on a partition returns :class:`ElectionResults`."""

    def __init__(self, name: str, parties_to_columns: Union[Dict, List], alias: Optional[str] = None):
        self.name = name
        self.alias = alias if alias is not None else name
        if isinstance(parties_to_columns, dict):
            pairs = list(parties_to_columns.items())
        elif isinstance(parties_to_columns, list):
            pairs = [(col, col) for col in parties_to_columns]
        else:
            raise TypeError("Election expects parties_to_columns to be a dict or list")
        for _, col in pairs:
            if not isinstance(col, str):
                raise TypeError("vote columns must be node-attribute names; the device tallies graph columns")
        self.parties = [party for party, _ in pairs]
        self.columns = [col for _, col in pairs]
        self.parties_to_columns = dict(pairs)

    def __str__(self):
        return "Election '{}' with vote totals for parties {} from columns {}.".format(
            self.name, str(self.parties), str(self.columns))

    def __repr__(self):
        return "Election(parties={}, columns={}, alias={})".format(str(self.parties), str(self.columns), str(self.alias))

    def __call__(self, partition) -> "ElectionResults":
        matrix = np.stack([partition._tally((col,)) for col in self.columns])
        return ElectionResults(self, matrix, list(partition._district_labels))


class ElectionResults:
    """Per-district vote totals with the reference's queries (updaters/election.py:188-367).
    Held as one ``[party, district]`` int64 matrix; the dict views the reference exposes
    (``totals_for_party``, ``totals``, ``percents_for_party``) are derived from it."""

    def __init__(self, election: Election, matrix: np.ndarray, regions: List):
        self.election = election
        self.regions = regions
        self._matrix = np.asarray(matrix, dtype=np.int64)
        self._row = {party: i for i, party in enumerate(election.parties)}
        self._col = {region: j for j, region in enumerate(regions)}
        district_totals = self._matrix.sum(axis=0)
        self.totals_for_party = {party: {r: int(self._matrix[i, j]) for r, j in self._col.items()}
                                 for party, i in self._row.items()}
        self.totals = {r: int(district_totals[j]) for r, j in self._col.items()}
        self.percents_for_party = {
            party: {r: (self._matrix[i, j] / district_totals[j] if district_totals[j] > 0 else math.nan)
                    for r, j in self._col.items()}
            for party, i in self._row.items()}

    def __str__(self):
        lines = ["Election Results for {}".format(self.election.name)]
        for region in self.regions:
            lines.append("{}:".format(region))
            lines.extend("  {}: {}".format(party, round(self.percents_for_party[party][region], 4))
                         for party in self.election.parties)
        return "\n".join(lines)

    def won(self, party: str, region) -> bool:
        column = self._matrix[:, self._col[region]]
        mine = column[self._row[party]]
        return bool(all(mine > other for i, other in enumerate(column) if i != self._row[party]))

    def seats(self, party: str) -> int:
        return sum(self.won(party, region) for region in self.regions)

    wins = seats

    def count(self, party: str, region=None) -> int:
        if region is not None:
            return int(self._matrix[self._row[party], self._col[region]])
        return int(self._matrix[self._row[party]].sum())

    def counts(self, party: str):
        return tuple(int(x) for x in self._matrix[self._row[party]])

    votes = counts

    def percent(self, party: str, region=None) -> float:
        if region is not None:
            return self.percents_for_party[party][region]
        return self.count(party) / self.total_votes()

    def percents(self, party: str):
        return tuple(self.percents_for_party[party][region] for region in self.regions)

    def total_votes(self) -> int:
        return int(self._matrix.sum())

    # partisan scores (updaters/election.py:369-422 -> metrics/partisan.py)
    def mean_median(self) -> float:
        from . import metrics
        return metrics.mean_median(self)

    def mean_thirdian(self) -> float:
        from . import metrics
        return metrics.mean_thirdian(self)

    def efficiency_gap(self) -> float:
        from . import metrics
        return metrics.efficiency_gap(self)

    def partisan_bias(self) -> float:
        from . import metrics
        return metrics.partisan_bias(self)

    def partisan_gini(self) -> float:
        from . import metrics
        return metrics.partisan_gini(self)
