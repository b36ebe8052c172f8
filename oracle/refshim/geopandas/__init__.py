"""Three-line stand-in for geopandas so the pure-Python reference imports here.

TEST INFRASTRUCTURE ONLY (see oracle/README.md). The reference only uses
`GeoDataFrame` as a type hint on the path we exercise
(/root/reference/gerrychain/graph/adjacency.py:13, graph/geo.py:12).
"""


class GeoDataFrame:  # pragma: no cover - type-hint stub
    pass


class _Options:
    use_pygeos = False


options = _Options()
