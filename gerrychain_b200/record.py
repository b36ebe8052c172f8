"""``Record`` / ``Replay``: store a chain run as assignment deltas and play it back.

The reference recommends wrapping a chain in ``pcompress.Record(chain, "run.chain")`` and
reading it back with ``pcompress.Replay(graph, "run.chain", updaters=...)``
(/root/reference/docs/topics/reproducibility.rst:21-60); pcompress itself is a separate Rust
tool that is not part of the reference tree, so this module offers the same two wrappers
over its own small container:

    for partition in Record(chain, "run.rcd"):       # unchanged driver loop
        ...
    for partition in Replay(graph, "run.rcd", updaters=my_updaters):
        ...

A ReCom step relabels the nodes of two districts only, so a step is stored as the list of
nodes whose district changed (index gaps and new district, LEB128 varints), the whole stream
deflated.  File layout::

    b"RCD1" | u32 header length | header JSON | zlib stream of
        initial plan: N bytes (district rank per node, node order of the graph)
        per step:     varint n_changed, then n_changed x (varint gap to the previous changed
                      node index + 1 ... first gap is index + 1, varint new district rank)

Replayed partitions are built on the dense vector (no dict copy); their ``Tally`` /
``cut_edges`` values come from the device on first use, like any other partition here.
"""

from __future__ import annotations

import json
import struct
import zlib
from typing import Dict, Iterable, Iterator, Optional

import numpy as np

from ._context import context_of
from .partition import Partition

MAGIC = b"RCD1"


def _put_varint(out: bytearray, x: int) -> None:
    while x >= 0x80:
        out.append((x & 0x7F) | 0x80)
        x >>= 7
    out.append(x)


class _Reader:
    """Incremental reader over a deflated byte stream."""

    def __init__(self, fh):
        self._fh = fh
        self._z = zlib.decompressobj()
        self._buf = bytearray()
        self._pos = 0
        self._eof = False

    def _fill(self, need: int) -> bool:
        while len(self._buf) - self._pos < need and not self._eof:
            chunk = self._fh.read(1 << 16)
            if not chunk:
                self._buf += self._z.flush()
                self._eof = True
                break
            self._buf += self._z.decompress(chunk)
        if self._pos > (1 << 20):
            del self._buf[: self._pos]
            self._pos = 0
        return len(self._buf) - self._pos >= need

    def take(self, n: int) -> bytes:
        if not self._fill(n):
            raise EOFError("truncated recording")
        b = bytes(self._buf[self._pos: self._pos + n])
        self._pos += n
        return b

    def varint(self) -> Optional[int]:
        """Next varint, or None at a clean end of the stream."""
        x = shift = 0
        first = True
        while True:
            if not self._fill(1):
                if first:
                    return None
                raise EOFError("truncated recording")
            byte = self._buf[self._pos]
            self._pos += 1
            first = False
            x |= (byte & 0x7F) << shift
            if byte < 0x80:
                return x
            shift += 7


class Record:
    """Iterate ``chain`` and write every state it yields to ``filename``.

    ``chain`` is anything that yields partitions of one graph (``MarkovChain``, or an
    iterable of ``Partition``).  States are passed through untouched."""

    def __init__(self, chain: Iterable[Partition], filename: str, level: int = 6):
        self.chain = chain
        self.filename = filename
        self.level = level

    def __len__(self):
        return len(self.chain)

    def __iter__(self) -> Iterator[Partition]:
        prev = None
        z = zlib.compressobj(self.level)
        with open(self.filename, "wb") as fh:
            for state in self.chain:
                vec = state.assignment_vector()
                if prev is None:
                    header = json.dumps({
                        "n_nodes": int(vec.shape[0]),
                        "district_labels": [_jsonable(x) for x in state._district_labels],
                        "label_bytes": int(vec.dtype.itemsize),
                        "format": "initial plan + per-step (gap, district) varints, zlib",
                    }).encode()
                    fh.write(MAGIC + struct.pack("<I", len(header)) + header)
                    fh.write(z.compress(vec.tobytes()))
                else:
                    changed = np.nonzero(vec != prev)[0]
                    out = bytearray()
                    _put_varint(out, int(changed.size))
                    last = -1
                    for i, d in zip(changed.tolist(), vec[changed].tolist()):
                        _put_varint(out, i - last)
                        _put_varint(out, d)
                        last = i
                    fh.write(z.compress(bytes(out)))
                prev = vec.copy()
                yield state
            fh.write(z.flush())


def _jsonable(x):
    return x.item() if isinstance(x, np.generic) else x


class Replay:
    """Yield the recorded states as partitions of ``graph`` with the given ``updaters``
    (``cut_edges`` is attached by default, as for ``Partition``)."""

    def __init__(self, graph, filename: str, updaters: Optional[Dict] = None, use_default_updaters: bool = True,
                 partition_class=Partition):
        self.graph = graph
        self.filename = filename
        self.updaters = updaters
        self.use_default_updaters = use_default_updaters
        self.partition_class = partition_class

    def __iter__(self) -> Iterator[Partition]:
        ctx = context_of(self.graph)
        with open(self.filename, "rb") as fh:
            if fh.read(4) != MAGIC:
                raise ValueError(f"{self.filename} is not a chain recording")
            (hlen,) = struct.unpack("<I", fh.read(4))
            header = json.loads(fh.read(hlen).decode())
            n = header["n_nodes"]
            if n != ctx.n_nodes:
                raise ValueError(f"recording has {n} nodes, the graph has {ctx.n_nodes}")
            labels = [tuple(x) if isinstance(x, list) else x for x in header["district_labels"]]
            rd = _Reader(fh)
            lb = int(header.get("label_bytes", 1))  # 2 for plans with more than 255 districts
            vec = np.frombuffer(rd.take(n * lb), dtype=np.uint8 if lb == 1 else np.dtype("<u2")).copy()
            first = self.partition_class(self.graph, {lab: labels[d] for lab, d in zip(ctx.labels, vec.tolist())},
                                         self.updaters, use_default_updaters=self.use_default_updaters)
            # the recorded ranks index the recorded label list; keep exactly that order
            first._district_labels = labels
            first._rank = {p: i for i, p in enumerate(labels)}
            first._vec = vec.copy()
            yield first
            parent = first
            while True:
                n_changed = rd.varint()
                if n_changed is None:
                    return
                vec = vec.copy()
                i = -1
                flips = {}
                for _ in range(n_changed):
                    i += rd.varint()
                    d = rd.varint()
                    vec[i] = d
                    flips[ctx.labels[i]] = labels[d]
                child = self.partition_class.__new__(self.partition_class)
                child._inherit(parent)
                child._flips = flips  # the nodes whose district changed
                child._vec = vec
                child.parent = parent
                parent.parent = None  # as MarkovChain does (chain.py:188-189): two generations alive
                yield child
                parent = child
