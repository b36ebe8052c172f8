"""ctypes binding of the CPU oracle (oracle/librecom_oracle.so).

TEST INFRASTRUCTURE ONLY - imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs; never by gerrychain_b200.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from gerrychain_b200 import _abi
from gerrychain_b200.config import EngineConfig, strides

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}  # label width (8 / 16) -> CDLL


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def load(label_bits=8):
    """The oracle built with ``label_bits``-wide district labels (as the engine: 8 up to 255
    districts, 16 beyond)."""
    _lib = _libs.get(label_bits)
    if _lib is None:
        path = os.path.join(HERE, "librecom_oracle.so" if label_bits == 8 else "librecom_oracle_l16.so")
        if not os.path.exists(path):
            build()
        _lib = _libs[label_bits] = C.CDLL(path)
        _lib.rco_work_create.restype = C.c_void_p
        _lib.rco_work_create.argtypes = [C.POINTER(_abi.RcGraph)]
        _lib.rco_work_destroy.argtypes = [C.c_void_p]
        _lib.rco_replay.argtypes = [
            C.POINTER(_abi.RcGraph), C.POINTER(_abi.RcConfig), C.c_void_p,
            C.POINTER(_abi.RcReplay)] + [C.c_void_p] * 6
        _lib.rco_init_chains.argtypes = [
            C.POINTER(_abi.RcGraph), C.POINTER(_abi.RcConfig), C.POINTER(_abi.RcState)]
        _lib.rco_run_chains.argtypes = [
            C.POINTER(_abi.RcGraph), C.POINTER(_abi.RcConfig), C.c_int64,
            C.POINTER(_abi.RcState), C.c_int]
        _lib.rco_balanced_cuts_of_tree.argtypes = [
            C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int32, C.c_void_p]
        _lib.rco_mst_of_graph.argtypes = [C.POINTER(_abi.RcGraph), C.c_void_p, C.c_void_p]
        _lib.rco_seed_plan.argtypes = [
            C.POINTER(_abi.RcGraph), C.POINTER(_abi.RcConfig), C.c_void_p, C.c_uint32, C.c_int32,
            C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_abi.RcReplay)]
        _lib.rco_region_splits.argtypes = [C.POINTER(_abi.RcGraph), C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.c_void_p]
        _lib.rco_contiguity.argtypes = [C.POINTER(_abi.RcGraph), C.c_int32, C.c_void_p, C.c_void_p]
        _lib.rco_replay_reversible.argtypes = [
            C.POINTER(_abi.RcGraph), C.POINTER(_abi.RcConfig), C.c_void_p, C.c_int32, C.c_void_p,
            C.POINTER(_abi.RcReplay)] + [C.c_void_p] * 7
        _lib.rco_philox.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.rco_philox.restype = None
    return _lib


def philox(k0, k1, ctr):
    out = np.zeros(4, dtype=np.uint32)
    c = np.asarray(ctr, dtype=np.uint32)
    load().rco_philox(k0, k1, c.ctypes.data, out.ctypes.data)
    return out


def balanced_cuts_of_tree(n, tree_edges, pops, ideal, eps, root):
    uv = np.ascontiguousarray(tree_edges, dtype=np.int32)
    p = np.ascontiguousarray(pops, dtype=np.int32)
    out = np.zeros((n, 2), dtype=np.int32)
    nb = load().rco_balanced_cuts_of_tree(n, uv.ctypes.data, p.ctypes.data, ideal, eps, root,
                                          out.ctypes.data)
    return [tuple(map(int, r)) for r in out[:nb]]


def mst_of_graph(csr, weights):
    w = np.ascontiguousarray(weights, dtype=np.float64)
    out = np.zeros(csr.n_nodes, dtype=np.int32)
    g = _abi.graph_struct(csr)
    nt = load().rco_mst_of_graph(C.byref(g), w.ctypes.data, out.ctypes.data)
    if nt < 0:
        raise RuntimeError(_abi.STATUS_NAMES[-nt])
    return np.sort(out[:nt])


def region_splits(csr, k, assignment, region_ids, n_regions=None):
    """(regions with an inner cut edge, regions over > 1 district, pieces) of one plan."""
    bits = _abi.label_bits_for(k)
    a = np.ascontiguousarray(assignment, dtype=_abi.label_dtype(bits))
    r = np.ascontiguousarray(region_ids, dtype=np.int32)
    if n_regions is None:
        n_regions = int(r.max()) + 1
    out = np.zeros(3, dtype=np.int32)
    g = _abi.graph_struct(csr)
    load(bits).rco_region_splits(C.byref(g), k, a.ctypes.data, r.ctypes.data, n_regions, out.ctypes.data)
    return tuple(int(x) for x in out)


def contiguity(csr, k, assignment):
    """int32 [k]: connected pieces of every district of one plan."""
    bits = _abi.label_bits_for(k)
    a = np.ascontiguousarray(assignment, dtype=_abi.label_dtype(bits))
    out = np.zeros(k, dtype=np.int32)
    g = _abi.graph_struct(csr)
    load(bits).rco_contiguity(C.byref(g), k, a.ctypes.data, out.ctypes.data)
    return out


#: oracle/recom_oracle.h: RcoRevStep
REV_STEP_DTYPE = np.dtype([("d_out", "<i4"), ("d_in", "<i4"), ("pair_edge", "<i4"), ("root", "<i4"),
                           ("cut_edge", "<i4"), ("pad", "<i4"), ("tree_idx", "<i8"), ("accept_u", "<f8")])
REV_NO_PAIR, REV_NO_CUT, REV_REJECTED, REV_MOVED = range(4)


def replay_reversible(csr, cfg: EngineConfig, init_assignment, steps, wilson_root, stack_ptr, stack_val):
    """Apply recorded reference calls of reversible_recom (oracle/record_reversible.py) to one plan.
    steps: REV_STEP_DTYPE array; the Wilson arrays as in tests/_golden.ReplayArrays.
    Returns (status, assignment after every call [S][N], info [S][4] = trees, balanced cuts, seam
    length, outcome REV_*)."""
    bits = _abi.label_bits_for(cfg.n_districts)
    lib = load(bits)
    assert steps.dtype == REV_STEP_DTYPE and steps.flags.c_contiguous
    g = _abi.graph_struct(csr)
    c = cfg.to_struct()
    S, N, k = len(steps), csr.n_nodes, cfg.n_districts
    assign = np.ascontiguousarray(init_assignment, dtype=_abi.label_dtype(bits)).copy()
    tally = np.zeros(k, dtype=np.int64)
    mask = np.zeros(strides(csr.n_nodes, csr.n_edges)[1], dtype=np.uint32)
    cut_count = np.zeros(1, dtype=np.int32)
    counters = np.zeros(_abi.RC_N_COUNTERS, dtype=np.int64)
    trace = np.zeros((S, N), dtype=assign.dtype)
    info = np.zeros((S, 4), dtype=np.int32)
    rw = _abi.RcReplay()
    wr = np.ascontiguousarray(wilson_root, dtype=np.int32)
    sp = np.ascontiguousarray(stack_ptr, dtype=np.int64)
    sv = np.ascontiguousarray(stack_val, dtype=np.int32)
    rw.wilson_root, rw.stack_ptr, rw.stack_val = wr.ctypes.data, sp.ctypes.data, sv.ctypes.data
    lib.rco_init_state.argtypes = [C.POINTER(_abi.RcGraph), C.POINTER(_abi.RcConfig)] + [C.c_void_p] * 4
    rc = lib.rco_init_state(C.byref(g), C.byref(c), assign.ctypes.data, tally.ctypes.data, mask.ctypes.data,
                            cut_count.ctypes.data)
    assert rc == 0, rc
    w = lib.rco_work_create(C.byref(g))
    try:
        rc = lib.rco_replay_reversible(C.byref(g), C.byref(c), w, S, steps.ctypes.data, C.byref(rw),
                                       assign.ctypes.data, tally.ctypes.data, mask.ctypes.data,
                                       cut_count.ctypes.data, counters.ctypes.data, trace.ctypes.data,
                                       info.ctypes.data)
    finally:
        lib.rco_work_destroy(w)
    return rc, trace, info, tally, int(cut_count[0])


def seed_plans(csr, cfg: EngineConfig, n_plans=1, max_attempts=10000, rec=None):
    """recursive_tree_part on the CPU: (status [n], plans [n][N] uint8, tally [n][k]).
    rec: tests/_golden.ReplayArrays of recorded reference randomness (one plan)."""
    bits = _abi.label_bits_for(cfg.n_districts)
    lib = load(bits)
    g = _abi.graph_struct(csr)
    c = cfg.to_struct()
    plans = np.zeros((n_plans, csr.n_nodes), dtype=_abi.label_dtype(bits))
    tally = np.zeros((n_plans, cfg.n_districts), dtype=np.int64)
    status = np.zeros(n_plans, dtype=np.int32)
    counters = np.zeros((n_plans, _abi.RC_N_COUNTERS), dtype=np.int64)
    rp = None
    if rec is not None:
        rp = _abi.RcReplay()
        rp.n_steps = len(rec.steps)
        rp.steps = rec.steps.ctypes.data
        rp.weights = rec.weights.ctypes.data
    w = lib.rco_work_create(C.byref(g))
    try:
        for i in range(n_plans):
            status[i] = lib.rco_seed_plan(C.byref(g), C.byref(c), w, cfg.chain_offset + i, max_attempts,
                                          plans[i].ctypes.data, tally[i].ctypes.data,
                                          counters[i].ctypes.data, C.byref(rp) if rp is not None else None)
    finally:
        lib.rco_work_destroy(w)
    return status, plans, tally, counters


class OracleChains:
    """Host-resident twin of the device engine's state, advanced by the C oracle."""

    def __init__(self, csr, cfg: EngineConfig, assignment, label_bits=None):
        self.label_bits = label_bits or _abi.label_bits_for(cfg.n_districts)
        self.label_dtype = _abi.label_dtype(self.label_bits)
        self.lib = load(self.label_bits)
        self.csr = csr
        self.cfg = cfg
        self.cstruct = cfg.to_struct()
        self.gstruct = _abi.graph_struct(csr)
        n, k = cfg.n_chains, cfg.n_districts
        self.assign_stride, self.mask_stride = strides(csr.n_nodes, csr.n_edges)
        a = np.asarray(assignment, dtype=self.label_dtype)
        if a.ndim == 1:
            a = np.broadcast_to(a, (n, csr.n_nodes))
        self.assignment = np.zeros((n, self.assign_stride), dtype=self.label_dtype)
        self.assignment[:, : csr.n_nodes] = a
        self.tally = np.zeros((n, k), dtype=np.int64)
        self.cut_mask = np.zeros((n, self.mask_stride), dtype=np.uint32)
        self.cut_count = np.zeros(n, dtype=np.int32)
        self.step_no = np.zeros(n, dtype=np.int64)
        self.proposal_no = np.zeros(n, dtype=np.uint32)
        self.status = np.zeros(n, dtype=np.int32)
        self.flags = np.zeros(n, dtype=np.uint32)
        self.counters = np.zeros((n, _abi.RC_N_COUNTERS), dtype=np.int64)
        st = _abi.RcState()
        st.d_assignment = self.assignment.ctypes.data
        st.d_tally = self.tally.ctypes.data
        st.d_cut_mask = self.cut_mask.ctypes.data
        st.d_cut_count = self.cut_count.ctypes.data
        st.d_step_no = self.step_no.ctypes.data
        st.d_proposal_no = self.proposal_no.ctypes.data
        st.d_status = self.status.ctypes.data
        st.d_flags = self.flags.ctypes.data
        st.d_counters = self.counters.ctypes.data
        self.last_pair = np.full((n, 2), -1, dtype=np.int32)
        st.d_last_pair = self.last_pair.ctypes.data
        st.assign_stride = self.assign_stride
        st.mask_stride = self.mask_stride
        self.state = st
        rc = self.lib.rco_init_chains(C.byref(self.gstruct), C.byref(self.cstruct), C.byref(st))
        if rc:
            raise RuntimeError(_abi.STATUS_NAMES.get(rc, rc))

    def run(self, n_steps, n_threads=1):
        self.lib.rco_run_chains(C.byref(self.gstruct), C.byref(self.cstruct), int(n_steps),
                                C.byref(self.state), int(n_threads))
        return self

    def assignments(self):
        return self.assignment[:, : self.csr.n_nodes]

    def replay(self, rec, chain=0):
        """rec: tests/_golden.ReplayArrays (host numpy).  Returns (assign_trace, info)."""
        n_steps = len(rec.steps)
        trace = np.zeros((n_steps, self.csr.n_nodes), dtype=self.label_dtype)
        info = np.zeros((n_steps, 4), dtype=np.int32)
        rp = _abi.RcReplay()
        rp.n_steps = n_steps
        rp.chain = chain
        rp.steps = rec.steps.ctypes.data
        rp.weights = rec.weights.ctypes.data if rec.weights is not None else None
        rp.weight_rank = None
        rp.wilson_root = rec.wilson_root.ctypes.data if rec.wilson_root is not None else None
        rp.stack_ptr = rec.stack_ptr.ctypes.data if rec.stack_ptr is not None else None
        rp.stack_val = rec.stack_val.ctypes.data if rec.stack_val is not None else None
        rp.assign_trace = trace.ctypes.data
        rp.info_trace = info.ctypes.data
        w = self.lib.rco_work_create(C.byref(self.gstruct))
        try:
            rc = self.lib.rco_replay(
                C.byref(self.gstruct), C.byref(self.cstruct), w, C.byref(rp),
                self.assignment[chain].ctypes.data, self.tally[chain].ctypes.data,
                self.cut_mask[chain].ctypes.data,
                self.cut_count[chain:].ctypes.data, self.counters[chain].ctypes.data,
                self.flags[chain:].ctypes.data)
        finally:
            self.lib.rco_work_destroy(w)
        return rc, trace, info
