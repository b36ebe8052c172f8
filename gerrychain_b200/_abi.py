"""ctypes mirror of include/recom_b200.h (struct layouts, enums, loader).

The CUDA library is the ONLY compute path of this package: if it cannot be loaded the
import of anything that needs it raises (no CPU fallback, no oracle import).
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

RC_OK = 0
RC_ERR_INVALID_ARG = 1
RC_ERR_CUDA = 2
RC_ERR_CAPACITY = 3
RC_ERR_MAX_ATTEMPTS = 4
RC_ERR_ALL_PAIRS_BAD = 5
RC_ERR_DISCONNECTED = 6
RC_ERR_NO_CUT_EDGES = 7
RC_ERR_REPLAY = 8
RC_ERR_NO_ROOT = 9
RC_ERR_INVALID_STREAK = 10
RC_ERR_BALANCE = 11
RC_ERR_REVERSIBILITY = 12
RC_PROPOSAL_RECOM = 0
RC_PROPOSAL_REVERSIBLE = 1
RC_SEED_REMAINING = 255  # in the 8-bit-label build; 65535 in librecom_b200_l16.so
RC_MAX_DISTRICTS = {8: 255, 16: 32767}  # include/recom_b200.h: RC_MAX_DISTRICTS per label width

STATUS_NAMES = {
    0: "RC_OK",
    1: "RC_ERR_INVALID_ARG",
    2: "RC_ERR_CUDA",
    3: "RC_ERR_CAPACITY",
    4: "RC_ERR_MAX_ATTEMPTS",
    5: "RC_ERR_ALL_PAIRS_BAD",
    6: "RC_ERR_DISCONNECTED",
    7: "RC_ERR_NO_CUT_EDGES",
    8: "RC_ERR_REPLAY",
    9: "RC_ERR_NO_ROOT",
    10: "RC_ERR_INVALID_STREAK",
    11: "RC_ERR_BALANCE",
    12: "RC_ERR_REVERSIBILITY",
}

RC_FLAG_WARNED = 1

RC_TREE_MST = 0
RC_TREE_WILSON = 1

RC_ACCEPT_ALWAYS = 0
RC_ACCEPT_CUT_EDGE = 1

RC_CUT_UNIFORM = 0
RC_CUT_REGION_PREF = 1
RC_CUT_MAX_WEIGHT = 2

(RC_CTR_TREES, RC_CTR_ATTEMPTS, RC_CTR_NODES, RC_CTR_SLOTS, RC_CTR_INVALID,
 RC_CTR_RESELECT, RC_CTR_WALK, RC_CTR_ROUNDS) = range(8)
RC_N_COUNTERS = 8

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)


class RcGraph(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32),
        ("n_edges", C.c_int32),
        ("row_ptr", C.c_void_p),
        ("col_idx", C.c_void_p),
        ("slot_edge", C.c_void_p),
        ("edge_uv", C.c_void_p),
        ("pop", C.c_void_p),
        ("n_region_cols", C.c_int32),
        ("region_ids", C.c_void_p),
        ("surcharges", C.c_void_p),
    ]


class RcConfig(C.Structure):
    _fields_ = [
        ("n_districts", C.c_int32),
        ("n_chains", C.c_int32),
        ("pop_target", C.c_double),
        ("epsilon", C.c_double),
        ("node_repeats", C.c_int32),
        ("max_attempts", C.c_int32),
        ("warn_attempts", C.c_int32),
        ("allow_pair_reselection", C.c_int32),
        ("tree_kind", C.c_int32),
        ("cut_policy", C.c_int32),
        ("pop_lower", C.c_double),
        ("pop_upper", C.c_double),
        ("seed", C.c_uint64),
        ("chain_offset", C.c_int32),
        ("cap_nodes", C.c_int32),
        ("cap_edges", C.c_int32),
        ("threads_per_cta", C.c_int32),
        ("max_invalid_streak", C.c_int32),
        ("accept_rule", C.c_int32),
        ("proposal_kind", C.c_int32),
        ("rev_M", C.c_int32),
        ("wilson_walkers", C.c_int32),
    ]


class RcState(C.Structure):
    _fields_ = [
        ("d_assignment", C.c_void_p),
        ("d_tally", C.c_void_p),
        ("d_cut_mask", C.c_void_p),
        ("d_cut_count", C.c_void_p),
        ("d_step_no", C.c_void_p),
        ("d_proposal_no", C.c_void_p),
        ("d_status", C.c_void_p),
        ("d_flags", C.c_void_p),
        ("d_counters", C.c_void_p),
        ("assign_stride", C.c_int64),
        ("mask_stride", C.c_int64),
        ("d_last_pair", C.c_void_p),
    ]


class RcReplayStep(C.Structure):
    _fields_ = [
        ("pair_edge", C.c_int32),
        ("n_trees", C.c_int32),
        ("first_tree", C.c_int64),
        ("root", C.c_int32),
        ("cut_edge", C.c_int32),
    ]


class RcReplay(C.Structure):
    _fields_ = [
        ("n_steps", C.c_int32),
        ("chain", C.c_int32),
        ("steps", C.c_void_p),
        ("weights", C.c_void_p),
        ("weight_rank", C.c_void_p),
        ("wilson_root", C.c_void_p),
        ("stack_ptr", C.c_void_p),
        ("stack_val", C.c_void_p),
        ("assign_trace", C.c_void_p),
        ("info_trace", C.c_void_p),
    ]


class RcEngineInfo(C.Structure):
    _fields_ = [
        ("cap_nodes", C.c_int32),
        ("cap_edges", C.c_int32),
        ("threads_per_cta", C.c_int32),
        ("ctas_per_sm", C.c_int32),
        ("n_sms", C.c_int32),
        ("grid", C.c_int32),
        ("smem_bytes", C.c_int64),
        ("scratch_bytes", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("wilson_walkers", C.c_int32),
        ("spill", C.c_int32),
    ]


EXPORTS = (
    "rc_last_error",
    "rc_abi_version",
    "rc_label_bits",
    "rc_state_strides",
    "rc_engine_create",
    "rc_engine_destroy",
    "rc_engine_bind_state",
    "rc_engine_init_state",
    "rc_engine_run",
    "rc_engine_step_host",
    "rc_engine_seed",
    "rc_engine_seed_replay",
    "rc_engine_replay",
    "rc_engine_histograms",
    "rc_engine_tally",
    "rc_engine_region_splits",
    "rc_engine_contiguity",
    "rc_engine_info",
    "rc_engine_last_run_ms",
)

LIB_NAME = "librecom_b200.so"          # district labels in one byte (k <= 255)
LIB_NAME_L16 = "librecom_b200_l16.so"  # the same sources built with 16-bit labels (k <= 32767)
_libs: dict = {}


def label_bits_for(n_districts: int) -> int:
    """Width of the district labels a plan with ``n_districts`` districts needs."""
    if n_districts <= RC_MAX_DISTRICTS[8]:
        return 8
    if n_districts <= RC_MAX_DISTRICTS[16]:
        return 16
    raise ValueError(f"at most {RC_MAX_DISTRICTS[16]} districts (16-bit labels)")


def label_dtype(label_bits: int):
    import numpy as np

    return np.uint8 if label_bits == 8 else np.uint16


def lib_path(label_bits: int = 8) -> str:
    """The in-tree library; RECOM_B200_LIB (RECOM_B200_LIB_L16 for the 16-bit-label build)
    selects another build of the same ABI (A/B runs)."""
    override = os.environ.get("RECOM_B200_LIB" if label_bits == 8 else "RECOM_B200_LIB_L16")
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME if label_bits == 8 else LIB_NAME_L16)


def load(label_bits: int = 8) -> C.CDLL:
    """Load the in-tree CUDA library (the build with ``label_bits``-wide district labels);
    raise loudly if it is missing."""
    if label_bits not in (8, 16):
        raise ValueError("label_bits must be 8 or 16")
    if label_bits in _libs:
        return _libs[label_bits]
    path = lib_path(label_bits)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  gerrychain_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    lib.rc_last_error.restype = C.c_char_p
    lib.rc_abi_version.restype = C.c_int
    lib.rc_label_bits.restype = C.c_int
    for name in EXPORTS[3:]:
        getattr(lib, name).restype = C.c_int
    lib.rc_state_strides.argtypes = [C.POINTER(RcGraph), C.POINTER(RcConfig), _i64p, _i64p]
    lib.rc_engine_create.argtypes = [
        C.POINTER(RcGraph), C.POINTER(RcConfig), C.c_int, C.POINTER(C.c_void_p)]
    lib.rc_engine_destroy.argtypes = [C.c_void_p]
    lib.rc_engine_bind_state.argtypes = [C.c_void_p, C.POINTER(RcState)]
    lib.rc_engine_init_state.argtypes = [C.c_void_p, C.c_void_p]
    lib.rc_engine_run.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.rc_engine_step_host.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p]
    lib.rc_engine_seed.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.rc_engine_seed_replay.argtypes = [C.c_void_p, C.c_int32, C.POINTER(RcReplay), C.c_void_p]
    lib.rc_engine_replay.argtypes = [C.c_void_p, C.POINTER(RcReplay), C.c_void_p]
    lib.rc_engine_histograms.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
    lib.rc_engine_tally.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rc_engine_region_splits.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.rc_engine_contiguity.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rc_engine_info.argtypes = [C.c_void_p, C.POINTER(RcEngineInfo)]
    lib.rc_engine_last_run_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    if lib.rc_abi_version() != 1:
        raise ImportError(f"{path}: ABI version mismatch")
    if lib.rc_label_bits() != label_bits:
        raise ImportError(f"{path} was built with {lib.rc_label_bits()}-bit labels, not {label_bits}")
    _libs[label_bits] = lib
    return lib


def graph_struct(csr) -> RcGraph:
    """RcGraph over the numpy arrays of a GraphCSR (keeps no references: the caller must
    keep ``csr`` alive for the duration of the call it passes the struct to)."""
    g = RcGraph()
    g.n_nodes = csr.n_nodes
    g.n_edges = csr.n_edges
    g.row_ptr = csr.row_ptr.ctypes.data
    g.col_idx = csr.col_idx.ctypes.data
    g.slot_edge = csr.slot_edge.ctypes.data
    g.edge_uv = csr.edge_uv.ctypes.data
    g.pop = csr.pop.ctypes.data
    g.n_region_cols = csr.n_region_cols
    g.region_ids = csr.region_ids.ctypes.data if csr.n_region_cols else None
    g.surcharges = csr.surcharges.ctypes.data if csr.n_region_cols else None
    return g
