"""Device engine: torch-owned chain state + the C-ABI calls of librecom_b200.so.

PyTorch is plumbing here (device buffers, streams, torch.distributed); every compute
step is a CUDA kernel of the in-tree library reached through ctypes.  There is no CPU
fallback: constructing an ``Engine`` without the library or without a CUDA device raises.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _abi
from .config import EngineConfig, strides
from .csr import GraphCSR


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str = ""):
        self.code = code
        name = _abi.STATUS_NAMES.get(code, str(code))
        super().__init__(f"{name}: {message}" if message else name)


def _check(lib, rc: int):
    if rc != _abi.RC_OK:
        raise EngineError(rc, lib.rc_last_error().decode())


class Engine:
    """n_chains independent ReCom chains resident on one GPU."""

    def __init__(self, csr: GraphCSR, cfg: EngineConfig, assignment, device=None, label_bits: Optional[int] = None):
        """``label_bits``: 8 or 16, which build of the library holds the plans (default: 8 up
        to 255 districts, 16 beyond; 16 may be forced for a small plan - the results are the same)."""
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("gerrychain_b200.Engine needs a CUDA device (no CPU fallback)")
        # plans with more than 255 districts run on the 16-bit-label build of the library
        self.label_bits = max(_abi.label_bits_for(cfg.n_districts), int(label_bits or 8))
        self.label_dtype = _abi.label_dtype(self.label_bits)
        self.lib = _abi.load(self.label_bits)
        self.torch = torch
        # (labels <= 32767, so the signed torch type holds them; numpy views are unsigned)
        self.torch_label = torch.uint8 if self.label_bits == 8 else torch.int16
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.csr = csr
        self.cfg = cfg
        n, k = cfg.n_chains, cfg.n_districts
        self.assign_stride, self.mask_stride = strides(csr.n_nodes, csr.n_edges)
        dev = self.device
        a = np.asarray(assignment)
        if a.ndim == 1:
            a = np.broadcast_to(a, (n, csr.n_nodes))
        if a.shape != (n, csr.n_nodes):
            raise ValueError("assignment must be [N] or [n_chains, N]")
        if a.min() < 0 or a.max() >= k:
            raise ValueError("assignment entries must be district indices in [0, k)")
        self.assignment = self._labels_to_torch(a).to(dev)
        self.tally = torch.zeros((n, k), dtype=torch.int64, device=dev)
        self.cut_mask = torch.zeros((n, self.mask_stride), dtype=torch.int32, device=dev)
        self.cut_count = torch.zeros(n, dtype=torch.int32, device=dev)
        self.step_no = torch.zeros(n, dtype=torch.int64, device=dev)
        self.proposal_no = torch.zeros(n, dtype=torch.int32, device=dev)
        self.status = torch.zeros(n, dtype=torch.int32, device=dev)
        self.flags = torch.zeros(n, dtype=torch.int32, device=dev)
        self.counters = torch.zeros((n, _abi.RC_N_COUNTERS), dtype=torch.int64, device=dev)
        self._handle = C.c_void_p()
        g = _abi.graph_struct(csr)
        c = cfg.to_struct()
        _check(self.lib, self.lib.rc_engine_create(C.byref(g), C.byref(c), self.device.index, C.byref(self._handle)))
        st = _abi.RcState()
        st.d_assignment = self.assignment.data_ptr()
        st.d_tally = self.tally.data_ptr()
        st.d_cut_mask = self.cut_mask.data_ptr()
        st.d_cut_count = self.cut_count.data_ptr()
        st.d_step_no = self.step_no.data_ptr()
        st.d_proposal_no = self.proposal_no.data_ptr()
        st.d_status = self.status.data_ptr()
        st.d_flags = self.flags.data_ptr()
        st.d_counters = self.counters.data_ptr()
        self.last_pair = torch.full((n, 2), -1, dtype=torch.int32, device=dev)
        st.d_last_pair = self.last_pair.data_ptr()
        st.assign_stride = self.assign_stride
        st.mask_stride = self.mask_stride
        self._state = st
        _check(self.lib, self.lib.rc_engine_bind_state(self._handle, C.byref(st)))
        self.init_state()

    def _labels_to_torch(self, a):
        """[n_chains, N] district indices -> CPU tensor [n_chains, assign_stride] of the label type."""
        host = np.zeros((self.cfg.n_chains, self.assign_stride), dtype=self.label_dtype)
        host[:, : self.csr.n_nodes] = a
        return self.torch.from_numpy(host if self.label_bits == 8 else host.view(np.int16))

    def _labels_to_numpy(self, t):
        a = t.cpu().numpy()
        return a if self.label_bits == 8 else a.view(np.uint16)

    # -- lifecycle ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) and self._handle.value:
            self.torch.cuda.synchronize(self.device)
            self.lib.rc_engine_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    # -- ABI calls ---------------------------------------------------------------------
    def init_state(self):
        _check(self.lib, self.lib.rc_engine_init_state(self._handle, self._stream()))

    def set_assignment(self, assignment):
        """Host -> device copy of new plans (one per chain or one for all) + re-init."""
        a = np.asarray(assignment, dtype=self.label_dtype)
        if a.ndim == 1:
            a = np.broadcast_to(a, (self.cfg.n_chains, self.csr.n_nodes))
        self.assignment.copy_(self._labels_to_torch(a), non_blocking=False)
        self.init_state()

    def run(self, n_steps: int):
        """Advance every chain by ``n_steps`` accepted steps (asynchronous)."""
        _check(self.lib, self.lib.rc_engine_run(self._handle, int(n_steps), self._stream()))
        return self

    def host_buffers(self):
        """Pinned host buffers shaped for :meth:`step_host` (assignment, tally, cut_count, status)."""
        t = self.torch
        n, k = self.cfg.n_chains, self.cfg.n_districts
        # rows padded to the device pitch: each direction is then one contiguous copy
        return (t.zeros((n, self.assign_stride), dtype=self.torch_label).pin_memory()[:, : self.csr.n_nodes],
                t.empty((n, k), dtype=t.int64).pin_memory(),
                t.empty(n, dtype=t.int32).pin_memory(),
                t.empty(n, dtype=t.int32).pin_memory())

    def step_host(self, h_assignment, n_steps: int, h_out=None, h_tally=None, h_cut_count=None,
                  h_status=None):
        """Host plans in, ``n_steps`` ReCom steps on the device, host plans / Tally /
        len(cut_edges) / status out (``rc_engine_step_host``); synchronous.  Arguments are
        CPU torch tensors (pinned for asynchronous copies) or C-contiguous numpy arrays."""
        def ptr(x):
            if x is None:
                return None
            return x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data

        def pitch(x):
            # row pitch in labels (include/recom_b200.h: h_stride)
            return int(x.stride(0)) if hasattr(x, "data_ptr") else int(x.strides[0]) // x.itemsize

        n, N = self.cfg.n_chains, self.csr.n_nodes
        if tuple(h_assignment.shape) != (n, N):
            raise ValueError("h_assignment must be [n_chains, N] district labels")
        isz = h_assignment.element_size() if hasattr(h_assignment, "data_ptr") else h_assignment.itemsize
        if isz * 8 != self.label_bits:
            raise ValueError(f"h_assignment must hold {self.label_bits}-bit labels")
        if h_out is not None and (tuple(h_out.shape) != (n, N) or pitch(h_out) != pitch(h_assignment)):
            raise ValueError("h_out must have the shape and row pitch of h_assignment")
        _check(self.lib, self.lib.rc_engine_step_host(
            self._handle, ptr(h_assignment), pitch(h_assignment), int(n_steps), ptr(h_out), ptr(h_tally),
            ptr(h_cut_count), ptr(h_status), self._stream()))
        return self

    def seed(self, max_attempts: int = 10000):
        """Random seed plans for every chain (``rc_engine_seed``: the reference's
        ``recursive_tree_part``); the engine must cover the whole graph
        (``cap_nodes = N, cap_edges = E``).  Asynchronous."""
        _check(self.lib, self.lib.rc_engine_seed(self._handle, int(max_attempts or 0), self._stream()))
        return self

    def seed_replay(self, rec, chain: int = 0, max_attempts: int = 10000):
        """TEST HOOK: recursive_tree_part with recorded reference randomness."""
        t = self.torch
        keep = []

        def up(arr):
            x = t.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(self.device)
            keep.append(x)
            return x.data_ptr()

        rp = _abi.RcReplay()
        rp.n_steps = len(rec.steps)
        rp.chain = chain
        rp.steps = up(rec.steps)
        rp.weight_rank = up(rec.weight_rank)
        _check(self.lib, self.lib.rc_engine_seed_replay(self._handle, int(max_attempts or 0), C.byref(rp), self._stream()))
        t.cuda.synchronize(self.device)
        return self

    def last_run_ms(self) -> float:
        ms = C.c_float()
        _check(self.lib, self.lib.rc_engine_last_run_ms(self._handle, C.byref(ms)))
        return float(ms.value)

    def info(self) -> _abi.RcEngineInfo:
        out = _abi.RcEngineInfo()
        _check(self.lib, self.lib.rc_engine_info(self._handle, C.byref(out)))
        return out

    def histograms(self, n_cut_bins: int, n_dev_bins: int, dev_bin_width: float):
        t = self.torch
        cut = t.zeros(n_cut_bins, dtype=t.int64, device=self.device)
        dev = t.zeros(n_dev_bins, dtype=t.int64, device=self.device)
        _check(self.lib, self.lib.rc_engine_histograms(
            self._handle, cut.data_ptr(), n_cut_bins, dev.data_ptr(), n_dev_bins,
            float(dev_bin_width), self._stream()))
        return cut, dev

    def tally_column(self, column):
        """int64 [n_chains, k]: per-district sums of an integer node column (``rc_engine_tally``)."""
        t = self.torch
        col = t.as_tensor(np.ascontiguousarray(column, dtype=np.int32)).to(self.device)
        if col.numel() != self.csr.n_nodes:
            raise ValueError("column must have one entry per node")
        out = t.empty((self.cfg.n_chains, self.cfg.n_districts), dtype=t.int64, device=self.device)
        _check(self.lib, self.lib.rc_engine_tally(self._handle, col.data_ptr(), out.data_ptr(), self._stream()))
        self.synchronize()  # `col` must outlive the kernel
        return out

    def region_splits(self, region_ids, n_regions: Optional[int] = None):
        """int32 [n_chains, 3] (``rc_engine_region_splits``): per chain the regions with a cut
        edge inside them (the reference's ``total_reg_splits``), the regions spread over more
        than one district, and the total number of (region, district) pieces."""
        t = self.torch
        reg = np.ascontiguousarray(region_ids, dtype=np.int32)
        if reg.size != self.csr.n_nodes:
            raise ValueError("region_ids must have one entry per node")
        if n_regions is None:
            n_regions = int(reg.max()) + 1 if reg.size else 0
        col = t.as_tensor(reg).to(self.device)
        out = t.empty((self.cfg.n_chains, 3), dtype=t.int32, device=self.device)
        _check(self.lib, self.lib.rc_engine_region_splits(self._handle, col.data_ptr(), int(n_regions),
                                                          out.data_ptr(), self._stream()))
        self.synchronize()  # `col` must outlive the kernel
        return out

    def contiguity(self):
        """int32 [n_chains, k] (``rc_engine_contiguity``): connected pieces of every district
        (0 = empty district); a plan is contiguous iff no entry exceeds 1."""
        t = self.torch
        out = t.empty((self.cfg.n_chains, self.cfg.n_districts), dtype=t.int32, device=self.device)
        _check(self.lib, self.lib.rc_engine_contiguity(self._handle, out.data_ptr(), self._stream()))
        return out

    def replay(self, rec, chain: int = 0):
        """TEST HOOK: apply recorded reference steps (tests/_golden.ReplayArrays)."""
        t = self.torch
        dev = self.device
        n_steps = len(rec.steps)
        keep = []

        def up(arr):
            if arr is None:
                return None
            x = t.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(dev)
            keep.append(x)
            return x.data_ptr()

        trace = t.zeros((n_steps, self.csr.n_nodes), dtype=self.torch_label, device=dev)
        info = t.zeros((n_steps, 4), dtype=t.int32, device=dev)
        rp = _abi.RcReplay()
        rp.n_steps = n_steps
        rp.chain = chain
        rp.steps = up(rec.steps)
        rp.weights = None  # float64 weights are the oracle's input; the device takes ranks
        rp.weight_rank = up(rec.weight_rank)
        rp.wilson_root = up(rec.wilson_root)
        rp.stack_ptr = up(rec.stack_ptr)
        rp.stack_val = up(rec.stack_val)
        rp.assign_trace = trace.data_ptr()
        rp.info_trace = info.data_ptr()
        _check(self.lib, self.lib.rc_engine_replay(self._handle, C.byref(rp), self._stream()))
        t.cuda.synchronize(dev)
        return self._labels_to_numpy(trace), info.cpu().numpy()

    # -- read-backs --------------------------------------------------------------------
    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

    def assignments(self) -> np.ndarray:
        return self._labels_to_numpy(self.assignment[:, : self.csr.n_nodes])

    def check_status(self):
        """Raise the reference's exception for the first chain that stopped."""
        st = self.status.cpu().numpy()
        bad = np.nonzero(st)[0]
        if len(bad):
            raise EngineError(int(st[bad[0]]), f"chain {int(bad[0])} stopped ({len(bad)} chains in error)")
