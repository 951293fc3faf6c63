"""ctypes binding of the C ABI in include/adsel_b200.h (``csrc/libadsel_b200.so``).

This is the only door to the compute path.  There is no CPU fallback: if the library is missing
``load()`` raises, and every compute call raises ``NativeError`` when no CUDA device is present.
"""
import ctypes
import os

import numpy as np

from . import build as _build

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_RESELECTED, ERR_COMM, ERR_PEER_ABORT = 0, -1, -2, -3, -4, -5, -6
IPC_HANDLE_BYTES = 64
SPATIAL = {"none": 0, "geodesic": 1, "euclid": 2}
NORMALIZE = {"raw": 0, "exp": 1, "linear": 2}
AGGREGATE = {"sum": 0, "min": 1, "max": 2}


class NativeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"adsel_b200 native call failed (status {status}): {message}")
        self.status = status
        self.message = message


class RetrySharded(NativeError):
    """A rank's neighbour-list pool was too small (it has been grown): every rank repeats begin/end."""


class BevParams(ctypes.Structure):
    _fields_ = [("range", ctypes.c_float * 6), ("voxel", ctypes.c_float * 3), ("reserved", ctypes.c_int32)]


class SelectParams(ctypes.Structure):
    _fields_ = [
        ("spatial", ctypes.c_int32), ("temporal", ctypes.c_int32), ("normalize", ctypes.c_int32),
        ("aggregate", ctypes.c_int32), ("lambda_t", ctypes.c_double), ("max_spatial", ctypes.c_double),
        ("max_temporal", ctypes.c_double), ("cost_b", ctypes.c_double), ("cost_f", ctypes.c_double),
        ("cost_start", ctypes.c_double), ("budget", ctypes.c_double), ("forbid_sampled", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


_P = ctypes.c_void_p
_SIGNATURES = {
    "adsel_version": (ctypes.c_char_p, []),
    "adsel_last_error": (ctypes.c_char_p, []),
    "adsel_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "adsel_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_P)]),
    "adsel_destroy": (ctypes.c_int, [_P]),
    "adsel_get_stream": (ctypes.c_int, [_P, ctypes.POINTER(_P)]),
    "adsel_synchronize": (ctypes.c_int, [_P]),
    "adsel_timer_start": (ctypes.c_int, [_P]),
    "adsel_timer_stop": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_float)]),
    "adsel_set_pool": (ctypes.c_int, [_P, ctypes.c_int32, _P, _P, _P]),
    "adsel_set_pool_device": (ctypes.c_int, [_P, ctypes.c_int32, _P, _P, _P]),
    "adsel_select_device": (ctypes.c_int, [_P, ctypes.POINTER(SelectParams), _P, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)]),
    "adsel_copy_selected_device": (ctypes.c_int, [_P, _P, ctypes.c_int32]),
    "adsel_bev_set_points_device": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32]),
    "adsel_bev_descriptors_device": (ctypes.c_int, [_P, ctypes.POINTER(_P), ctypes.POINTER(_P), ctypes.POINTER(_P),
                                                    ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "adsel_build_knn_graph": (ctypes.c_int, [_P, ctypes.c_int32]),
    "adsel_get_knn": (ctypes.c_int, [_P, _P, _P]),
    "adsel_get_csr": (ctypes.c_int, [_P, _P, _P, _P, ctypes.POINTER(ctypes.c_int64)]),
    "adsel_geodesic_rows": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.c_double, _P]),
    "adsel_geodesic_max_finite": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double)]),
    "adsel_select": (ctypes.c_int, [_P, ctypes.POINTER(SelectParams), _P, ctypes.c_int32, ctypes.c_int32, _P,
                                    ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)]),
    "adsel_get_fps": (ctypes.c_int, [_P, _P]),
    "adsel_get_stage_ms": (ctypes.c_int, [_P, _P]),
    "adsel_get_counters": (ctypes.c_int, [_P, _P]),
    "adsel_get_greedy_info": (ctypes.c_int, [_P, _P]),
    "adsel_exp_f64": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P]),
    "adsel_exp_f32": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P]),
    "adsel_pairdist_set_descriptors": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.c_int32]),
    "adsel_pairdist_build": (ctypes.c_int, [_P]),
    "adsel_pairdist_get_rows": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, _P]),
    "adsel_select_dense": (ctypes.c_int, [_P, ctypes.POINTER(SelectParams), _P, _P, ctypes.c_int32, ctypes.c_int32, _P,
                                          ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)]),
    "adsel_comm_init": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_int64, _P]),
    "adsel_comm_attach": (ctypes.c_int, [_P, _P]),
    "adsel_comm_detach": (ctypes.c_int, [_P]),
    "adsel_comm_attach_local": (ctypes.c_int, [_P, ctypes.POINTER(_P)]),
    "adsel_comm_set_timeout": (ctypes.c_int, [_P, ctypes.c_double]),
    "adsel_comm_barrier": (ctypes.c_int, [_P]),
    "adsel_comm_check": (ctypes.c_int, [_P]),
    "adsel_shard_range": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "adsel_shard_prepare": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.POINTER(SelectParams), ctypes.c_int32, ctypes.c_int32]),
    "adsel_build_knn_graph_sharded": (ctypes.c_int, [_P, ctypes.c_int32]),
    "adsel_select_sharded_begin": (ctypes.c_int, [_P, ctypes.POINTER(SelectParams), _P, ctypes.c_int32, ctypes.c_int32,
                                                  ctypes.c_int32]),
    "adsel_select_replicated_begin": (ctypes.c_int, [_P, ctypes.POINTER(SelectParams), _P, ctypes.c_int32, ctypes.c_int32,
                                                     ctypes.c_int32]),
    "adsel_select_sharded_end": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32),
                                                ctypes.POINTER(ctypes.c_double)]),
    "adsel_select_fused_feature": (ctypes.c_int, [_P, ctypes.POINTER(SelectParams), ctypes.c_double, ctypes.c_int32, _P, ctypes.c_int32,
                                                  ctypes.c_int32, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32),
                                                  ctypes.POINTER(ctypes.c_double)]),
    "adsel_check_ego_locations": (ctypes.c_int, [ctypes.c_int, _P, _P, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                                 ctypes.POINTER(ctypes.c_int64)]),
    "adsel_feature_distance_build": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "adsel_pairdist_build_sharded": (ctypes.c_int, [_P]),
    "adsel_select_dense_prepare": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32]),
    "adsel_select_dense_sharded_begin": (ctypes.c_int, [_P, ctypes.POINTER(SelectParams), _P, _P, ctypes.c_int32,
                                                        ctypes.c_int32, ctypes.c_int32]),
    "adsel_select_dense_sharded_end": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32),
                                                      ctypes.POINTER(ctypes.c_double)]),
    "adsel_bev_set_points": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32]),
    "adsel_bev_generate": (ctypes.c_int, [_P, ctypes.c_int32, _P, ctypes.c_int32, ctypes.c_uint64]),
    "adsel_bev_get_points": (ctypes.c_int, [_P, _P]),
    "adsel_bev_build": (ctypes.c_int, [_P, ctypes.POINTER(BevParams), ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "adsel_bev_get_descriptors": (ctypes.c_int, [_P, _P, _P, _P, ctypes.POINTER(ctypes.c_int32)]),
}
_lib = None
HOST_INPUT_BYTES = 0        # bytes of caller arrays handed to the library from HOST memory (tests: the device-pointer entry points add none)


def library_path() -> str:
    return _build.LIB


def load():
    """Load (building in-tree if sources are newer) and type the library; raises if unavailable."""
    global _lib
    if _lib is None:
        override = os.environ.get("ADSEL_LIB")          # an alternative build of the same sources (experiments)
        if override:
            path = override
        else:
            path = _build.build() if (_build.needs_build() and _nvcc_available()) else _build.LIB
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing and nvcc is not available to build it: the selection hot path has "
                "no CPU fallback (run __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            if hasattr(lib, name):
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _nvcc_available() -> bool:
    from shutil import which
    return which(os.environ.get("NVCC", "nvcc")) is not None


def exported_symbols():
    return list(_SIGNATURES)


def _check(status):
    if status != OK:
        raise NativeError(status, load().adsel_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


def shard_bounds_256(n, world, rank):
    """Frames ``[lo, hi)`` of ``rank``: contiguous 256-aligned slices (csrc/comm.cu:update_shard)."""
    if world <= 1:
        return 0, n
    per = -(-n // world)
    per = -(-per // 256) * 256
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def bf16_bits(levels) -> np.ndarray:
    """uint16 bf16 bit patterns of small non-negative integers (exact for 0..256): the descriptor format of
    ``Handle.pairdist_set_descriptors``."""
    f = np.asarray(levels, dtype=np.float32)
    return (f.view(np.uint32) >> 16).astype(np.uint16)


def device_count() -> int:
    c = ctypes.c_int(0)
    _check(load().adsel_device_count(ctypes.byref(c)))
    return c.value


def check_ego_locations(device: int, pose9, loc_xy, fused: bool):
    """``(n_bad, first_bad)``: every frame's ego location redone on the device in the literal order of
    spatial_temporal_selector.py:83-91 and compared bit-for-bit with ``loc_xy`` (csrc/ego_check.cu)."""
    pose9 = np.ascontiguousarray(pose9, dtype=np.float64).reshape(-1, 9)
    loc_xy = np.ascontiguousarray(loc_xy, dtype=np.float64).reshape(-1, 2)
    assert len(pose9) == len(loc_xy)
    n_bad, first = ctypes.c_int64(0), ctypes.c_int64(-1)
    _check(load().adsel_check_ego_locations(int(device), _ptr(pose9), _ptr(loc_xy), len(loc_xy), int(bool(fused)),
                                            ctypes.byref(n_bad), ctypes.byref(first)))
    return n_bad.value, first.value


class Handle:
    """One selection context on one CUDA device (owns all device buffers and a stream)."""

    def __init__(self, device: int = 0):
        self._h = _P()
        self.n = 0
        self.k = 0
        self.rank, self.world = 0, 1
        _check(load().adsel_create(device, ctypes.byref(self._h)))

    def close(self):
        if self._h:
            load().adsel_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- pool / graph ----------------------------------------------------------------------
    def set_pool(self, loc_xy, group_id, n_boxes):
        loc_xy = np.ascontiguousarray(loc_xy, dtype=np.float64)
        group_id = np.ascontiguousarray(group_id, dtype=np.int32)
        n_boxes = np.ascontiguousarray(n_boxes, dtype=np.int32)
        assert loc_xy.ndim == 2 and loc_xy.shape[1] == 2 and len(group_id) == len(loc_xy) == len(n_boxes)
        self.n = len(loc_xy)
        global HOST_INPUT_BYTES
        HOST_INPUT_BYTES += loc_xy.nbytes + group_id.nbytes + n_boxes.nbytes
        _check(load().adsel_set_pool(self._h, self.n, _ptr(loc_xy), _ptr(group_id), _ptr(n_boxes)))

    def set_pool_device(self, n, loc_ptr, group_ptr, boxes_ptr):
        """``adsel_set_pool`` from DEVICE pointers (float64 [n,2], int32 [n], int32 [n] on this handle's device)."""
        self.n = int(n)
        _check(load().adsel_set_pool_device(self._h, self.n, _P(loc_ptr), _P(group_ptr), _P(boxes_ptr)))

    def select_device(self, sampled_ptr, n_sampled, *, spatial, temporal, normalize, aggregate="sum", lambda_t=1.0,
                      max_spatial=0.0, max_temporal=0.0, cost_b=0.04, cost_f=0.12, cost_start=0.0, budget=0.0,
                      forbid_sampled=False, first_index=-1):
        """``adsel_select`` with the list of already-selected frames read from a DEVICE pointer (int32); returns
        ``(m, cost)``: the m selected indices stay on the device (``copy_selected_device``)."""
        p = self._select_params(spatial, temporal, normalize, aggregate, lambda_t, max_spatial, max_temporal, cost_b, cost_f,
                                cost_start, budget, forbid_sampled)
        m, cost = ctypes.c_int32(0), ctypes.c_double(0.0)
        status = load().adsel_select_device(self._h, ctypes.byref(p), _P(sampled_ptr) if n_sampled else None, int(n_sampled),
                                            int(first_index), ctypes.byref(m), ctypes.byref(cost))
        if status == ERR_RESELECTED:
            raise AssertionError(load().adsel_last_error().decode())
        _check(status)
        return m.value, cost.value

    def copy_selected_device(self, dst_ptr, m):
        _check(load().adsel_copy_selected_device(self._h, _P(dst_ptr), int(m)))

    def build_knn_graph(self, k: int = 8):
        self.k = k
        _check(load().adsel_build_knn_graph(self._h, k))

    def get_knn(self):
        idx = np.empty((self.n, self.k + 1), dtype=np.int32)
        dist = np.empty((self.n, self.k + 1), dtype=np.float64)
        _check(load().adsel_get_knn(self._h, _ptr(idx), _ptr(dist)))
        return dist, idx

    def get_csr(self):
        nnz = ctypes.c_int64(0)
        _check(load().adsel_get_csr(self._h, None, None, None, ctypes.byref(nnz)))
        row_ptr = np.empty(self.n + 1, dtype=np.int32)
        col = np.empty(nnz.value, dtype=np.int32)
        w = np.empty(nnz.value, dtype=np.float64)
        _check(load().adsel_get_csr(self._h, _ptr(row_ptr), _ptr(col), _ptr(w), ctypes.byref(nnz)))
        return row_ptr, col, w

    # -- geodesic ----------------------------------------------------------------------------
    def geodesic_rows(self, sources, radius=np.inf):
        sources = np.ascontiguousarray(sources, dtype=np.int32)
        out = np.empty((len(sources), self.n), dtype=np.float64)
        _check(load().adsel_geodesic_rows(self._h, _ptr(sources), len(sources), float(radius), _ptr(out)))
        return out

    def geodesic_max_finite(self) -> float:
        v = ctypes.c_double(0.0)
        _check(load().adsel_geodesic_max_finite(self._h, ctypes.byref(v)))
        return v.value

    # -- selection ---------------------------------------------------------------------------
    def select(self, *, spatial, temporal, normalize, aggregate="sum", lambda_t=1.0, max_spatial=0.0,
               max_temporal=0.0, cost_b=0.04, cost_f=0.12, cost_start=0.0, budget=0.0, forbid_sampled=False,
               sampled=(), first_index=-1, out_cap=None):
        p = SelectParams(SPATIAL[spatial], int(bool(temporal)), NORMALIZE[normalize], AGGREGATE[aggregate],
                         float(lambda_t), float(max_spatial), float(max_temporal), float(cost_b), float(cost_f),
                         float(cost_start), float(budget), int(bool(forbid_sampled)), 0)
        sampled = np.ascontiguousarray(sampled, dtype=np.int32)
        out_cap = self.n if out_cap is None else out_cap
        out = np.empty(out_cap, dtype=np.int32)
        m, cost = ctypes.c_int32(0), ctypes.c_double(0.0)
        status = load().adsel_select(self._h, ctypes.byref(p), _ptr(sampled) if len(sampled) else None, len(sampled),
                                     int(first_index), _ptr(out), out_cap, ctypes.byref(m), ctypes.byref(cost))
        if status == ERR_RESELECTED:
            # the reference's AssertionError (spatial_temporal_selector.py:182), same message
            raise AssertionError(load().adsel_last_error().decode())
        _check(status)
        return out[: m.value].tolist(), cost.value

    # -- pool sharded over several ranks ---------------------------------------
    def comm_init(self, rank, world, n, k=8, dim=0, list_cap=0) -> bytes:
        """Allocate the peer-visible arena; returns this rank's 64-byte CUDA IPC handle."""
        out = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
        _check(load().adsel_comm_init(self._h, int(rank), int(world), int(n), int(k), int(dim), int(list_cap), out))
        self.rank, self.world = int(rank), int(world)
        return out.raw

    def comm_attach(self, ipc_handles):
        """``ipc_handles``: the world's handles in rank order (other processes on the same box)."""
        blob = b"".join(ipc_handles)
        assert len(blob) == IPC_HANDLE_BYTES * self.world
        _check(load().adsel_comm_attach(self._h, ctypes.c_char_p(blob)))

    def comm_detach(self):
        """Unmap every peer's arena (before any rank frees or re-sizes its own)."""
        _check(load().adsel_comm_detach(self._h))

    def comm_attach_local(self, handles):
        """``handles``: the world's ``Handle`` objects of THIS process in rank order."""
        arr = (_P * len(handles))(*[h._h for h in handles])
        _check(load().adsel_comm_attach_local(self._h, arr))

    def comm_set_timeout(self, seconds):
        _check(load().adsel_comm_set_timeout(self._h, float(seconds)))

    def comm_barrier(self):
        _check(load().adsel_comm_barrier(self._h))

    def comm_check(self):
        _check(load().adsel_comm_check(self._h))

    def shard_range(self):
        lo, hi = ctypes.c_int32(0), ctypes.c_int32(0)
        _check(load().adsel_shard_range(self._h, ctypes.byref(lo), ctypes.byref(hi)))
        return lo.value, hi.value

    def build_knn_graph_sharded(self, k: int = 8):
        self.k = k
        _check(load().adsel_build_knn_graph_sharded(self._h, k))

    @staticmethod
    def _select_params(spatial, temporal, normalize, aggregate, lambda_t, max_spatial, max_temporal, cost_b, cost_f,
                       cost_start, budget, forbid_sampled):
        return SelectParams(SPATIAL[spatial], int(bool(temporal)), NORMALIZE[normalize], AGGREGATE[aggregate],
                            float(lambda_t), float(max_spatial), float(max_temporal), float(cost_b), float(cost_f),
                            float(cost_start), float(budget), int(bool(forbid_sampled)), 0)

    def shard_prepare(self, k=8, n_sampled=0, out_cap=None, **kw):
        p = self._select_params(kw.get("spatial", "geodesic"), kw.get("temporal", True), kw.get("normalize", "exp"),
                                kw.get("aggregate", "sum"), 1.0, 0.0, 0.0, 0.04, 0.12, 0.0, 0.0, False)
        _check(load().adsel_shard_prepare(self._h, int(k), ctypes.byref(p), int(n_sampled),
                                          self.n if out_cap is None else int(out_cap)))

    def select_sharded_begin(self, *, spatial, temporal, normalize, aggregate="sum", lambda_t=1.0, max_spatial=0.0,
                             max_temporal=0.0, cost_b=0.04, cost_f=0.12, cost_start=0.0, budget=0.0,
                             forbid_sampled=False, sampled=(), first_index=-1, out_cap=None, replicate=False):
        """``replicate``: sharded build, every rank receives every list entry and runs the whole greedy loop itself."""
        p = self._select_params(spatial, temporal, normalize, aggregate, lambda_t, max_spatial, max_temporal, cost_b,
                                cost_f, cost_start, budget, forbid_sampled)
        sampled = np.ascontiguousarray(sampled, dtype=np.int32)
        self._out_cap = self.n if out_cap is None else int(out_cap)
        fn = load().adsel_select_replicated_begin if replicate else load().adsel_select_sharded_begin
        _check(fn(self._h, ctypes.byref(p), _ptr(sampled) if len(sampled) else None, len(sampled), int(first_index),
                  self._out_cap))

    def select_sharded_end(self):
        """Returns ``(selected, cost)``; raises ``RetrySharded`` when every rank must call begin/end again."""
        out = np.empty(self._out_cap, dtype=np.int32)
        m, cost = ctypes.c_int32(0), ctypes.c_double(0.0)
        status = load().adsel_select_sharded_end(self._h, _ptr(out), self._out_cap, ctypes.byref(m), ctypes.byref(cost))
        if status == ERR_RESELECTED:
            raise AssertionError(load().adsel_last_error().decode())
        message = load().adsel_last_error().decode() if status != OK else ""
        if status == ERR_PEER_ABORT or (status == ERR_CAPACITY and "neighbour-list" in message):
            raise RetrySharded(status, message)
        _check(status)
        return out[: m.value].tolist(), cost.value

    def get_fps(self):
        fps = np.empty(self.n, dtype=np.float64)
        _check(load().adsel_get_fps(self._h, _ptr(fps)))
        return fps

    def stage_ms(self):
        ms = np.zeros(8, dtype=np.float32)
        _check(load().adsel_get_stage_ms(self._h, _ptr(ms)))
        return dict(zip(("knn", "csr", "geodesic_lists", "greedy", "geodesic_max", "bev_histogram", "pairdist",
                         "dense_greedy"), ms.tolist()))

    def counters(self):
        c = np.zeros(4, dtype=np.int64)
        _check(load().adsel_get_counters(self._h, _ptr(c)))
        return dict(zip(("launches", "steps", "list_entries", "nnz"), c.tolist()))

    GREEDY_KERNELS = ("greedy_select_kernel", "greedy_fast_kernel", "greedy_resident_kernel", "greedy_lazy_kernel",
                      "fused_feature_kernel", "greedy_warp_kernel", "greedy_warp_cluster_kernel")

    def greedy_info(self):
        """Kernel of the last selection and its cluster size (csrc/greedy.cu, csrc/lazy.cu)."""
        v = np.zeros(2, dtype=np.int32)
        _check(load().adsel_get_greedy_info(self._h, _ptr(v)))
        return {"kernel": self.GREEDY_KERNELS[int(v[0])], "cluster": int(v[1])}

    # -- BEV descriptors ------------------------------------------------------------------------
    NUSC_RANGE = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)     # examples/active/cbgs_spatial_temporal.py:280

    def bev_set_points_device(self, points_ptr, frame_off):
        """``adsel_bev_set_points`` with the float32 [P,5] points read from a DEVICE pointer (frame offsets: host metadata)."""
        frame_off = np.ascontiguousarray(frame_off, dtype=np.int64)
        self.bev_batch, self.bev_points = len(frame_off) - 1, int(frame_off[-1])
        _check(load().adsel_bev_set_points_device(self._h, _P(points_ptr), _ptr(frame_off), self.bev_batch))

    def bev_descriptors_device(self):
        """Device pointers ``(counts, levels_bf16, norm2, frames, cells)`` of the descriptors of the last ``bev_build``."""
        c, l, q = _P(), _P(), _P()
        frames, cells = ctypes.c_int32(0), ctypes.c_int32(0)
        _check(load().adsel_bev_descriptors_device(self._h, ctypes.byref(c), ctypes.byref(l), ctypes.byref(q),
                                                   ctypes.byref(frames), ctypes.byref(cells)))
        return c.value, l.value, q.value, frames.value, cells.value

    def bev_set_points(self, points, frame_off):
        points = np.ascontiguousarray(points, dtype=np.float32)
        frame_off = np.ascontiguousarray(frame_off, dtype=np.int64)
        assert points.ndim == 2 and points.shape[1] == 5 and frame_off[-1] == len(points)
        self.bev_batch, self.bev_points = len(frame_off) - 1, int(frame_off[-1])
        global HOST_INPUT_BYTES
        HOST_INPUT_BYTES += points.nbytes
        _check(load().adsel_bev_set_points(self._h, _ptr(points), _ptr(frame_off), self.bev_batch))

    def bev_generate(self, frame_off, n_sweeps=10, seed=20221107, first_frame=0):
        """Synthetic LiDAR for the batch whose first frame has pool index ``first_frame``."""
        frame_off = np.ascontiguousarray(frame_off, dtype=np.int64)
        self.bev_batch, self.bev_points = len(frame_off) - 1, int(frame_off[-1])
        _check(load().adsel_bev_generate(self._h, self.bev_batch, _ptr(frame_off), n_sweeps,
                                         (int(seed) & 0xffffffff) | (int(first_frame) << 32)))

    def bev_get_points(self):
        pts = np.empty((self.bev_points, 5), dtype=np.float32)
        _check(load().adsel_bev_get_points(self._h, _ptr(pts)))
        return pts

    def bev_build(self, point_range=NUSC_RANGE, voxel=(0.8, 0.8, 8.0), aggregate_atomics=False, frame_base=0,
                  total_frames=0):
        """Descriptors of the resident point batch -> frames [frame_base, frame_base+batch) of the pool."""
        p = BevParams((ctypes.c_float * 6)(*point_range), (ctypes.c_float * 3)(*voxel), 0)
        _check(load().adsel_bev_build(self._h, ctypes.byref(p), int(bool(aggregate_atomics)), int(frame_base),
                                      int(total_frames)))
        self.bev_frames = int(total_frames) if total_frames > 0 else self.bev_batch
        self.pd_n = self.bev_frames

    def bev_get_descriptors(self, counts=True, levels=True):
        cells = ctypes.c_int32(0)
        _check(load().adsel_bev_get_descriptors(self._h, None, None, None, ctypes.byref(cells)))
        c = np.empty((self.bev_frames, cells.value), dtype=np.int32) if counts else None
        lv = np.empty((self.bev_frames, cells.value), dtype=np.uint16) if levels else None
        n2 = np.empty(self.bev_frames, dtype=np.int32)
        _check(load().adsel_bev_get_descriptors(self._h, _ptr(c), _ptr(lv), _ptr(n2), ctypes.byref(cells)))
        return c, lv, n2

    # -- pairwise distances + dense greedy ------------------------------------------------------
    def pairdist_set_descriptors(self, levels_bf16_bits):
        lv = np.ascontiguousarray(levels_bf16_bits, dtype=np.uint16)
        assert lv.ndim == 2
        self.pd_n = lv.shape[0]
        _check(load().adsel_pairdist_set_descriptors(self._h, _ptr(lv), lv.shape[0], lv.shape[1]))

    def pairdist_build(self):
        _check(load().adsel_pairdist_build(self._h))
        self.pd_cols = self.pd_n

    def feature_distance_build(self, feats, p=2):
        """float32 [n, dim] features -> the reference's N x N feature-distance map (feature_selector.py:87-109)."""
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        assert feats.ndim == 2
        self.pd_n = self.pd_cols = feats.shape[0]
        _check(load().adsel_feature_distance_build(self._h, _ptr(feats), feats.shape[0], feats.shape[1], int(p)))

    def select_fused_feature(self, *, temporal, aggregate="sum", lambda_t=1.0, lambda_f=1.0, seed_spatial_only=False,
                             max_temporal=0.0, cost_b=0.04, cost_f=0.12, cost_start=0.0, budget=0.0, forbid_sampled=False,
                             sampled=(), first_index=-1, out_cap=None):
        """Greedy loop of SpatialFeatureSelector (``temporal=False``, ``seed_spatial_only=True``) /
        SpatialTemporalFeatureSelector (``temporal=True``) on the built graph and feature map (csrc/fused.cu)."""
        p = SelectParams(SPATIAL["geodesic"], int(bool(temporal)), NORMALIZE["exp"], AGGREGATE[aggregate], float(lambda_t), 0.0,
                         float(max_temporal), float(cost_b), float(cost_f), float(cost_start), float(budget),
                         int(bool(forbid_sampled)), 0)
        sampled = np.ascontiguousarray(sampled, dtype=np.int32)
        out_cap = self.n if out_cap is None else out_cap
        out = np.empty(out_cap, dtype=np.int32)
        m, cost = ctypes.c_int32(0), ctypes.c_double(0.0)
        status = load().adsel_select_fused_feature(self._h, ctypes.byref(p), float(lambda_f), int(bool(seed_spatial_only)),
                                                   _ptr(sampled) if len(sampled) else None, len(sampled), int(first_index),
                                                   _ptr(out), out_cap, ctypes.byref(m), ctypes.byref(cost))
        if status == ERR_RESELECTED:
            raise AssertionError(load().adsel_last_error().decode())
        _check(status)
        return out[: m.value].tolist(), cost.value

    def feature_distance_rows(self, row0, rows):
        return self.pairdist_get_rows(row0, rows).view(np.float32)

    def pairdist_build_sharded(self):
        _check(load().adsel_pairdist_build_sharded(self._h))
        lo, hi = shard_bounds_256(self.pd_n, self.world, self.rank)
        self.pd_cols = hi - lo

    def select_dense_prepare(self, n_sampled=0, out_cap=None):
        _check(load().adsel_select_dense_prepare(self._h, int(n_sampled), self.pd_n if out_cap is None else int(out_cap)))

    def select_dense_sharded_begin(self, n_boxes, *, cost_b=0.04, cost_f=0.12, cost_start=0.0, budget=0.0,
                                   forbid_sampled=True, sampled=(), first_index=-1, out_cap=None):
        p = SelectParams(0, 0, 0, 0, 0.0, 0.0, 0.0, float(cost_b), float(cost_f), float(cost_start), float(budget),
                         int(bool(forbid_sampled)), 0)
        n_boxes = np.ascontiguousarray(n_boxes, dtype=np.int32)
        sampled = np.ascontiguousarray(sampled, dtype=np.int32)
        self._out_cap = len(n_boxes) if out_cap is None else int(out_cap)
        _check(load().adsel_select_dense_sharded_begin(self._h, ctypes.byref(p), _ptr(n_boxes),
                                                       _ptr(sampled) if len(sampled) else None, len(sampled),
                                                       int(first_index), self._out_cap))

    def select_dense_sharded_end(self):
        out = np.empty(self._out_cap, dtype=np.int32)
        m, cost = ctypes.c_int32(0), ctypes.c_double(0.0)
        status = load().adsel_select_dense_sharded_end(self._h, _ptr(out), self._out_cap, ctypes.byref(m), ctypes.byref(cost))
        if status == ERR_RESELECTED:
            raise AssertionError(load().adsel_last_error().decode())
        _check(status)
        return out[: m.value].tolist(), cost.value

    def pairdist_get_rows(self, row0, rows):
        out = np.empty((rows, getattr(self, "pd_cols", None) or self.pd_n), dtype=np.int32)
        _check(load().adsel_pairdist_get_rows(self._h, int(row0), int(rows), _ptr(out)))
        return out

    def select_dense(self, n_boxes, *, cost_b=0.04, cost_f=0.12, cost_start=0.0, budget=0.0, forbid_sampled=True,
                     sampled=(), first_index=-1, out_cap=None):
        p = SelectParams(0, 0, 0, 0, 0.0, 0.0, 0.0, float(cost_b), float(cost_f), float(cost_start), float(budget),
                         int(bool(forbid_sampled)), 0)
        n_boxes = np.ascontiguousarray(n_boxes, dtype=np.int32)
        sampled = np.ascontiguousarray(sampled, dtype=np.int32)
        out_cap = len(n_boxes) if out_cap is None else out_cap
        out = np.empty(out_cap, dtype=np.int32)
        m, cost = ctypes.c_int32(0), ctypes.c_double(0.0)
        status = load().adsel_select_dense(self._h, ctypes.byref(p), _ptr(n_boxes), _ptr(sampled) if len(sampled) else None,
                                           len(sampled), int(first_index), _ptr(out), out_cap, ctypes.byref(m),
                                           ctypes.byref(cost))
        if status == ERR_RESELECTED:
            raise AssertionError(load().adsel_last_error().decode())
        _check(status)
        return out[: m.value].tolist(), cost.value

    # -- one interface for the dense (BEV) path on a single GPU and on a sharded world -------------------
    @staticmethod
    def bev_cells(point_range=NUSC_RANGE, voxel=(0.8, 0.8, 8.0)) -> int:
        """ny*nx of the BEV grid: round((hi-lo)/voxel) in float32 (bev_ops.py:27-30)."""
        r, v = np.asarray(point_range, dtype=np.float32), np.asarray(voxel, dtype=np.float32)
        nx, ny = int(np.rint((r[3] - r[0]) / v[0])), int(np.rint((r[4] - r[1]) / v[1]))
        return nx * ny

    def dense_build(self, n, batch_frames, load_batch, point_range=NUSC_RANGE, voxel=(0.8, 0.8, 8.0)):
        """Descriptors of all ``n`` frames (``load_batch(handle, lo, hi)`` puts the points of frames [lo, hi) on
        the handle), then the pairwise distance matrix."""
        for lo in range(0, n, batch_frames):
            hi = min(n, lo + batch_frames)
            load_batch(self, lo, hi)
            self.bev_build(point_range, voxel, frame_base=lo, total_frames=n)
        self.pairdist_build()

    def dense_select(self, n_boxes, **kw):
        return self.select_dense(n_boxes, **kw)

    def stream_ptr(self) -> int:
        s = _P()
        _check(load().adsel_get_stream(self._h, ctypes.byref(s)))
        return s.value or 0

    def timer_start(self):
        _check(load().adsel_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = ctypes.c_float(0.0)
        _check(load().adsel_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def synchronize(self):
        _check(load().adsel_synchronize(self._h))

    def exp_f32(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        _check(load().adsel_exp_f32(self._h, _ptr(x), x.size, _ptr(out)))
        return out

    def exp_f64(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        _check(load().adsel_exp_f64(self._h, _ptr(x), x.size, _ptr(out)))
        return out
