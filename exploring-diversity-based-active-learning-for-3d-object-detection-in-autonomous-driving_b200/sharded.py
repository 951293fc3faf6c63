"""Pool sharded over the GPUs of one NVSwitch box.

The reference has no multi-GPU selection: under ``torch.distributed.launch`` every rank recomputes the
whole selection and rank 0 writes (det3d/selectors/base_selector.py:57, tools/active_select.py:94-101).
Here every rank owns a contiguous slice of the pool's columns.  All data-path communication happens
inside the CUDA kernels over peer-mapped memory (kNN rows stored into every rank's table, one
``(value, index)`` mailbox exchange per greedy step); ``torch.distributed`` is used only to hand the
64-byte CUDA IPC handles around once.

Two drivers with the interface of ``native.Handle`` (``set_pool``, ``build_knn_graph``,
``geodesic_max_finite``, ``select``, ``stage_ms``, ``counters``, ``close``):

* ``DistWorld``  -- one process per GPU (``torchrun``); what the selectors use when
  ``torch.distributed`` is initialised with more than one rank.
* ``LocalWorld`` -- all ranks driven by ONE host thread, on one device or several; used by the tests
  (the exchange protocol runs unchanged between CTAs of one GPU) and for single-process multi-GPU.
"""
import os

import numpy as np

from . import native


def shard_bounds(n: int, world: int, rank: int):
    """Frames ``[lo, hi)`` owned by ``rank``: contiguous, 256-aligned slices (mirrors csrc/comm.cu:update_shard)."""
    return native.shard_bounds_256(n, world, rank)


REPLICATE_MAX_FRAMES = 440_000      # 16 CTAs x 27.5 k columns: what a thread-block cluster's shared memory holds


def shard_mode(n: int) -> str:
    """How several ranks split one selection.

    ``replicate`` (default while the pool fits one GPU's resident greedy kernel): the kNN rows and the neighbour lists
    are BUILT by the rank that owns the source frames and stored into every rank's arena over NVLink; each rank then
    runs the whole greedy loop in its own shared memory.  A resident step costs about as much as ONE NVLink round trip,
    so exchanging a (value, index) pair per step -- ``columns``, the round-1 schedule, kept for pools beyond 440 k
    frames -- can only lose (measured: 9.8 ms on one GPU, 14.5 ms on eight).
    ``ADSEL_SHARD=columns|replicate`` forces one; ``ADSEL_SHARD=0`` makes every rank an independent replica."""
    env = os.environ.get("ADSEL_SHARD", "auto")
    if env in ("columns", "replicate"):
        return env
    return "replicate" if n <= REPLICATE_MAX_FRAMES else "columns"


def lazy_rows(n: int, kw) -> bool:
    """Mirrors csrc/lazy.cu:lazy_wanted -- rows without a saturation radius on a large pool are searched per selected
    frame by every rank on its own (no lists to shard)."""
    if kw.get("spatial") != "geodesic" or kw.get("normalize") == "exp":
        return False
    env = os.environ.get("ADSEL_LAZY", "")
    if env:
        return env != "0"
    return n >= 16384


def pick_global_best(records):
    """The exchange's reduction rule on the host: ``records`` = ``[(value_key, index), ...]`` per rank ->
    the record with the largest key, lowest index among equal keys (np.argmax's first maximum)."""
    best = None
    for key, idx in records:
        if best is None or key > best[0] or (key == best[0] and idx < best[1]):
            best = (key, idx)
    return best


class _WorldBase:
    MAX_RETRIES = 4

    def __init__(self):
        self.n = 0
        self.k = 0
        self._pool = None
        self._pool_dirty = False
        self._arena_shape = None
        self._list_cap = 0          # entries of every rank's neighbour-list region (0: library default)

    def set_pool(self, loc_xy, group_id, n_boxes):
        self._pool = (np.ascontiguousarray(loc_xy, dtype=np.float64), np.ascontiguousarray(group_id, dtype=np.int32),
                      np.ascontiguousarray(n_boxes, dtype=np.int32))
        self.n = len(self._pool[0])
        self._pool_dirty = True


class LocalWorld(_WorldBase):
    """``world`` ranks of this process; ``devices[r]`` is rank r's CUDA device (default: all on device 0)."""

    def __init__(self, world: int, devices=None, timeout: float = 20.0, mode: str = "columns"):
        super().__init__()
        self.mode = mode
        self.world = int(world)
        devices = [0] * self.world if devices is None else list(devices)
        assert len(devices) == self.world
        self.handles = [native.Handle(d) for d in devices]
        self.timeout = timeout

    def close(self):
        for h in self.handles:
            h.close()
        self.handles = []

    def _ensure(self, k: int, dim: int = 0):
        if self.mode == "replicate" and self._list_cap == 0 and k > 0:
            self._list_cap = self.n * 288            # every rank holds every entry
        shape = (self.n, int(k), int(dim), self._list_cap)
        if self._arena_shape != shape:
            for r, h in enumerate(self.handles):
                h.comm_init(r, self.world, self.n, k, dim, self._list_cap)
                h.comm_set_timeout(self.timeout)
            for h in self.handles:
                h.comm_attach_local(self.handles)
            self._arena_shape = shape
            self._pool_dirty = True
        if self._pool_dirty and self._pool is not None:
            for h in self.handles:
                h.set_pool(*self._pool)
            self._pool_dirty = False

    # -- dense (BEV descriptor) path ---------------------------------------------------------------
    def dense_build(self, n, batch_frames, load_batch, point_range=native.Handle.NUSC_RANGE, voxel=(0.8, 0.8, 8.0)):
        """Every rank histograms the frames of its own slice (``load_batch(handle, lo, hi)``) and stores the
        descriptors into every rank's arena; then each rank computes the distance tiles of its column slice."""
        self.n = n
        self._ensure(self.k, dim=native.Handle.bev_cells(point_range, voxel))
        for r, h in enumerate(self.handles):
            lo, hi = shard_bounds(n, self.world, r)
            for b in range(lo, hi, batch_frames):
                e = min(hi, b + batch_frames)
                load_batch(h, b, e)
                h.bev_build(point_range, voxel, frame_base=b, total_frames=n)
        for h in self.handles:
            h.pd_n = n
            h.pairdist_build_sharded()
        for h in self.handles:
            h.comm_check()

    def dense_set_descriptors(self, levels_bf16_bits):
        """Caller-supplied descriptors (tests): every rank uploads the full matrix into its own arena."""
        self.n = levels_bf16_bits.shape[0]
        self._ensure(self.k, dim=levels_bf16_bits.shape[1])
        for h in self.handles:
            h.pairdist_set_descriptors(levels_bf16_bits)
        for h in self.handles:
            h.pairdist_build_sharded()
        for h in self.handles:
            h.comm_check()

    def dense_select(self, n_boxes, **kw):
        for h in self.handles:
            h.select_dense_prepare(len(kw.get("sampled", ())), kw.get("out_cap"))
        for h in self.handles:
            h.select_dense_sharded_begin(n_boxes, **kw)
        results, error = [], None
        for h in self.handles:
            try:
                results.append(h.select_dense_sharded_end())
            except (AssertionError, native.NativeError) as e:
                error = error or e
        if error is not None:
            raise error
        assert all(r == results[0] for r in results), "ranks disagree on the selection"
        return results[0]

    def shard_ranges(self):
        return [h.shard_range() for h in self.handles]

    def build_knn_graph(self, k: int = 8):
        self.k = k
        self._ensure(k)
        for h in self.handles:       # allocations first: none may happen while a rank's kernel waits for its peers
            h.shard_prepare(k=k)
        for h in self.handles:
            h.build_knn_graph_sharded(k)
        for h in self.handles:
            h.comm_check()

    def geodesic_max_finite(self) -> float:
        vals = [h.geodesic_max_finite() for h in self.handles]
        assert all(v == vals[0] for v in vals)
        return vals[0]

    def select(self, **kw):
        self._ensure(self.k)
        sampled = kw.get("sampled", ())
        prep = {key: kw[key] for key in ("spatial", "temporal", "normalize", "aggregate") if key in kw}
        for _ in range(self.MAX_RETRIES):
            for h in self.handles:
                h.shard_prepare(k=max(self.k, 1), n_sampled=len(sampled), out_cap=kw.get("out_cap"), **prep)
            for h in self.handles:
                h.select_sharded_begin(replicate=self.mode == "replicate", **kw)
            results, retry, error = [], False, None
            for h in self.handles:   # every rank is collected even if one raises (their kernels are all in flight)
                try:
                    results.append(h.select_sharded_end())
                except native.RetrySharded:
                    retry = True
                except (AssertionError, native.NativeError) as e:
                    error = error or e
            if error is not None and not retry:
                raise error
            if retry:       # a rank's list region overflowed: new arenas sized for the largest need, graph rebuilt
                self._list_cap = int(max(h.counters()["list_entries"] for h in self.handles) * 1.1) + 1024
                self.build_knn_graph(self.k)
                continue
            assert all(r == results[0] for r in results), "ranks disagree on the selection"
            return results[0]
        raise native.NativeError(native.ERR_CAPACITY, "sharded selection kept overflowing its neighbour-list pools")

    def get_fps(self):
        """The min-distance vector assembled from every rank's slice."""
        out = np.empty(self.n, dtype=np.float64)
        for h in self.handles:
            lo, hi = h.shard_range()
            out[lo:hi] = h.get_fps()[lo:hi]
        return out

    def stage_ms(self):
        return self.handles[0].stage_ms()

    def counters(self):
        c = self.handles[0].counters()
        c["launches"] = sum(h.counters()["launches"] for h in self.handles)
        return c


class DistWorld(_WorldBase):
    """This process's rank of a ``torch.distributed`` world whose ranks all sit on one NVSwitch box."""

    def __init__(self, device: int, group=None, timeout: float = 20.0):
        super().__init__()
        import torch.distributed as dist
        assert dist.is_available() and dist.is_initialized(), "DistWorld needs an initialised torch.distributed"
        self._dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("the sharded path spans one 8-GPU NVSwitch box (world <= 8); set ADSEL_SHARD=0 for replicas")
        self.handle = native.Handle(device)
        self.device = device
        self.timeout = timeout

    def close(self):
        if self.handle is not None:
            self.handle.close()
            self.handle = None

    def _ensure(self, k: int, dim: int = 0):
        self.mode = shard_mode(self.n)
        if self.mode == "replicate" and self._list_cap == 0 and k > 0:
            self._list_cap = self.n * 288            # every rank holds every entry
        shape = (self.n, int(k), int(dim), self._list_cap)
        if self._arena_shape != shape:
            if self._arena_shape is not None:       # nobody may still map an arena that is about to be freed
                self.handle.comm_detach()
                self._dist.barrier(group=self.group)
            ipc = self.handle.comm_init(self.rank, self.world, self.n, k, dim, self._list_cap)
            self.handle.comm_set_timeout(self.timeout)
            gathered = [None] * self.world
            self._dist.all_gather_object(gathered, ipc, group=self.group)   # 64 bytes per rank, once per pool shape
            self.handle.comm_attach(gathered)
            self._arena_shape = shape
            self._pool_dirty = True
        if self._pool_dirty and self._pool is not None:
            self.handle.set_pool(*self._pool)
            self._pool_dirty = False

    # -- dense (BEV descriptor) path ---------------------------------------------------------------
    def dense_build(self, n, batch_frames, load_batch, point_range=native.Handle.NUSC_RANGE, voxel=(0.8, 0.8, 8.0)):
        """This rank histograms the frames of its own slice and stores their descriptors into every rank's
        arena (fused all-gather over NVLink); then the distance tiles of its column slice."""
        self.n = n
        self._ensure(self.k, dim=native.Handle.bev_cells(point_range, voxel))
        lo, hi = shard_bounds(n, self.world, self.rank)
        for b in range(lo, hi, batch_frames):
            e = min(hi, b + batch_frames)
            load_batch(self.handle, b, e)
            self.handle.bev_build(point_range, voxel, frame_base=b, total_frames=n)
        self.handle.pd_n = n
        self.handle.pairdist_build_sharded()

    def dense_select(self, n_boxes, **kw):
        self.handle.select_dense_sharded_begin(n_boxes, **kw)
        out = self.handle.select_dense_sharded_end()
        self.handle.comm_check()        # a device barrier of the build that timed out left partial peer data behind
        return out

    def shard_range(self):
        return self.handle.shard_range()

    def build_knn_graph(self, k: int = 8):
        self.k = k
        self._ensure(k)
        self.handle.build_knn_graph_sharded(k)

    def geodesic_max_finite(self) -> float:
        return self.handle.geodesic_max_finite()

    def select(self, **kw):
        from . import dist as _d
        self._ensure(self.k)
        if lazy_rows(self.n, kw):            # one search per selected frame: nothing to shard, every rank a replica
            return self.handle.select(**kw)
        for _ in range(self.MAX_RETRIES):
            self.handle.select_sharded_begin(replicate=self.mode == "replicate", **kw)
            out, code, error = None, 0, None
            try:
                out = self.handle.select_sharded_end()
                self.handle.comm_check()     # a timed-out device barrier (graph / list build) left partial peer data
            except native.RetrySharded:
                code = 1
            except (AssertionError, native.NativeError) as e:
                code, error = 2, e
            # every rank takes the same branch: a rank that failed alone must not leave its peers in a collective
            worst = _d.all_ranks_max(code)
            if worst == 2:
                raise error if error is not None else native.NativeError(native.ERR_PEER_ABORT, "a peer rank failed the selection")
            if worst == 0:
                return out
            need = _d.all_ranks_max(self.handle.counters()["list_entries"])
            self._list_cap = int(need * 1.1) + 1024
            self.build_knn_graph(self.k)     # new arenas sized for the largest need, graph rebuilt
        raise native.NativeError(native.ERR_CAPACITY, "sharded selection kept overflowing its neighbour-list pools")

    def get_fps(self):
        return self.handle.get_fps()

    def stage_ms(self):
        return self.handle.stage_ms()

    def counters(self):
        return self.handle.counters()

    def greedy_info(self):
        return self.handle.greedy_info()


_ENGINES = {}


class _Lease:
    """An engine borrowed from the per-process cache: ``close()`` returns it instead of destroying it, so that
    repeated selections (active-learning rounds in one process, the bench) reuse the CUDA context's buffers,
    the peer mappings and the loaded kernels rather than paying cudaMalloc / cudaFree / IPC setup every call."""

    def __init__(self, key, engine):
        self._key, self._engine = key, engine

    def __getattr__(self, name):
        return getattr(self._engine, name)

    def close(self):
        pass


def make_engine(device: int):
    """``native.Handle`` for a single process, ``DistWorld`` when torch.distributed runs more than one rank
    (``ADSEL_SHARD=0`` keeps the reference's behaviour: every rank selects on its own replica).
    Engines are cached per (kind, device); ``release_engines()`` destroys them."""
    from .dist import get_dist_info
    _, world = get_dist_info()
    sharded_mode = world > 1 and os.environ.get("ADSEL_SHARD", "1") != "0"
    key = ("dist" if sharded_mode else "single", int(device), world)
    if key not in _ENGINES:
        _ENGINES[key] = DistWorld(device) if sharded_mode else native.Handle(device)
    return _Lease(key, _ENGINES[key])


def release_engines():
    for engine in _ENGINES.values():
        engine.close()
    _ENGINES.clear()


import atexit  # noqa: E402

atexit.register(release_engines)
