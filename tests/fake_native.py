"""Test double of ``native.Handle`` for the CPU (gloo) tests of the multi-rank host logic.

It is TEST INFRASTRUCTURE: rows come from the oracle (oracle/restatement.py) and the per-step exchange of the
CUDA kernels' peer-memory mailboxes is replaced by ``torch.distributed.all_gather_object`` followed by the
same reduction rule (``sharded.pick_global_best``).  The product never imports this file.
"""
import struct

import numpy as np
import torch.distributed as dist

import restatement as R
from adsel_b200 import native, sharded


EGO_CHECKS = []     # (device, frames, fused) of every all-frames ego-location check the host asked for


def check_ego_locations(device, pose9, loc_xy, fused):
    """Stand-in for csrc/ego_check.cu: the 3-term dot product in numpy's separately rounded order, else accepted as the
    fused order (python has no fma before 3.13) -- enough for the host logic, which only needs (n_bad, first_bad)."""
    pose9 = np.asarray(pose9, dtype=np.float64).reshape(-1, 9)
    loc_xy = np.asarray(loc_xy, dtype=np.float64).reshape(-1, 2)
    assert len(pose9) == len(loc_xy)
    EGO_CHECKS.append((int(device), len(loc_xy), bool(fused)))
    if not fused:
        t = pose9[:, [2, 5, 8]]
        want = np.stack([-((t[:, 0] * pose9[:, j] + t[:, 1] * pose9[:, 3 + j]) + t[:, 2] * pose9[:, 6 + j]) for j in (0, 1)], axis=1)
        bad = np.flatnonzero((want.view(np.uint64) != loc_xy.view(np.uint64)).any(axis=1))
        return len(bad), (int(bad[0]) if len(bad) else -1)
    approx = np.stack([-(pose9[:, 2] * pose9[:, j] + pose9[:, 5] * pose9[:, 3 + j] + pose9[:, 8] * pose9[:, 6 + j]) for j in (0, 1)], axis=1)
    bad = np.flatnonzero(~np.isclose(approx, loc_xy, rtol=1e-12, atol=1e-9).all(axis=1))
    return len(bad), (int(bad[0]) if len(bad) else -1)


class FakeShardHandle:
    instances = []
    overflow_once = False      # tests: the first select of every handle reports a neighbour-list overflow

    def __init__(self, device=0):
        self.device, self.n, self.k = device, 0, 0
        self.rank, self.world = 0, 1
        self.attached = False
        self.csr = None
        self.calls = []
        FakeShardHandle.instances.append(self)

    def close(self):
        self.calls.append("close")

    # -- comm ---------------------------------------------------------------------------------
    def comm_init(self, rank, world, n, k=8, dim=0, list_cap=0):
        self.rank, self.world, self.arena = rank, world, (n, k, dim)
        self.attached = world == 1
        self.calls.append(("comm_init", rank, world, n, k, dim))
        return struct.pack("<4i", rank, world, n, k).ljust(native.IPC_HANDLE_BYTES, b"\0")

    def comm_detach(self):
        self.attached = False

    def comm_attach(self, handles):
        assert len(handles) == self.world and all(len(h) == native.IPC_HANDLE_BYTES for h in handles)
        for r, blob in enumerate(handles):                      # rank order, every rank sized its arena identically
            assert struct.unpack("<4i", blob[:16]) == (r, self.world, self.arena[0], self.arena[1])
        self.attached = True
        self.calls.append("comm_attach")

    def comm_set_timeout(self, seconds):
        self.timeout = seconds

    def shard_range(self):
        return sharded.shard_bounds(self.n, self.world, self.rank)

    # -- pool / graph ---------------------------------------------------------------------------
    def set_pool(self, loc, group, n_boxes):
        self.loc, self.group, self.n_boxes = np.asarray(loc), np.asarray(group), np.asarray(n_boxes)
        self.n = len(self.loc)
        self.csr = None

    def build_knn_graph_sharded(self, k=8):
        assert self.attached
        self.k = k
        self.csr = R.knn_graph_csr(self.loc, k)

    def geodesic_max_finite(self):
        return R.geodesic_max_finite(self.csr)

    # -- selection --------------------------------------------------------------------------------
    def _row(self, i, kw):
        if kw["spatial"] == "geodesic":
            s = R.geodesic_rows(self.csr, [i])[0]
        elif kw["spatial"] == "euclid":
            return R.euclid_row(self.loc, self.group, i)
        else:
            return R.temporal_row(self.group, i)
        if not kw["temporal"]:
            return s
        t = R.temporal_row(self.group, i)
        if kw["normalize"] == "linear":
            s, t = s / kw["max_spatial"], t / kw["max_temporal"]
        else:
            s, t = R.exp_normalize(s), R.exp_normalize(t)
        agg = kw.get("aggregate", "sum")
        if agg == "sum":
            return s + kw.get("lambda_t", 1.0) * t
        return np.stack([s, t]).min(0) if agg == "min" else np.stack([s, t]).max(0)

    def select_sharded_begin(self, replicate=False, **kw):
        assert self.attached
        self.replicate = replicate          # the schedule (who runs the loop) does not change what is selected
        self.pending = kw

    def comm_check(self):
        self.calls.append(("comm_check",))

    def _argmax_exchange(self, fps, lo):
        if len(fps):
            j = int(np.argmax(fps))
            local = (float(fps[j]), lo + j)
        else:
            local = (-1.0, 2 ** 31 - 1)
        records = [None] * self.world
        dist.all_gather_object(records, local)                   # stands in for the mailbox exchange
        return sharded.pick_global_best(records)[1]

    def select_sharded_end(self):
        if FakeShardHandle.overflow_once and not getattr(self, "_overflowed", False):
            self._overflowed = True                              # like the device: rank 0 overflows, the others see its abort
            dist.barrier()
            raise native.RetrySharded(native.ERR_CAPACITY if self.rank == 0 else native.ERR_PEER_ABORT, "neighbour-list pool too small")
        kw = self.pending
        lo, hi = self.shard_range()
        sampled = list(kw.get("sampled", ()))
        cost_b, cost_f = kw.get("cost_b", 0.04), kw.get("cost_f", 0.12)
        cost = kw.get("cost_start", 0.0)
        if sampled:
            fps = np.stack([self._row(i, kw)[lo:hi] for i in sampled]).min(0)
            selected = [self._argmax_exchange(fps, lo)]
        else:
            selected = [int(kw["first_index"])]
            fps = self._row(selected[-1], kw)[lo:hi]
        chosen = set(selected) | (set(sampled) if kw.get("forbid_sampled") else set())
        cost += cost_f
        cost += int(self.n_boxes[selected[-1]]) * cost_b
        while True:
            fps = np.minimum(fps, self._row(selected[-1], kw)[lo:hi])
            nxt = self._argmax_exchange(fps, lo)
            assert nxt not in chosen, f"id: {nxt} has been selected"
            cost += cost_f
            cost += int(self.n_boxes[nxt]) * cost_b
            if cost > kw["budget"]:
                break
            selected.append(nxt)
            chosen.add(nxt)
        return selected, cost

    def stage_ms(self):
        return {}

    def counters(self):
        return {"launches": 0, "list_entries": 1000 + 10 * self.rank}
