"""Multi-GPU IVFPQ: one process per GPU, the inverted lists sharded by `list id % world == rank`
(BASELINE.json north_star).  The reference keeps one BoundedPriorityQueue for all probed lists
(IVFPQ.java:409,445); here every rank scans the probed lists it owns, the per-rank top-k (+ offer sequence
numbers) are all-gathered over NCCL and merged on the device with the queue's ordering; exact ties cut at the
k-th boundary go through the ordered tie pass.  torch / torch.distributed is plumbing only (device buffers,
stream, the collective); every arithmetic step is a libmmidx kernel."""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _capi
from ._capi import check, lib
from .datastructures import IVFPQ


def _p(t):
    return C.c_void_p(t.data_ptr())


PACK_FIELDS = (("iids", torch.int32, True), ("dist", torch.float64, True), ("seq", torch.int64, True),
               ("tie", torch.float64, False), ("cnt", torch.int32, False))


def packed_layout(nq, k):
    """Byte layout of one rank's packed result row: [iids i32[nq][k] | dist f64 | seq i64 | tie f64[nq] | cnt i32[nq]],
    every field 16-byte aligned, so ONE all-gather moves a rank's whole partial result."""
    sizes = [nq * (k if per_k else 1) * torch.empty(0, dtype=dt).element_size() for _, dt, per_k in PACK_FIELDS]
    offs = [0]
    for s in sizes:
        offs.append((offs[-1] + s + 15) & ~15)
    return offs, sizes


def gather_partials(local, world, nq, k, group=None, out=None):
    """all-gather the packed rows and return {field: tensor[world][nq][k] or [world][nq]} (contiguous copies)."""
    offs, sizes = packed_layout(nq, k)
    gathered = torch.empty(world * offs[-1], dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    g = gathered.view(world, -1)
    parts = out if out is not None else {}
    for i, (name, dt, per_k) in enumerate(PACK_FIELDS):
        shape = (world, nq, k) if per_k else (world, nq)
        if name not in parts:
            parts[name] = torch.empty(shape, dtype=dt, device=local.device)
        parts[name].view(torch.uint8).view(world, -1).copy_(g[:, offs[i]:offs[i] + sizes[i]])
    return parts


def owner_of_list(list_id, world):
    """the default sharding rule of mmidx_params.shard_rank / shard_count"""
    return list_id % world


def balanced_shard_map(list_sizes, world):
    """list -> shard, longest-processing-time greedy on the expected scan work of a list.  A list of length n is
    probed about proportionally to n (queries follow the data) and costs n per probe: weight n^2."""
    import numpy as np

    wgt = np.asarray(list_sizes, dtype=np.float64) ** 2
    owner = np.zeros(len(wgt), dtype=np.int32)
    load = np.zeros(world)
    for l in np.argsort(-wgt, kind="stable"):
        r = int(np.argmin(load))
        owner[l] = r
        load[r] += wgt[l]
    return owner


class ShardedIVFPQ:
    """One rank's shard of an IVFPQ index + the collective search.

    search(k, dQ) per batch (G ranks, nq queries, sl = ceil(nq/G)):
      0. coarse probe of this rank's query slice, all-gather of the probe lists      (coarse cost / G)
      1. mmidx_search_shard_dev: scan of the probed lists this rank owns, all nq queries -> partial top-k
      2. all-to-all: rank r receives every rank's partials for query slice r          (one packed row per peer)
      3. mmidx_merge_topk_dev on the slice (queue order, tie detection)               (merge cost / G)
      4. all-gather of the merged slices: every rank ends with the full result
      5. only if some query's k-th boundary was an exact tie that got cut: ordered tie pass (rare)
    """

    def __init__(self, vectorLength, maxNumVectors, numSubVectors, numProductCentroids, transformation,
                 numCoarseCentroids, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.index = IVFPQ(vectorLength, maxNumVectors, numSubVectors, numProductCentroids, transformation,
                           numCoarseCentroids, device=self.device.index, shard_rank=self.rank, shard_count=self.world)
        self._bufs = {}

    def __getattr__(self, name):  # loadCoarseQuantizer, loadProductQuantizer, setW, indexVectors, ...
        return getattr(self.index, name)

    def indexVectorsBalanced(self, X):
        """Encode everything (every rank does, the quantizers are replicated), derive a load-balanced list -> shard
        map from the list sizes (identical on all ranks), then store the lists this rank owns. Returns (lists, codes)."""
        import numpy as np

        lists, codes = self.index.encode(X)
        owner = balanced_shard_map(np.bincount(lists, minlength=self.index.numCoarseCentroids), self.world)
        check(lib.mmidx_set_shard_map(self.index._h, C.c_void_p(owner.ctypes.data)))
        self.index.indexPQCodes(None, lists, codes)
        self.shard_map = owner
        return lists, codes

    def _buffers(self, nq, k, w):
        key = (nq, k, w)
        if key not in self._bufs:
            dev, G = self.device, self.world
            sl = (nq + G - 1) // G
            nqp = sl * G
            offs, sizes = packed_layout(sl, k)  # one destination slice
            b = dict(sl=sl, nqp=nqp, offs=offs, sizes=sizes)
            b["probes_local"] = torch.zeros((sl, w), dtype=torch.int32, device=dev)
            b["probes_all"] = torch.zeros((nqp, w), dtype=torch.int32, device=dev)
            # partial results of this shard for all (padded) queries
            b["part"] = {name: torch.zeros((nqp, k) if per_k else (nqp,), dtype=dt, device=dev)
                         for name, dt, per_k in PACK_FIELDS}
            b["send"] = torch.empty(G * offs[-1], dtype=torch.uint8, device=dev)
            b["recv"] = torch.empty(G * offs[-1], dtype=torch.uint8, device=dev)
            b["parts"] = {name: torch.empty((G, sl, k) if per_k else (G, sl), dtype=dt, device=dev)
                          for name, dt, per_k in PACK_FIELDS}
            b["slice"] = dict(iids=torch.empty((sl, k), dtype=torch.int32, device=dev),
                              dist=torch.empty((sl, k), dtype=torch.float64, device=dev),
                              cnt=torch.empty(sl, dtype=torch.int32, device=dev),
                              amb_list=torch.zeros(sl, dtype=torch.int32, device=dev),
                              amb_count=torch.zeros(1, dtype=torch.int32, device=dev))
            fo, fs = [0], []
            for n_el, es in ((sl * k, 4), (sl * k, 8), (sl, 4), (sl + 1, 4)):  # iids | dist | cnt | amb (count + list)
                fs.append(n_el * es)
                fo.append((fo[-1] + fs[-1] + 15) & ~15)
            b["fo"], b["fs"] = fo, fs
            b["fin_local"] = torch.empty(fo[-1], dtype=torch.uint8, device=dev)
            b["fin_all"] = torch.empty(G * fo[-1], dtype=torch.uint8, device=dev)
            b["out"] = dict(iids=torch.empty((nqp, k), dtype=torch.int32, device=dev),
                            dist=torch.empty((nqp, k), dtype=torch.float64, device=dev),
                            cnt=torch.empty(nqp, dtype=torch.int32, device=dev),
                            amb=torch.empty((G, sl + 1), dtype=torch.int32, device=dev))
            self._bufs[key] = b
        return self._bufs[key]

    def search_dev(self, k, dQ):
        """dQ: float64 CUDA tensor [nq][d], the same on every rank. Returns (iids[nq][k], dist[nq][k], cnt[nq], state);
        every rank gets the full merged result. Asynchronous on the current stream."""
        nq = dQ.shape[0]
        G, rank, ix = self.world, self.rank, self.index
        w = ix.w
        b = self._buffers(nq, k, w)
        sl, nqp, offs, sizes = b["sl"], b["nqp"], b["offs"], b["sizes"]
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        marks = [] if os.environ.get("MMIDX_SHARD_TIMING") else None

        def mark(name):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        # 0. coarse probes of my query slice, all-gathered
        q0 = min(nq, rank * sl)
        q1 = min(nq, q0 + sl)
        if q1 > q0:
            check(lib.mmidx_coarse_probe_dev(ix._h, q1 - q0, C.c_void_p(dQ.data_ptr() + q0 * dQ.shape[1] * 8), w,
                                             _p(b["probes_local"]), st))
        mark("coarse")
        dist.all_gather_into_tensor(b["probes_all"].view(-1), b["probes_local"].view(-1), group=self.group)
        mark("allgather_probes")
        # 1. this shard's partial top-k for all queries
        part = b["part"]
        check(lib.mmidx_search_shard_dev(ix._h, nq, _p(dQ), k, _p(b["probes_all"]), _p(part["iids"]), _p(part["dist"]),
                                         _p(part["seq"]), _p(part["tie"]), _p(part["cnt"]), st))
        mark("shard_search")
        # 2. all-to-all of the packed query slices
        send = b["send"].view(G, -1)
        for i, (name, dt, per_k) in enumerate(PACK_FIELDS):
            send[:, offs[i]:offs[i] + sizes[i]].copy_(part[name].view(torch.uint8).view(G, -1))
        mark("pack")
        dist.all_to_all_single(b["recv"], b["send"], group=self.group)
        mark("all_to_all")
        recv = b["recv"].view(G, -1)
        parts = b["parts"]
        for i, (name, dt, per_k) in enumerate(PACK_FIELDS):
            parts[name].view(torch.uint8).view(G, -1).copy_(recv[:, offs[i]:offs[i] + sizes[i]])
        # 3. merge my slice
        sc = b["slice"]
        check(lib.mmidx_merge_topk_dev(sl, k, G, _p(parts["iids"]), _p(parts["dist"]), _p(parts["seq"]), _p(parts["tie"]),
                                       _p(parts["cnt"]), _p(sc["iids"]), _p(sc["dist"]), None, _p(sc["cnt"]),
                                       _p(sc["amb_list"]), _p(sc["amb_count"]), st))
        mark("unpack+merge")
        # 4. all-gather the merged slices (+ each slice's ambiguous-query list)
        fo, fs, fl = b["fo"], b["fs"], b["fin_local"]
        fl[fo[0]:fo[0] + fs[0]].copy_(sc["iids"].view(torch.uint8).view(-1))
        fl[fo[1]:fo[1] + fs[1]].copy_(sc["dist"].view(torch.uint8).view(-1))
        fl[fo[2]:fo[2] + fs[2]].copy_(sc["cnt"].view(torch.uint8).view(-1))
        fl[fo[3]:fo[3] + 4].copy_(sc["amb_count"].view(torch.uint8).view(-1))
        fl[fo[3] + 4:fo[3] + fs[3]].copy_(sc["amb_list"].view(torch.uint8).view(-1))
        mark("pack2")
        dist.all_gather_into_tensor(b["fin_all"], fl, group=self.group)
        mark("allgather_final")
        fa = b["fin_all"].view(G, -1)
        out = b["out"]
        out["iids"].view(torch.uint8).view(G, -1).copy_(fa[:, fo[0]:fo[0] + fs[0]])
        out["dist"].view(torch.uint8).view(G, -1).copy_(fa[:, fo[1]:fo[1] + fs[1]])
        out["cnt"].view(torch.uint8).view(G, -1).copy_(fa[:, fo[2]:fo[2] + fs[2]])
        out["amb"].view(torch.uint8).view(G, -1).copy_(fa[:, fo[3]:fo[3] + fs[3]])
        mark("unpack2")
        if marks is not None:
            torch.cuda.synchronize()
            self.last_phase_ms = {n: marks[i - 1][1].elapsed_time(e) for i, (n, e) in enumerate(marks) if i > 0}
        return out["iids"][:nq], out["dist"][:nq], out["cnt"][:nq], b

    def resolve_ties(self, k, dQ, b):
        """Rare path: exact binary64 ties cut at the k-th boundary. Host-syncs on the ambiguous counts."""
        amb = b["out"]["amb"].cpu()
        counts = amb[:, 0].tolist()
        if sum(counts) == 0:
            return 0
        nq = dQ.shape[0]
        dev, G, sl = self.device, self.world, b["sl"]
        glob = [int(amb[r, 1 + i]) + r * sl for r in range(G) for i in range(counts[r])]
        na = len(glob)
        amb_list = torch.tensor(glob, dtype=torch.int32, device=dev)
        amb_count = torch.tensor([na], dtype=torch.int32, device=dev)
        out = b["out"]
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        nqp = b["nqp"]
        l_seq = torch.zeros((nqp, k), dtype=torch.int64, device=dev)
        l_iid = torch.zeros((nqp, k), dtype=torch.int32, device=dev)
        l_eq = torch.zeros((nqp, k), dtype=torch.int32, device=dev)
        l_cnt = torch.zeros(nqp, dtype=torch.int32, device=dev)
        check(lib.mmidx_tie_collect_shard_dev(self.index._h, nq, _p(dQ), k, _p(out["dist"]), _p(amb_list), _p(amb_count),
                                              _p(l_seq), _p(l_iid), _p(l_eq), _p(l_cnt), st))
        g = [torch.empty((G,) + t.shape, dtype=t.dtype, device=dev) for t in (l_seq, l_iid, l_eq, l_cnt)]
        for dst, src in zip(g, (l_seq, l_iid, l_eq, l_cnt)):
            dist.all_gather_into_tensor(dst, src, group=self.group)
        check(lib.mmidx_tie_finish_dev(nqp, k, G, _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(amb_list), _p(amb_count),
                                       _p(out["iids"]), _p(out["dist"]), st))
        return na

    def search(self, k, dQ):
        iids, d, cnt, b = self.search_dev(k, dQ)
        self.resolve_ties(k, dQ, b)
        return iids, d, cnt


class HybridIVFPQ:
    """G = S x R ranks: S list shards (the NCCL exchange above) x R query groups (each group owns nq/R of a batch).

    Rank r is list shard r % S of query group r // S.  S == G is pure list sharding (what a database that does not fit
    one GPU needs); S == 1 is pure query parallelism over replicas (best for a database as small as 12 MB, where the
    per-query fixed work of a shard does not shrink with S).  Every rank ends with the full result."""

    def __init__(self, vectorLength, maxNumVectors, numSubVectors, numProductCentroids, transformation,
                 numCoarseCentroids, list_shards):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        S = max(1, min(int(list_shards), self.world))
        while self.world % S:
            S -= 1
        self.S, self.R = S, self.world // S
        self.qgroup = self.rank // S
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.sharded = None
        if S > 1:
            groups = [dist.new_group(ranks=list(range(g * S, (g + 1) * S))) for g in range(self.R)]
            self.sharded = ShardedIVFPQ(vectorLength, maxNumVectors, numSubVectors, numProductCentroids, transformation,
                                        numCoarseCentroids, group=groups[self.qgroup])
            self.index = self.sharded.index
        else:
            self.index = IVFPQ(vectorLength, maxNumVectors, numSubVectors, numProductCentroids, transformation,
                               numCoarseCentroids, device=self.device.index)
        self._bufs = {}

    def __getattr__(self, name):
        return getattr(self.index, name)

    def indexAll(self, X):
        if self.sharded is not None:
            return self.sharded.indexVectorsBalanced(X)
        return self.index.indexVectors(None, X, return_codes=True)

    def slice_of(self, nq):
        sl = (nq + self.R - 1) // self.R
        q0 = min(nq, self.qgroup * sl)
        return q0, min(nq, q0 + sl), sl

    def search(self, k, dQ):
        """dQ: all nq queries (device, float64). Returns full (iids[nq][k], dist[nq][k], cnt[nq]) on every rank."""
        nq = dQ.shape[0]
        q0, q1, sl = self.slice_of(nq)
        key = (nq, k)
        if key not in self._bufs:
            dev = self.device
            fo, fs = [0], []
            for n_el, es in ((sl * k, 4), (sl * k, 8), (sl, 4)):
                fs.append(n_el * es)
                fo.append((fo[-1] + fs[-1] + 15) & ~15)
            self._bufs[key] = dict(fo=fo, fs=fs, loc=torch.zeros(fo[-1], dtype=torch.uint8, device=dev),
                                   all=torch.empty(self.world * fo[-1], dtype=torch.uint8, device=dev),
                                   iids=torch.empty((self.R * sl, k), dtype=torch.int32, device=dev),
                                   dist=torch.empty((self.R * sl, k), dtype=torch.float64, device=dev),
                                   cnt=torch.empty(self.R * sl, dtype=torch.int32, device=dev),
                                   s_iids=torch.full((sl, k), -1, dtype=torch.int32, device=dev),
                                   s_dist=torch.full((sl, k), float("inf"), dtype=torch.float64, device=dev),
                                   s_cnt=torch.zeros(sl, dtype=torch.int32, device=dev))
        b = self._bufs[key]
        n_loc = q1 - q0
        dq = dQ[q0:q1]
        if n_loc > 0:
            if self.sharded is not None:
                i, d, c = self.sharded.search(k, dq)
                b["s_iids"][:n_loc].copy_(i)
                b["s_dist"][:n_loc].copy_(d)
                b["s_cnt"][:n_loc].copy_(c)
            else:
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                check(lib.mmidx_search_dev(self.index._h, n_loc, _p(dq), k, _p(b["s_iids"]), _p(b["s_dist"]), _p(b["s_cnt"]), st))
        if self.R == 1:
            return b["s_iids"][:nq], b["s_dist"][:nq], b["s_cnt"][:nq]
        if self.S == 1:
            # every rank is a whole query group: its slice is one contiguous block of each output array, so the three
            # results are all-gathered in place (no pack / unpack kernels around the collective)
            dist.all_gather_into_tensor(b["iids"].view(-1), b["s_iids"].view(-1))
            dist.all_gather_into_tensor(b["dist"].view(-1), b["s_dist"].view(-1))
            dist.all_gather_into_tensor(b["cnt"], b["s_cnt"])
            return b["iids"][:nq], b["dist"][:nq], b["cnt"][:nq]
        fo, fs, loc = b["fo"], b["fs"], b["loc"]
        loc[fo[0]:fo[0] + fs[0]].copy_(b["s_iids"].view(torch.uint8).view(-1))
        loc[fo[1]:fo[1] + fs[1]].copy_(b["s_dist"].view(torch.uint8).view(-1))
        loc[fo[2]:fo[2] + fs[2]].copy_(b["s_cnt"].view(torch.uint8).view(-1))
        dist.all_gather_into_tensor(b["all"], loc)
        ga = b["all"].view(self.R, self.S, -1)[:, 0, :]  # one list shard of every query group holds the group's result
        b["iids"].view(torch.uint8).view(self.R, -1).copy_(ga[:, fo[0]:fo[0] + fs[0]])
        b["dist"].view(torch.uint8).view(self.R, -1).copy_(ga[:, fo[1]:fo[1] + fs[1]])
        b["cnt"].view(torch.uint8).view(self.R, -1).copy_(ga[:, fo[2]:fo[2] + fs[2]])
        return b["iids"][:nq], b["dist"][:nq], b["cnt"][:nq]
