"""Multi-GPU IVFPQ: one process per GPU, the inverted lists sharded by `list id % world == rank`
(BASELINE.json north_star).  The reference keeps one BoundedPriorityQueue for all probed lists
(IVFPQ.java:409,445); here every rank scans the probed lists it owns, the per-rank top-k (+ offer sequence
numbers) are all-gathered over NCCL and merged on the device with the queue's ordering; exact ties cut at the
k-th boundary go through the ordered tie pass.  torch / torch.distributed is plumbing only (device buffers,
stream, the collective); every arithmetic step is a libmmidx kernel."""
import ctypes as C

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
    """the sharding rule of mmidx_params.shard_rank / shard_count"""
    return list_id % world


class ShardedIVFPQ:
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

    def _buffers(self, nq, k):
        key = (nq, k)
        if key not in self._bufs:
            dev = self.device
            offs, sizes = packed_layout(nq, k)
            local = torch.empty(offs[-1], dtype=torch.uint8, device=dev)
            out = dict(iids=torch.empty((nq, k), dtype=torch.int32, device=dev),
                       dist=torch.empty((nq, k), dtype=torch.float64, device=dev),
                       cnt=torch.empty(nq, dtype=torch.int32, device=dev),
                       amb_list=torch.empty(nq, dtype=torch.int32, device=dev),
                       amb_count=torch.zeros(1, dtype=torch.int32, device=dev))
            self._bufs[key] = (offs, local, out, {})
        return self._bufs[key]

    def search_dev(self, k, dQ):
        """dQ: float64 CUDA tensor [nq][d] (nq <= 32768). Returns device tensors (iids, dist, cnt); every rank
        gets the full merged result."""
        nq = dQ.shape[0]
        offs, local, out, parts = self._buffers(nq, k)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        lp = {f[0]: C.c_void_p(local.data_ptr() + offs[i]) for i, f in enumerate(PACK_FIELDS)}
        check(lib.mmidx_search_shard_dev(self.index._h, nq, _p(dQ), k, lp["iids"], lp["dist"], lp["seq"], lp["tie"],
                                         lp["cnt"], st))
        parts = gather_partials(local, self.world, nq, k, self.group, parts)
        check(lib.mmidx_merge_topk_dev(nq, k, self.world, _p(parts["iids"]), _p(parts["dist"]), _p(parts["seq"]),
                                       _p(parts["tie"]), _p(parts["cnt"]), _p(out["iids"]), _p(out["dist"]), None,
                                       _p(out["cnt"]), _p(out["amb_list"]), _p(out["amb_count"]), st))
        return out["iids"], out["dist"], out["cnt"], out

    def resolve_ties(self, k, dQ, out):
        """Rare path: exact binary64 ties cut at the k-th boundary. Host-syncs on the ambiguous count."""
        na = int(out["amb_count"].item())
        if na == 0:
            return 0
        nq = dQ.shape[0]
        dev, W = self.device, self.world
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        l_seq = torch.zeros((nq, k), dtype=torch.int64, device=dev)
        l_iid = torch.zeros((nq, k), dtype=torch.int32, device=dev)
        l_eq = torch.zeros((nq, k), dtype=torch.int32, device=dev)
        l_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
        check(lib.mmidx_tie_collect_shard_dev(self.index._h, nq, _p(dQ), k, _p(out["dist"]), _p(out["amb_list"]),
                                              _p(out["amb_count"]), _p(l_seq), _p(l_iid), _p(l_eq), _p(l_cnt), st))
        g = [torch.empty((W,) + t.shape, dtype=t.dtype, device=dev) for t in (l_seq, l_iid, l_eq, l_cnt)]
        for dst, src in zip(g, (l_seq, l_iid, l_eq, l_cnt)):
            dist.all_gather_into_tensor(dst, src, group=self.group)
        check(lib.mmidx_tie_finish_dev(nq, k, W, _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(out["amb_list"]),
                                       _p(out["amb_count"]), _p(out["iids"]), _p(out["dist"]), st))
        return na

    def search(self, k, dQ):
        iids, d, cnt, out = self.search_dev(k, dQ)
        self.resolve_ties(k, dQ, out)
        return iids, d, cnt
