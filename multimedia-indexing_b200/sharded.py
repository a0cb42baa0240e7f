"""Multi-GPU IVFPQ: one process per GPU of one NVSwitch node, G = S x R ranks (rank = group * S + shard).

S list shards hold one copy of the index (BASELINE.json north_star: "the code database shards by IVF list across GPUs");
R groups each serve their own query batch.  The reference keeps ONE BoundedPriorityQueue for all probed lists of a query
(IVFPQ.java:409,445); a shard's queue therefore is a partial result and the owner of a query slice merges the S partial
queues in the queue's order, with the ordered tie rule (csrc/tie_resolve.cuh) when a k-th boundary tie was cut.

The whole step runs inside libmmidx (`mmidx_search_multi_dev`, csrc/comm.cuh): the kernels store their result rows
straight into the exchange windows of the consuming ranks over NVLink (CUDA IPC peer memory) and the exchange points are
epoch flags in those windows -- no pack / unpack kernels, no collective launches, no host synchronisation.  torch /
torch.distributed is plumbing here: it carries the 128-byte window handles at set-up and provides device buffers."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._capi import check, lib
from .datastructures import IVFPQ


def _p(t):
    return C.c_void_p(t.data_ptr())


# ---- the sharding rules (host mirror of csrc/comm.cuh; tests/test_sharded_gloo.py replays them on CPU) ----------
def owner_of_list(list_id, world):
    """the default sharding rule of mmidx_params.shard_rank / shard_count"""
    return list_id % world


def balanced_shard_map(list_sizes, world):
    """list -> shard, longest-processing-time greedy on the expected scan work of a list.  A list of length n is
    probed about proportionally to n (queries follow the data) and costs n per probe: weight n^2."""
    wgt = np.asarray(list_sizes, dtype=np.float64) ** 2
    owner = np.zeros(len(wgt), dtype=np.int32)
    load = np.zeros(world)
    for l in np.argsort(-wgt, kind="stable"):
        r = int(np.argmin(load))
        owner[l] = r
        load[r] += wgt[l]
    return owner


def slice_len(gq, S):
    """queries per merge slice: shard t of a group merges the final queue of group queries [t*sl, (t+1)*sl)"""
    return (gq + S - 1) // S


def route_row(q, shard, gq, S):
    """(owner shard, row) of shard `shard`'s partial queue of group query q inside the owner's [S][sl][k] arrays
    (PeerSink mode 1, kernels.cuh sink_dest)"""
    sl = slice_len(gq, S)
    owner = q // sl
    return owner, shard * sl + (q - owner * sl)


def final_row(q, group, gq, S):
    """row of group `group`'s query q in the job-wide result arrays [R * gqp][k] (gqp = S * sl)"""
    return group * slice_len(gq, S) * S + q


class _DevArray:
    """exposes a raw device pointer to torch (as_tensor) without copying"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _view(ptr, shape, typestr, device):
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


class MultiIVFPQ:
    """One rank of an S x R job.  `index` is this rank's shard (an `IVFPQ` created with shard_rank / shard_count);
    quantizers are replicated.  Usage on every rank:
        mi = MultiIVFPQ(d, maxN, m, ks, transformation, nlist, list_shards=S)
        mi.loadCoarseQuantizer(C); mi.loadProductQuantizer(P); mi.setW(w); mi.indexAll(X)      # or indexAllDev
        mi.connect(max_gq, k_max)                                    # windows + handle rendezvous
        iids, dist, cnt, row0, nrows = mi.search(k, dQ_group, gather_all=True)
    """

    def __init__(self, vectorLength, maxNumVectors, numSubVectors, numProductCentroids, transformation,
                 numCoarseCentroids, list_shards, group=None):
        self.pg = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        S = max(1, min(int(list_shards), self.world))
        if self.world % S:
            raise ValueError(f"list_shards = {S} must divide the world size {self.world}")
        self.S, self.R = S, self.world // S
        self.shard, self.group = self.rank % S, self.rank // S
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.index = IVFPQ(vectorLength, maxNumVectors, numSubVectors, numProductCentroids, transformation,
                           numCoarseCentroids, device=self.device.index, shard_rank=self.shard, shard_count=S)
        self.shard_map = None
        self.connected = False

    def __getattr__(self, name):  # loadCoarseQuantizer, loadProductQuantizer, setW, listSizes, ...
        return getattr(self.index, name)

    # ---- indexing ----
    def indexAll(self, X, balanced=True):
        """Every rank is given the same vectors (the quantizers are replicated, so every rank computes the same codes)
        and keeps the lists it owns.  balanced: derive a load-balanced list -> shard map from the list sizes first."""
        ix = self.index
        if self.S == 1:
            return ix.indexVectors(None, X, return_codes=True)
        if not balanced:
            return ix.indexVectors(None, X, return_codes=True)
        lists, codes = ix.encode(X)
        owner = balanced_shard_map(np.bincount(lists, minlength=ix.numCoarseCentroids), self.S)
        check(lib.mmidx_set_shard_map(ix._h, C.c_void_p(owner.ctypes.data)))
        ix.indexPQCodes(None, lists, codes)
        self.shard_map = owner
        return lists, codes

    def setShardMap(self, owner):
        owner = np.ascontiguousarray(owner, dtype=np.int32)
        check(lib.mmidx_set_shard_map(self.index._h, C.c_void_p(owner.ctypes.data)))
        self.shard_map = owner

    def indexDev(self, dX, d_lists=None, d_codes=None):
        """append device-resident vectors (mmidx_add_dev); optional device outputs for the list ids / codes.
        mmidx_add_dev runs on the index's own stream: the producer of dX (torch's current stream) is synchronised first."""
        torch.cuda.current_stream().synchronize()
        check(lib.mmidx_add_dev(self.index._h, dX.shape[0], _p(dX), _p(d_lists) if d_lists is not None else None,
                                _p(d_codes) if d_codes is not None else None))

    # ---- exchange ----
    def connect(self, max_gq, k_max):
        handle = (C.c_ubyte * _capi.COMM_HANDLE_BYTES)()
        check(lib.mmidx_comm_create(self.index._h, self.rank, self.world, self.S, int(max_gq), int(k_max), handle))
        mine = bytes(handle)
        if self.world > 1:
            allh = [None] * self.world
            dist.all_gather_object(allh, mine, group=self.pg)
        else:
            allh = [mine]
        buf = b"".join(allh)
        check(lib.mmidx_comm_attach(self.index._h, C.c_char_p(buf)))
        if self.world > 1:
            dist.barrier(group=self.pg)  # nobody starts a step before every window is mapped everywhere
        self.connected = True
        self.max_gq, self.k_max = int(max_gq), int(k_max)

    def group_slice(self, nq_total):
        """the batch of this rank's group when a job-wide batch of nq_total queries is split over the R groups"""
        per = (nq_total + self.R - 1) // self.R
        q0 = min(nq_total, self.group * per)
        return q0, min(nq_total, q0 + per)

    def search(self, k, dQ, gather_all=True):
        """dQ: this GROUP's queries (float64 CUDA tensor [gq][d], the same on every shard of the group).  Asynchronous on
        the current stream.  Returns (iids, dist, cnt, row0, nrows): views of the job-wide result arrays in this rank's
        window ([R*gqp][k]; group g's query q is row g*gqp + q) and the rows this rank produced.  The views stay valid
        until the second next call."""
        assert self.connected, "connect() first"
        gq = dQ.shape[0]
        pi, pd, pc = C.c_void_p(), C.c_void_p(), C.c_void_p()
        row0, nrows = C.c_int64(), C.c_int64()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib.mmidx_search_multi_dev(self.index._h, gq, _p(dQ), int(k), 1 if gather_all else 0, C.byref(pi), C.byref(pd),
                                         C.byref(pc), C.byref(row0), C.byref(nrows), st))
        gqp = slice_len(gq, self.S) * self.S
        rows = self.R * gqp
        iids = _view(pi.value, (rows, k), "<i4", self.device)
        dd = _view(pd.value, (rows, k), "<f8", self.device)
        cnt = _view(pc.value, (rows,), "<i4", self.device)
        return iids, dd, cnt, row0.value, nrows.value

    def search_host(self, k, Q, out_iids, out_dist, out_cnt):
        """host buffers in and out (pinned for full speed): this rank's rows of its group's batch.
        Returns (first_query, nrows) of the group batch that landed in out_*[:nrows]."""
        fq, nr = C.c_int64(), C.c_int64()
        check(lib.mmidx_search_multi(self.index._h, Q.shape[0], _p(Q) if isinstance(Q, torch.Tensor) else C.c_void_p(Q.ctypes.data),
                                     int(k), _p(out_iids), _p(out_dist), _p(out_cnt), C.byref(fq), C.byref(nr)))
        return fq.value, nr.value

    def close(self):
        check(lib.mmidx_comm_destroy(self.index._h))
        self.connected = False
        self.index.close()
