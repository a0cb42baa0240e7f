"""Host-side mirror of gr.iti.mklab.visual.datastructures (Linear, PQ, IVFPQ) over libmmidx's C ABI.

Same names, argument meaning and error behaviour as the reference classes, so the parity tests read like
tests of the reference (J/ = src/main/java/gr/iti/mklab/visual/):
  AbstractSearchStructure.indexVector(id, vector) -> bool            J/datastructures/AbstractSearchStructure.java:229-257
  AbstractSearchStructure.computeNearestNeighbors(k, vector|id)      :281-291, :320-328
  getInternalId / getId / isIndexed / getLoadCounter / close         :383, :403, :537, :711, :734
  IVFPQ.setW / loadCoarseQuantizer / loadProductQuantizer / indexPQCode   J/datastructures/IVFPQ.java:95, :297, :275, :357
The BDB JE id<->iid databases of the reference are out of scope: ids live in a Python dict.
Every arithmetic step runs in the sm_100a kernels; there is no CPU path here.
Batch overloads (indexVectors / computeNearestNeighborsBatch) are the form a GPU wants; the single-item
methods call them with a batch of one.
"""
import ctypes as C
import time

import numpy as np

from . import _capi
from ._capi import MmidxError, check, lib


class Answer:
    """J/utilities/Answer.java:45-58 (times in ms)."""

    def __init__(self, ids, distances, nameLookupTime, indexSearchTime):
        self._ids, self._distances = ids, distances
        self._nameLookupTime, self._indexSearchTime = nameLookupTime, indexSearchTime

    def getIds(self):
        return self._ids

    def getDistances(self):
        return self._distances

    def getIndexSearchTime(self):
        return self._indexSearchTime

    def getNameLookupTime(self):
        return self._nameLookupTime


class TransformationType:
    """PQ.java:30-32"""

    None_ = "None"
    RandomRotation = "RandomRotation"
    RandomPermutation = "RandomPermutation"


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def random_permutation(seed, dim):
    """J/utilities/RandomPermutation.java:29-40: java.util.Random(seed) + Collections.shuffle.
    Pure integer host logic (no arithmetic on vectors), so it lives on the host side like in the reference."""
    mask = (1 << 48) - 1
    s = (seed ^ 0x5DEECE66D) & mask

    def nxt(bits):
        nonlocal s
        s = (s * 0x5DEECE66D + 0xB) & mask
        return s >> (48 - bits)  # (int)(seed >>> (48 - bits)); non-negative for the bits = 31 used here

    def next_int(bound):
        r = nxt(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if u - r + m < (1 << 31):
                return r
            u = nxt(31)

    perm = list(range(dim))
    for i in range(dim, 1, -1):
        j = next_int(i)
        perm[i - 1], perm[j] = perm[j], perm[i - 1]
    return np.asarray(perm, dtype=np.int32)


class AbstractSearchStructure:
    _type = None

    def __init__(self, vectorLength, maxNumVectors, m=0, ks=0, nlist=0, w=0, device=-1, shard_rank=0, shard_count=0):
        self.vectorLength = int(vectorLength)
        self.maxNumVectors = int(maxNumVectors)
        self._h = C.c_void_p()
        p = _capi.Params(self._type, vectorLength, maxNumVectors, m, ks, nlist, w, device, shard_rank, shard_count)
        check(lib.mmidx_create(C.byref(p), C.byref(self._h)))
        self._id_to_iid = {}
        self._iid_to_id = []
        self.totalInternalVectorIndexingTime = 0

    # ---- ids (the reference's BDB maps, ASS.java:555-562) ----
    def isIndexed(self, id):
        return id in self._id_to_iid

    def getInternalId(self, id):
        return self._id_to_iid.get(id, -1)

    def getId(self, iid):
        if iid < 0 or iid >= len(self._iid_to_id):
            return None
        return self._iid_to_id[iid]

    def getLoadCounter(self):
        n = C.c_int64()
        check(lib.mmidx_size(self._h, C.byref(n)))
        return n.value

    # ---- indexing ----
    def indexVector(self, id, vector):
        """ASS.java:229-257: False when the index is full or the id is already indexed; raises on a wrong
        dimensionality (IVFPQ.java:310-312, PQ.java:233-235, Linear.java:112-114)."""
        if self.getLoadCounter() >= self.maxNumVectors:
            return False
        if self.isIndexed(id):
            return False
        vector = np.asarray(vector, dtype=np.float64)
        if vector.ndim != 1 or vector.shape[0] != self.vectorLength:
            raise MmidxError(_capi.ERR_DIM, "The dimensionality of the vector is wrong!")
        self._add(vector.reshape(1, -1))
        self._map([id])
        return True

    def indexVectors(self, ids, vectors, return_codes=False):
        """Batch form of indexVector: vectors[n][d]; ids must be new and distinct. Returns (list ids, codes)
        when return_codes, so a host can persist them like IVFPQ.appendPersistentIndex (IVFPQ.java:760-772)."""
        X = _f64(vectors)
        if X.ndim != 2 or X.shape[1] != self.vectorLength:
            raise MmidxError(_capi.ERR_DIM, "The dimensionality of the vector is wrong!")
        if ids is None:
            base = len(self._iid_to_id)
            ids = [str(base + i) for i in range(X.shape[0])]
        if len(ids) != X.shape[0]:
            raise MmidxError(_capi.ERR_INVALID, "ids and vectors differ in length")
        if len(set(ids)) != len(ids) or any(i in self._id_to_iid for i in ids):
            raise MmidxError(_capi.ERR_INVALID, "duplicate id")
        out = self._add(X, return_codes)
        self._map(ids)
        return out

    def _map(self, ids):
        for id in ids:
            self._id_to_iid[id] = len(self._iid_to_id)
            self._iid_to_id.append(id)

    def _add(self, X, return_codes=False):
        t0 = time.perf_counter_ns()
        n = X.shape[0]
        lists = codes = None
        if return_codes and self._type != _capi.MMIDX_LINEAR:
            lists = np.empty(n, dtype=np.int32) if self._type == _capi.MMIDX_IVFPQ else None
            codes = np.empty((n, self.numSubVectors), dtype=np.uint8 if self.numProductCentroids <= 256 else np.uint16)
        check(lib.mmidx_add(self._h, n, _ptr(X), _ptr(lists), _ptr(codes)))
        self.totalInternalVectorIndexingTime += time.perf_counter_ns() - t0
        return (lists, codes) if return_codes else None

    # ---- search ----
    def computeNearestNeighbors(self, k, query):
        """ASS.java:281-291 (vector) and :320-328 (id of an indexed vector)."""
        if isinstance(query, str):
            iid = self.getInternalId(query)
            if iid == -1:
                raise MmidxError(_capi.ERR_INVALID, "Id does not exist!")  # ASS.java:322-324
            return self._answer(k, self._search_by_iid(k, iid))
        q = np.asarray(query, dtype=np.float64).reshape(1, -1)
        return self._answer(k, self.searchBatch(k, q))

    def _search_by_iid(self, k, iid):
        raise MmidxError(_capi.ERR_UNSUPPORTED, "query-by-id is not available for this index type")

    def _answer(self, k, res):
        iids, dist, cnt, ms = res
        t0 = time.perf_counter()
        n = int(cnt[0])
        ids = [self.getId(int(i)) for i in iids[0, :n]]
        look = int((time.perf_counter() - t0) * 1000)
        return Answer(ids, dist[0, :n].copy(), look, int(ms))

    def searchBatch(self, k, Q):
        """nq queries at once: (iids[nq][k], dist[nq][k], count[nq], ms). Unused slots: iid -1, dist +inf."""
        Q = _f64(Q)
        if Q.ndim != 2 or Q.shape[1] != self.vectorLength:
            raise MmidxError(_capi.ERR_DIM, "The dimensionality of the vector is wrong!")
        k = int(k)
        nq = Q.shape[0]
        iids = np.empty((nq, max(k, 0)), dtype=np.int32)
        dist = np.empty((nq, max(k, 0)), dtype=np.float64)
        cnt = np.zeros(nq, dtype=np.int32)
        t0 = time.perf_counter()
        check(lib.mmidx_search(self._h, nq, _ptr(Q), k, _ptr(iids), _ptr(dist), _ptr(cnt)))
        return iids, dist, cnt, (time.perf_counter() - t0) * 1000

    def computeNearestNeighborsBatch(self, k, Q):
        iids, dist, cnt, ms = self.searchBatch(k, Q)
        out = []
        for r in range(Q.shape[0] if hasattr(Q, "shape") else len(Q)):
            n = int(cnt[r])
            out.append(Answer([self.getId(int(i)) for i in iids[r, :n]], dist[r, :n].copy(), 0, int(ms)))
        return out

    def lastTimings(self):
        t = np.zeros(5, dtype=np.float32)
        check(lib.mmidx_last_timings(self._h, _ptr(t)))
        return dict(zip(("coarse_ms", "lut_ms", "scan_ms", "merge_ms", "total_ms"), map(float, t)))

    def lastTimingsMulti(self):
        t = np.zeros(8, dtype=np.float32)
        check(lib.mmidx_last_timings_multi(self._h, _ptr(t)))
        return dict(zip(("coarse_ms", "prep_ms", "scan_ms", "after_scan_ms", "total_ms", "exchange_points_ms", "merge_ms", "tie_pass_ms"),
                        map(float, t)))

    def enableTimings(self, on=True):
        check(lib.mmidx_enable_timings(self._h, 1 if on else 0))

    def lastLaunches(self):
        n = C.c_int32()
        check(lib.mmidx_last_launches(self._h, C.byref(n)))
        return n.value

    def debugStats(self):
        """MMIDX_STATS=1 counters of the fused scan kernel: candidates, re-scanned lists, survivors, direct fallbacks"""
        out = np.zeros(4, dtype=np.uint64)
        check(lib.mmidx_debug_stats(self._h, _ptr(out)))
        return out

    def scanBytes(self, Q):
        Q = _f64(Q)
        out = C.c_int64()
        check(lib.mmidx_scan_bytes(self._h, Q.shape[0], _ptr(Q), C.byref(out)))
        return out.value

    def outputIndexingTimes(self):
        n = max(self.getLoadCounter(), 1)
        print(f"{self.totalInternalVectorIndexingTime / 1e6 / n} ms => internal indexing time")

    def close(self):
        if self._h:
            lib.mmidx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Linear(AbstractSearchStructure):
    """J/datastructures/Linear.java: exact squared-L2 kNN (computeNearestNeighborsInternal :138-163)."""

    _type = _capi.MMIDX_LINEAR

    def __init__(self, vectorLength, maxNumVectors, device=-1):
        super().__init__(vectorLength, maxNumVectors, device=device)

    def getVector(self, iid):
        """Linear.java:253-281"""
        v = np.empty(self.vectorLength, dtype=np.float64)
        check(lib.mmidx_get_vector(self._h, int(iid), _ptr(v)))
        return v

    def _search_by_iid(self, k, iid):
        # Linear.java:181-184: fetch the stored vector, then search by vector
        return self.searchBatch(k, self.getVector(iid).reshape(1, -1))


class _PQBase(AbstractSearchStructure):
    def _init_pq(self, numSubVectors, numProductCentroids, transformation, rotation=None):
        self.numSubVectors = int(numSubVectors)
        self.numProductCentroids = int(numProductCentroids)
        self.subVectorLength = self.vectorLength // self.numSubVectors
        self.transformation = transformation
        if transformation == TransformationType.RandomPermutation:
            # PQ.java:155-156: new RandomPermutation(seed = 1, vectorLength)
            self.setPermutation(random_permutation(1, self.vectorLength))
        elif transformation == TransformationType.RandomRotation:
            # PQ.java:153-154: new RandomRotation(seed, vectorLength) draws the matrix with EJML's
            # RandomMatrices.createOrthogonal -- un-vendored third-party code, so the d x d matrix has to
            # be supplied (e.g. dumped once from a JVM); it is then applied exactly as RandomRotation.rotate does
            if rotation is None:
                raise MmidxError(_capi.ERR_UNSUPPORTED, "RandomRotation needs EJML's matrix: pass rotation=R[d][d]")
            self.setRotation(rotation)

    def setRotation(self, R):
        """TransformationType.RandomRotation with a supplied matrix: transformed = v R (RandomRotation.java:44-49)"""
        R = _f64(R, (self.vectorLength, self.vectorLength))
        check(lib.mmidx_set_transform(self._h, 1, None, _ptr(R)))

    def setPermutation(self, perm):
        perm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        check(lib.mmidx_set_permutation(self._h, _ptr(perm)))

    def loadProductQuantizer(self, source):
        """PQ.java:210-223 / IVFPQ.java:275-292. `source`: CSV file (m*ks lines of subVectorLength values) or
        an array [m][ks][subVectorLength]."""
        if isinstance(source, str):
            rows = [[float(x) for x in line.strip().split(",")] for line in open(source) if line.strip()]
            P = np.asarray(rows[: self.numSubVectors * self.numProductCentroids], dtype=np.float64)
        else:
            P = _f64(source)
        P = _f64(P, (self.numSubVectors, self.numProductCentroids, self.subVectorLength))
        check(lib.mmidx_set_product_quantizer(self._h, _ptr(P)))

    def encode(self, X):
        """Arithmetic of indexVectorInternal without the append: (list ids or None, raw codes)."""
        X = _f64(X)
        n = X.shape[0]
        lists = np.empty(n, dtype=np.int32) if self._type == _capi.MMIDX_IVFPQ else None
        codes = np.empty((n, self.numSubVectors), dtype=np.uint8 if self.numProductCentroids <= 256 else np.uint16)
        check(lib.mmidx_encode(self._h, n, _ptr(X), _ptr(lists), _ptr(codes)))
        return lists, codes

    def computeLookupADC(self, V):
        """PQ.java:387-399: ADC tables [nq][m][ks] of already transformed vectors."""
        V = _f64(V)
        V = V.reshape(-1, self.vectorLength)
        out = np.empty((V.shape[0], self.numSubVectors, self.numProductCentroids), dtype=np.float64)
        check(lib.mmidx_pq_lut(self._h, V.shape[0], _ptr(V), _ptr(out)))
        return out


class PQ(_PQBase):
    """J/datastructures/PQ.java: product quantization index, exhaustive ADC search (computeKnnADC :290-322)."""

    _type = _capi.MMIDX_PQ

    def __init__(self, vectorLength, maxNumVectors, numSubVectors, numProductCentroids,
                 transformation=TransformationType.None_, device=-1, rotation=None):
        super().__init__(vectorLength, maxNumVectors, m=numSubVectors, ks=numProductCentroids, device=device)
        self._init_pq(numSubVectors, numProductCentroids, transformation, rotation)

    def indexPQCodes(self, ids, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8 if self.numProductCentroids <= 256 else np.uint16)
        check(lib.mmidx_add_codes(self._h, codes.shape[0], None, _ptr(codes)))
        if ids is None:
            base = len(self._iid_to_id)
            ids = [str(base + i) for i in range(codes.shape[0])]
        self._map(ids)


class IVFPQ(_PQBase):
    """J/datastructures/IVFPQ.java: IVFADC (computeKnnIVFADC :408-450)."""

    _type = _capi.MMIDX_IVFPQ

    def __init__(self, vectorLength, maxNumVectors, numSubVectors, numProductCentroids,
                 transformation=TransformationType.None_, numCoarseCentroids=1, device=-1, shard_rank=0, shard_count=0,
                 rotation=None):
        super().__init__(vectorLength, maxNumVectors, m=numSubVectors, ks=numProductCentroids,
                         nlist=numCoarseCentroids, w=0, device=device, shard_rank=shard_rank, shard_count=shard_count)
        self.numCoarseCentroids = int(numCoarseCentroids)
        self.w = int(numCoarseCentroids * 0.1)  # IVFPQ.java:188
        self._init_pq(numSubVectors, numProductCentroids, transformation, rotation)

    def setW(self, w):
        """IVFPQ.java:95-97"""
        self.w = int(w)
        check(lib.mmidx_set_w(self._h, int(w)))

    def loadCoarseQuantizer(self, source):
        """IVFPQ.java:297-300 -> AbstractFeatureAggregator.readQuantizer (AFA.java:234-254): one centroid per
        line, lines without a comma are skipped."""
        if isinstance(source, str):
            rows = [[float(x) for x in line.strip().split(",")] for line in open(source) if "," in line]
            Cq = np.asarray(rows[: self.numCoarseCentroids], dtype=np.float64)
        else:
            Cq = _f64(source)
        Cq = _f64(Cq, (self.numCoarseCentroids, self.vectorLength))
        check(lib.mmidx_set_coarse_quantizer(self._h, _ptr(Cq)))

    def indexPQCode(self, id, listId, code):
        """IVFPQ.java:357-386: insert a pre-computed code (Java bytes are code-128; `code` here is the signed
        byte array exactly as the reference passes it)."""
        if self.numProductCentroids > 256:
            raise MmidxError(_capi.ERR_INVALID, "Call the short[] variant of the method!")  # IVFPQ.java:358-361
        if self.getLoadCounter() >= self.maxNumVectors or self.isIndexed(id):
            return False
        raw = (np.asarray(code, dtype=np.int16) + 128).astype(np.uint8).reshape(1, -1)
        self.indexPQCodes([id], np.asarray([listId], dtype=np.int32), raw)
        return True

    def indexPQCodes(self, ids, listIds, codes):
        """Bulk (re)load, IVFPQ.loadIndexInMemory IVFPQ.java:680-728: raw codes 0..ks-1."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8 if self.numProductCentroids <= 256 else np.uint16)
        listIds = np.ascontiguousarray(listIds, dtype=np.int32)
        check(lib.mmidx_add_codes(self._h, codes.shape[0], _ptr(listIds), _ptr(codes)))
        if ids is None:
            base = len(self._iid_to_id)
            ids = [str(base + i) for i in range(codes.shape[0])]
        self._map(ids)

    def computeNearestCoarseIndices(self, Q, w=None):
        """IVFPQ.java:575-601 for a batch: [nq][w], ascending coarse distance."""
        Q = _f64(Q).reshape(-1, self.vectorLength)
        w = self.w if w is None else int(w)
        out = np.empty((Q.shape[0], max(w, 0)), dtype=np.int32)
        check(lib.mmidx_coarse_probe(self._h, Q.shape[0], _ptr(Q), w, _ptr(out)))
        return out

    def listSizes(self):
        """IVFPQ.outputItemsPerList IVFPQ.java:654-673"""
        out = np.zeros(self.numCoarseCentroids, dtype=np.int32)
        check(lib.mmidx_list_sizes(self._h, _ptr(out)))
        return out
