"""Independent pure-Python restatement of the reference's search / encode arithmetic, written from the Java
(not from oracle/mmidx_oracle.c) and used to pin the C oracle on small cases and to generate tests/golden/.
Python floats are IEEE binary64 and `a*b+c` is never fused, i.e. the Java numeric model.

PARITY UNPINNED by the reference itself: it has no tests or fixtures and cannot run here (no JVM)."""
import bisect


class BoundedPriorityQueue:
    """com.aliasi.util.BoundedPriorityQueue<Result> (LingPipe 4.0.1) with J/utilities/Result.java:38-45.
    A TreeSet of entries ordered comparator-larger first (= smaller distance first); entries that compare
    equal are ordered by creation id, the EARLIER entry sorting LATER.  offer(): if full and
    compare(e, last()) <= 0 reject, else add and drop last()."""

    def __init__(self, max_size):
        if max_size < 1:
            raise ValueError("Require maximum size >= 1")
        self.max_size = max_size
        self.keys = []  # (distance, -creation id) ascending == TreeSet iteration order
        self.ids = []
        self.next_id = 0

    def offer(self, rid, dist):
        if len(self.keys) >= self.max_size:
            last_dist = self.keys[-1][0]
            # Result.compare(e, last): -1 if e.dist > last.dist, 1 if <, else 0; reject when <= 0
            if not dist < last_dist:
                return False
        key = (dist, -self.next_id)
        self.next_id += 1
        pos = bisect.bisect_left(self.keys, key)
        self.keys.insert(pos, key)
        self.ids.insert(pos, rid)
        if len(self.keys) > self.max_size:
            self.keys.pop()
            self.ids.pop()
        return True

    def last_distance(self):
        return self.keys[-1][0]

    def size(self):
        return len(self.keys)

    def results(self):
        return list(self.ids), [k[0] for k in self.keys]


def linear_search(X, q, k):
    """Linear.computeNearestNeighborsInternal Linear.java:138-163"""
    nn = BoundedPriorityQueue(k)
    lowest = 1.7976931348623157e308
    for i, x in enumerate(X):
        skip = False
        l2 = 0.0
        for j in range(len(q)):
            l2 += (q[j] - x[j]) * (q[j] - x[j])
            if l2 > lowest:
                skip = True
                break
        if not skip:
            nn.offer(i, l2)
            if i >= k:
                lowest = nn.last_distance()
    return nn.results()


def lookup_adc(P, v):
    """PQ.computeLookupADC PQ.java:387-399; P[m][ks][S]"""
    m, ks, S = len(P), len(P[0]), len(P[0][0])
    lut = [[0.0] * ks for _ in range(m)]
    for i in range(m):
        start = i * S
        for j in range(ks):
            acc = 0.0
            for t in range(S):
                acc += (v[start + t] - P[i][j][t]) * (v[start + t] - P[i][j][t])
            lut[i][j] = acc
    return lut


def nearest_product_index(P, sub, subvec):
    """PQ.computeNearestProductIndex PQ.java:411-429"""
    best, min_d = -1, 1.7976931348623157e308
    for i in range(len(P[sub])):
        dist = 0.0
        for j in range(len(subvec)):
            dist += (P[sub][i][j] - subvec[j]) * (P[sub][i][j] - subvec[j])
            if dist >= min_d:
                break
        if dist < min_d:
            min_d, best = dist, i
    return best


def permute(perm, v):
    """RandomPermutation.permute RandomPermutation.java:50-56"""
    return [v[p] for p in perm] if perm is not None else list(v)


def pq_encode(P, x, perm=None):
    """PQ.indexVectorInternal PQ.java:232-268 -> raw centroid indices (the Java byte is index - 128)"""
    S = len(P[0][0])
    v = permute(perm, x)
    return [nearest_product_index(P, i, v[i * S:(i + 1) * S]) for i in range(len(P))]


def pq_search(P, codes, q, k, perm=None):
    """PQ.computeKnnADC PQ.java:290-322"""
    nn = BoundedPriorityQueue(k)
    lut = lookup_adc(P, permute(perm, q))
    for i, code in enumerate(codes):
        l2 = 0.0
        for j in range(len(code)):
            l2 += lut[j][code[j]]
        nn.offer(i, l2)
    return nn.results()


def nearest_coarse_index(C, v):
    """IVFPQ.computeNearestCoarseIndex IVFPQ.java:547-564"""
    best, min_d = -1, 1.7976931348623157e308
    for i, c in enumerate(C):
        dist = 0.0
        for j in range(len(v)):
            dist += (c[j] - v[j]) * (c[j] - v[j])
            if dist >= min_d:
                break
        if dist < min_d:
            min_d, best = dist, i
    return best


def nearest_coarse_indices(C, v, k):
    """IVFPQ.computeNearestCoarseIndices IVFPQ.java:575-601"""
    bpq = BoundedPriorityQueue(k)
    lowest = 1.7976931348623157e308
    for i, c in enumerate(C):
        skip = False
        l2 = 0.0
        for j in range(len(v)):
            l2 += (c[j] - v[j]) * (c[j] - v[j])
            if l2 > lowest:
                skip = True
                break
        if not skip:
            bpq.offer(i, l2)
            if i >= k:
                lowest = bpq.last_distance()
    ids, _ = bpq.results()
    if len(ids) < k:
        raise RuntimeError("NullPointerException: poll() on an exhausted queue (IVFPQ.java:598)")
    return ids


def residual(C, v, l):
    """IVFPQ.computeResidualVector IVFPQ.java:642-648: centroid MINUS vector"""
    return [C[l][i] - v[i] for i in range(len(v))]


def ivfpq_encode(C, P, x, perm=None):
    """IVFPQ.indexVectorInternal IVFPQ.java:309-355 -> (list id, raw code)"""
    l = nearest_coarse_index(C, x)
    v = permute(perm, residual(C, x, l))
    S = len(P[0][0])
    return l, [nearest_product_index(P, i, v[i * S:(i + 1) * S]) for i in range(len(P))]


def ivfpq_search(C, P, inverted_lists, list_codes, q, k, w, perm=None):
    """IVFPQ.computeKnnIVFADC IVFPQ.java:408-450; inverted_lists[l] = iids, list_codes[l] = raw codes, insertion order"""
    nn = BoundedPriorityQueue(k)
    probes = nearest_coarse_indices(C, q, w)
    for i in range(w):
        l = probes[i]
        lut = lookup_adc(P, permute(perm, residual(C, q, l)))
        for j, iid in enumerate(inverted_lists[l]):
            code = list_codes[l][j]
            l2 = 0.0
            for mm in range(len(code)):
                l2 += lut[mm][code[mm]]
            nn.offer(iid, l2)
    return nn.results()


def nearest_centroid(codebook, desc):
    """AbstractFeatureAggregator.computeNearestCentroid AFA.java:136-155"""
    return nearest_coarse_index(codebook, desc)


def vlad(codebook, descriptors):
    """VladAggregator.aggregateInternal VladAggregator.java:56-70"""
    K, D = len(codebook), len(codebook[0])
    out = [0.0] * (K * D)
    if len(descriptors) == 0:
        return out
    for x in descriptors:
        nn = nearest_centroid(codebook, x)
        for i in range(D):
            out[nn * D + i] += x[i] - codebook[nn][i]
    return out


class JavaRandom:
    """java.util.Random: 48-bit LCG, next(bits) = (int)(seed >>> (48 - bits))"""

    def __init__(self, seed):
        self.s = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def next(self, bits):
        self.s = (self.s * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        v = self.s >> (48 - bits)
        return v - (1 << 32) if v >= (1 << 31) else v

    def next_int(self, bound):
        r = self.next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if u - r + m < (1 << 31):  # Java: loop while (u - r + m) overflows to negative
                return r
            u = self.next(31)


def random_permutation(seed, dim):
    """RandomPermutation ctor RandomPermutation.java:29-40 + Collections.shuffle(list, rnd)"""
    rnd = JavaRandom(seed)
    perm = list(range(dim))
    for i in range(dim, 1, -1):
        j = rnd.next_int(i)
        perm[i - 1], perm[j] = perm[j], perm[i - 1]
    return perm
