"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs and against the committed golden vectors.  Bar (BASELINE.json north_star):
bit-exact PQ codes / list ids and neighbour id sets, distances within 1e-4 relative.  This implementation is
held to the stricter bar of bit-identical distances and identical result ORDER, because it reproduces the
reference's binary64 operation order."""
import os

import numpy as np
import pytest

import mmidx_b200 as M
import pyoracle as O
from multimedia_indexing_b200 import _capi, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-4  # north_star tolerance for distances (we assert equality first; this is the stated bound)


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def assert_same(res, ref, what=""):
    iids, dist, cnt = res[:3]
    oi, od, oc = ref
    assert (cnt == oc).all(), f"{what}: result counts differ"
    fin = np.isfinite(od)
    assert np.allclose(dist[fin], od[fin], rtol=REL_TOL, atol=0), f"{what}: distances beyond 1e-4 relative"
    for r in range(len(oc)):
        assert set(iids[r, :oc[r]]) == set(oi[r, :oc[r]]), f"{what}: id set differs for query {r}"
    assert (dist == od).all(), f"{what}: distances not bit-identical"
    assert (iids == oi).all(), f"{what}: result order differs"


def make_ivfpq(d, m, ks, nlist, w, Cq, P, perm=None, n=10 ** 6):
    ix = M.IVFPQ(d, n, m, ks, M.TransformationType.None_, nlist)
    ix.loadCoarseQuantizer(Cq)
    ix.loadProductQuantizer(P)
    if perm is not None:
        ix.setPermutation(perm)
    ix.setW(w)
    return ix


# ---------------------------------------------------------------- golden fixtures
def test_linear_golden():
    g = load("linear")
    ix = M.Linear(g["X"].shape[1], 1000)
    ix.indexVectors(None, g["X"])
    k = int(g["k"])
    assert_same(ix.searchBatch(k, g["Q"]), (g["ids"], g["dist"], np.full(len(g["Q"]), k, np.int32)), "linear golden")
    assert (ix.getVector(17) == g["X"][17]).all()
    a = ix.computeNearestNeighbors(k, "3")  # query by id -> getVector + search (Linear.java:181-184)
    b = ix.computeNearestNeighbors(k, g["X"][3])
    assert a.getIds() == b.getIds() and (a.getDistances() == b.getDistances()).all()


@pytest.mark.parametrize("name", ["ivfpq_a", "ivfpq_perm"])
def test_ivfpq_and_pq_golden(name):
    g = load(name)
    perm = g["perm"] if g["perm"].size else None
    Cq, P, X, Q, k, w = g["Cq"], g["P"], g["X"], g["Q"], int(g["k"]), int(g["w"])
    m, ks, S = P.shape
    ix = make_ivfpq(X.shape[1], m, ks, Cq.shape[0], w, Cq, P, perm)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    assert (lists == g["lists"]).all() and (codes == g["codes"]).all()
    assert (ix.computeNearestCoarseIndices(Q) == g["probes"]).all()
    assert_same(ix.searchBatch(k, Q), (g["ids"], g["dist"], g["cnt"]), name)
    assert (ix.computeLookupADC(Q[:1])[0] == g["lut0"]).all()
    pq = M.PQ(X.shape[1], 1000, m, ks)
    pq.loadProductQuantizer(P)
    if perm is not None:
        pq.setPermutation(perm)
    _, pc = pq.indexVectors(None, X, return_codes=True)
    assert (pc == g["pq_codes"]).all()
    assert_same(pq.searchBatch(k, Q), (g["pq_ids"], g["pq_dist"], np.full(len(Q), k, np.int32)), name + " flat PQ")


def test_vlad_golden():
    g = load("vlad")
    agg = M.VladAggregator(g["codebook"])
    out, assign = agg.aggregateBatch((g["desc"], g["offsets"]), return_assign=True)
    assert (out == g["out"]).all()
    assert (agg.aggregate(np.zeros((0, 6))) == 0).all()
    _, oa = O.vlad(g["codebook"], g["desc"], g["offsets"])
    assert (assign == oa).all()


# ---------------------------------------------------------------- seeded comparisons against the oracle
CASES = [
    # d, m, ks, nlist, w, n, nq, k, perm
    (32, 4, 64, 32, 8, 6000, 40, 10, False),
    (128, 8, 256, 64, 16, 20000, 64, 100, False),   # north-star geometry, scaled down
    (128, 16, 256, 32, 8, 8000, 32, 100, True),     # config-4 geometry (m=16) + RandomPermutation
    (24, 3, 300, 8, 8, 3000, 16, 20, False),        # ks > 256 -> short codes, w == nlist
    (20, 5, 7, 5, 2, 500, 9, 600, False),           # k > candidates, odd m*ks, k > 512
    (64, 8, 256, 16, 4, 5000, 3, 1, False),         # k = 1
]


@pytest.mark.parametrize("mode", ["fast", "exact"])
@pytest.mark.parametrize("case", CASES)
def test_ivfpq_vs_oracle(case, mode, monkeypatch):
    # "fast": fused fp32-filter + exact-verify kernel where the geometry allows it (ks <= 256, m in {8, 16});
    # "exact": MMIDX_MODE=exact forces the binary64 ADC-table kernels for every geometry.  Same bits either way.
    monkeypatch.setenv("MMIDX_MODE", mode)
    d, m, ks, nlist, w, n, nq, k, use_perm = case
    ce = synth.mixture_centers(d, 64)
    X, Q = synth.mixture(n, d, synth.SEED_DB, ce), synth.mixture(nq, d, synth.SEED_Q, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=min(n, 5000), iters=4, centers=ce)
    perm = M.random_permutation(1, d) if use_perm else None
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P, perm)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    ol, oc = O.ivfpq_encode(Cq, P, X, perm, threads=8)
    assert (lists == ol).all(), "coarse list ids differ"
    assert (codes == oc).all(), "PQ codes differ"
    assert (ix.listSizes() == np.bincount(ol, minlength=nlist)).all()
    assert (ix.computeNearestCoarseIndices(Q) == O.coarse_topw(Cq, Q, w)).all()
    off, cc, ii = synth.csr_from_assignments(ol, oc, nlist)
    assert_same(ix.searchBatch(k, Q), O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w, perm, threads=8), "ivfpq")
    # encode-only entry point gives the same codes without storing
    l2, c2 = ix.encode(X[:100])
    assert (l2 == ol[:100]).all() and (c2 == oc[:100]).all() and ix.getLoadCounter() == n


@pytest.mark.parametrize("mode", ["fast", "exact"])
@pytest.mark.parametrize("d,m,ks,n,nq,k,use_perm", [
    (32, 4, 64, 7000, 20, 10, False),
    (128, 8, 256, 30000, 16, 100, False),    # configs[1] geometry: the fused fp32-filter kernel over pseudo lists
    (128, 8, 256, 40000, 700, 100, True),    # ... one CTA per query, three pseudo lists, RandomPermutation
    (64, 16, 256, 20000, 40, 256, False),    # m = 16, k at the fast path's limit
    (16, 2, 1000, 2000, 8, 5, False)])
def test_pq_vs_oracle(d, m, ks, n, nq, k, use_perm, mode, monkeypatch):
    # "fast": a flat index with ks = 256, m in {8, 16} is searched as an IVFPQ with one zero centroid and the negated
    # codebook (same bits); "exact": the binary64 table kernel for every geometry
    monkeypatch.setenv("MMIDX_MODE", mode)
    ce = synth.mixture_centers(d, 64)
    X, Q = synth.mixture(n, d, synth.SEED_DB, ce), synth.mixture(nq, d, synth.SEED_Q, ce)
    P = synth.train_pq(d, m, ks, ntrain=min(n, 5000), iters=4, centers=ce)
    perm = M.random_permutation(3, d) if use_perm else None
    pq = M.PQ(d, n, m, ks)
    pq.loadProductQuantizer(P)
    if perm is not None:
        pq.setPermutation(perm)
    _, codes = pq.indexVectors(None, X[:n // 2], return_codes=True)
    oc = O.pq_encode(P, X, perm, threads=8)
    assert (codes == oc[:n // 2]).all()
    ref_half = O.pq_search(P, oc[:n // 2], Q, k, perm, threads=8)
    assert_same(pq.searchBatch(k, Q), ref_half, "pq, first half")
    pq.indexVectors(None, X[n // 2:])  # an add after a search re-seals the pseudo lists
    assert_same(pq.searchBatch(k, Q), O.pq_search(P, oc, Q, k, perm, threads=8), "pq")
    if perm is not None:
        return
    luts = pq.computeLookupADC(Q[:3])
    for i in range(3):
        assert (luts[i] == O.pq_lut(P, Q[i])).all()


@pytest.mark.parametrize("n,d,nq,k", [(10000, 64, 50, 10), (777, 5, 7, 1000), (4097, 128, 3, 100)])
def test_linear_vs_oracle(n, d, nq, k):
    ce = synth.mixture_centers(d, 64)
    X, Q = synth.mixture(n, d, synth.SEED_DB, ce), synth.mixture(nq, d, synth.SEED_Q, ce)
    ix = M.Linear(d, n)
    ix.indexVectors(None, X[: n // 2])
    ix.indexVectors(None, X[n // 2:])  # two appends: exercises the packed-block growth
    assert_same(ix.searchBatch(k, Q), O.linear_search(X, Q, k, threads=8), "linear")


def test_ties_at_the_kth_boundary_follow_the_queue_rules():
    """Exact duplicates -> many exact binary64 ties, including at the k-th boundary."""
    rng = np.random.default_rng(3)
    d, m, ks, nlist, w, k = 16, 4, 16, 6, 6, 25
    base = np.clip(np.rint(rng.normal(64, 20, size=(40, d))), 0, 255)
    X = base[rng.integers(0, 40, size=3000)]  # every vector repeated ~75 times
    Q = base[:10] + 1.0
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=2000, iters=3, centers=base)
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    assert_same(ix.searchBatch(k, Q), O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w), "ivfpq ties")
    pq = M.PQ(d, 5000, m, ks)
    pq.loadProductQuantizer(P)
    _, pc = pq.indexVectors(None, X, return_codes=True)
    assert_same(pq.searchBatch(k, Q), O.pq_search(P, pc, Q, k), "pq ties")
    lin = M.Linear(d, 5000)
    lin.indexVectors(None, X)
    assert_same(lin.searchBatch(k, Q), O.linear_search(X, Q, k), "linear ties")
    # coarse ties: duplicated coarse centroids
    Cd = np.vstack([Cq[:3], Cq[:3]])
    ix2 = make_ivfpq(d, m, ks, 6, 3, Cd, P)
    assert (ix2.computeNearestCoarseIndices(Q) == O.coarse_topw(Cd, Q, 3)).all()


@pytest.mark.parametrize("mode", ["fast", "exact"])
def test_coarse_probes_vs_oracle(mode, monkeypatch):
    """computeNearestCoarseIndices (IVFPQ.java:575-601): fp32 filter + exact verification == the exact kernels == oracle,
    including duplicated centroids (exact ties at and across the w-th boundary) and a tie class larger than the
    verification collector (the kernel's exact sweep + ordered tie pass)."""
    if mode == "exact":
        monkeypatch.setenv("MMIDX_MODE", "exact")
    rng = np.random.default_rng(11)
    d, m, ks = 24, 4, 16
    P = rng.normal(0, 10, size=(m, ks, d // m))
    # (a) generic real-valued centroids and queries, w from 1 to nlist
    Cq = rng.normal(60, 25, size=(300, d))
    Q = rng.normal(60, 25, size=(40, d))
    for w in (1, 7, 32, 300):
        ix = make_ivfpq(d, m, ks, 300, w, Cq, P)
        assert (ix.computeNearestCoarseIndices(Q) == O.coarse_topw(Cq, Q, w)).all(), f"w={w}"
    # (b) every centroid repeated 5 times: ties everywhere, cut at the boundary for w not a multiple of 5
    Cd = np.repeat(np.rint(rng.normal(60, 25, size=(60, d))), 5, axis=0)[rng.permutation(300)]
    Qi = np.rint(rng.normal(60, 25, size=(25, d)))
    for w in (3, 16, 64):
        ix = make_ivfpq(d, m, ks, 300, w, Cd, P)
        assert (ix.computeNearestCoarseIndices(Qi) == O.coarse_topw(Cd, Qi, w)).all(), f"dup w={w}"
    # (c) 1500 identical centroids among 1600: the tie class exceeds the 1024-entry collector
    Cb = np.vstack([np.tile(Cq[:1], (1500, 1)), Cq[1:101]])[rng.permutation(1600)]
    ix = make_ivfpq(d, m, ks, 1600, 40, Cb, P)
    Qb = np.vstack([Cq[:1] + 0.5, Q[:7]])
    assert (ix.computeNearestCoarseIndices(Qb) == O.coarse_topw(Cb, Qb, 40)).all(), "wide tie class"
    # (d) the register-key instantiations of the verification (16 and 32 keys per thread), ragged last block of keys,
    #     duplicated centroids among them
    for nl, w in ((2048, 64), (5000, 64), (8192, 33)):
        Cl = rng.normal(60, 25, size=(nl, d))
        Cl[nl // 2:nl // 2 + 40] = Cl[:40]
        ix = make_ivfpq(d, m, ks, nl, w, Cl, P)
        Ql = np.vstack([Cl[:3] + 0.25, Q[:13]])
        assert (ix.computeNearestCoarseIndices(Ql) == O.coarse_topw(Cl, Ql, w)).all(), f"nlist={nl}"


def test_bulk_reload_matches_incremental_index():
    """indexPQCode / loadIndexInMemory path (IVFPQ.java:357-386, 680-728)."""
    d, m, ks, nlist, w, k = 32, 4, 64, 16, 4, 10
    ce = synth.mixture_centers(d, 32)
    X, Q = synth.mixture(3000, d, 1, ce), synth.mixture(12, d, 2, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=3000, iters=3, centers=ce)
    a = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    lists, codes = a.indexVectors(None, X, return_codes=True)
    b = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    b.indexPQCodes(None, lists[:1000], codes[:1000])
    for i in range(1000, 1010):  # the reference's single-item form with Java signed bytes
        assert b.indexPQCode(str(i), int(lists[i]), (codes[i].astype(np.int16) - 128).astype(np.int8))
    assert not b.indexPQCode("1005", 0, np.zeros(m, np.int8))  # duplicate id -> False (ASS.java:237-240)
    b.indexPQCodes(None, lists[1010:], codes[1010:])
    ra, rb = a.searchBatch(k, Q), b.searchBatch(k, Q)
    assert (ra[0] == rb[0]).all() and (ra[1] == rb[1]).all()
    # through the reference's "ivfadc" tuple format (persistence.py: what a Java host would write to / scan from BDB)
    from multimedia_indexing_b200 import persistence
    c = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    assert persistence.load_ivfpq(c, persistence.ivfpq_records(lists, codes, ks), batch=700) == len(X)
    rc = c.searchBatch(k, Q)
    assert (ra[0] == rc[0]).all() and (ra[1] == rc[1]).all() and (c.listSizes() == a.listSizes()).all()
    # search, add more, search again: re-seal keeps insertion order
    a.indexVectors(None, X[:500] + 1.0)
    X2 = np.vstack([X, X[:500] + 1.0])
    ol, oc = O.ivfpq_encode(Cq, P, X2, threads=8)
    off, cc, ii = synth.csr_from_assignments(ol, oc, nlist)
    assert_same(a.searchBatch(k, Q), O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w), "after second add")


def test_reference_api_surface_and_errors():
    d, m, ks, nlist = 16, 4, 16, 20
    ce = synth.mixture_centers(d, 16)
    X = synth.mixture(50, d, 1, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=500, iters=2, centers=ce)
    ix = M.IVFPQ(d, 30, m, ks, M.TransformationType.None_, nlist)
    with pytest.raises(M.MmidxError) as e:  # quantizers not loaded
        ix.indexVector("a", X[0])
    assert e.value.code == _capi.ERR_STATE
    ix.loadCoarseQuantizer(Cq)
    ix.loadProductQuantizer(P)
    assert ix.w == 2  # default w = (int)(nlist * 0.1), IVFPQ.java:188
    assert ix.indexVector("img0", X[0]) is True
    assert ix.indexVector("img0", X[1]) is False  # duplicate id
    with pytest.raises(M.MmidxError) as e:
        ix.indexVector("bad", X[0][:5])
    assert "dimensionality of the vector is wrong" in str(e.value)
    for i in range(1, 30):
        assert ix.indexVector(f"img{i}", X[i])
    assert ix.indexVector("img30", X[30]) is False  # full (ASS.java:232-235)
    assert ix.getLoadCounter() == 30 and ix.getInternalId("img7") == 7 and ix.getId(7) == "img7"
    ans = ix.computeNearestNeighbors(5, X[3])
    assert len(ans.getIds()) <= 5 and all(s.startswith("img") for s in ans.getIds())
    assert (np.diff(ans.getDistances()) >= 0).all()
    ix.setW(0)
    with pytest.raises(M.MmidxError) as e:  # BoundedPriorityQueue ctor throws for size 0
        ix.computeNearestNeighbors(5, X[3])
    assert e.value.code == _capi.ERR_W
    ix.setW(nlist + 1)
    with pytest.raises(M.MmidxError) as e:  # NPE in the reference (IVFPQ.java:598)
        ix.computeNearestNeighbors(5, X[3])
    assert e.value.code == _capi.ERR_W
    ix.setW(nlist)
    with pytest.raises(M.MmidxError):
        ix.computeNearestNeighbors(0, X[3])
    # w = nlist with 30 vectors: fewer candidates than k
    ans = ix.computeNearestNeighbors(100, X[3])
    assert len(ans.getIds()) == 30
    ix.close()
    # empty index
    e2 = M.PQ(d, 10, m, ks)
    e2.loadProductQuantizer(P)
    iids, dist, cnt, _ = e2.searchBatch(3, X[:2])
    assert (cnt == 0).all() and (iids == -1).all() and np.isinf(dist).all()


def test_vlad_vs_oracle_ragged():
    K, D = 128, 64
    desc, offsets = synth.descriptors(40, D, mean=300, sd=150)
    offsets = np.concatenate([offsets[:10], [offsets[9]], offsets[10:]])  # insert an empty image
    cb = synth.kmeans(desc[:5000], K, 5, seed=1)
    agg = M.VladAggregator(cb)
    out, assign = agg.aggregateBatch((desc, offsets), return_assign=True)
    oo, oa = O.vlad(cb, desc, offsets, threads=8)
    assert (assign == oa).all(), "centroid assignment differs"
    assert (out == oo).all(), "VLAD vectors not bit-identical"
    assert (out[9] == 0).all()
    one = agg.aggregate(desc[offsets[3]:offsets[4]])
    assert (one == oo[3]).all()
    with pytest.raises(M.MmidxError):
        agg.aggregate(np.zeros((3, D + 1)))


# ---------------------------------------------------------------- full-size properties (BASELINE configs)
def test_full_size_config3_properties():
    """IVFPQ 1M x 128, nlist=1024, w=32, m=8, top-100: oracle agreement on a query sample + size-independent
    properties (sortedness, ids inside probed lists, idempotence, sharding-by-probe additivity)."""
    d, m, ks, nlist, w, k, n, nq = 128, 8, 256, 1024, 32, 100, 1_000_000, 2000
    ce = synth.mixture_centers(d)
    X, Q = synth.mixture(n, d, synth.SEED_DB, ce), synth.mixture(nq, d, synth.SEED_Q, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=50_000, iters=8, centers=ce)
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P, n=n)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    assert ix.getLoadCounter() == n and (ix.listSizes() == np.bincount(lists, minlength=nlist)).all()
    sel = np.random.default_rng(0).choice(n, 3000, replace=False)
    ol, oc = O.ivfpq_encode(Cq, P, X[sel], threads=O.num_threads())
    assert (lists[sel] == ol).all() and (codes[sel] == oc).all()
    iids, dist, cnt, _ = ix.searchBatch(k, Q)
    assert (cnt == k).all() and (np.diff(dist, axis=1) >= 0).all()
    probes = ix.computeNearestCoarseIndices(Q)
    for r in range(0, nq, 97):
        assert np.isin(lists[iids[r]], probes[r]).all()
        assert len(set(iids[r])) == k
    again = ix.searchBatch(k, Q)
    assert (again[0] == iids).all() and (again[1] == dist).all()
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    ref = O.ivfpq_search(Cq, P, off, cc, ii, Q[:200], k, w, threads=O.num_threads())
    assert_same((iids[:200], dist[:200], cnt[:200]), ref, "config 3 sample")
    # single query == row of the batch
    one = ix.searchBatch(k, Q[5:6])
    assert (one[0][0] == iids[5]).all() and (one[1][0] == dist[5]).all()


@pytest.mark.parametrize("mode", ["fast", "exact"])
def test_large_batch_single_split_path(mode, monkeypatch):
    monkeypatch.setenv("MMIDX_MODE", mode)
    _large_batch_single_split_path()


def _large_batch_single_split_path():
    """>= 592 queries in one chunk -> one CTA per query (nsplit == 1) and large select/merge grids.  Regression:
    TopK::init() lacked a barrier, so a CTA could act on a previous CTA's shared-memory garbage."""
    d, m, ks, nlist, w, n, nq, k = 64, 8, 256, 256, 16, 60000, 3000, 100
    ce = synth.mixture_centers(d, 512)
    X, Q = synth.mixture(n, d, synth.SEED_DB, ce), synth.mixture(nq, d, synth.SEED_Q, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=10000, iters=4, centers=ce)
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    for _ in range(3):
        assert (ix.computeNearestCoarseIndices(Q) == O.coarse_topw(Cq, Q, w)).all()
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    ref = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w, threads=O.num_threads())
    for _ in range(2):
        assert_same(ix.searchBatch(k, Q), ref, "nsplit=1")


def test_fast_path_overflow_falls_back_to_direct_kernel():
    """More exact duplicates than the fp32 collector can keep inside its error band (> 512 equal distances):
    the fast kernel hands the query to the table-free exact kernel and the tie pass; results must still follow the
    queue rules bit for bit."""
    rng = np.random.default_rng(11)
    d, m, ks, nlist, w, k = 64, 8, 256, 8, 8, 50
    base = np.clip(np.rint(rng.normal(64, 20, size=(5, d))), 0, 255)
    X = np.vstack([base[rng.integers(0, 5, size=6000)], np.clip(np.rint(rng.normal(64, 20, size=(500, d))), 0, 255)])
    X = X[rng.permutation(len(X))]
    Q = np.vstack([base + 1.0, np.clip(np.rint(rng.normal(64, 20, size=(11, d))), 0, 255)])
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=3000, iters=3, centers=np.vstack([base, base + 40.0]))
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    ref = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w)
    assert_same(ix.searchBatch(k, Q), ref, "overflow -> direct")
    big = np.vstack([Q] * 60)  # > 592 queries: one CTA per query
    assert_same(ix.searchBatch(k, big), tuple(np.concatenate([r] * 60) for r in ref), "overflow -> direct, nsplit=1")


def test_barrier_free_sweep_overflow_is_rolled_back(monkeypatch):
    """Adversarial offer order for the fused scan kernel: the nearest list holds only far candidates (the admission
    threshold it leaves is loose), the next lists hold thousands of candidates that all beat it, so the barrier-free
    sweep runs past the fp32 collector; its entries must be dropped and the list scanned again with round barriers.
    Lists are crafted through the bulk-reload entry point (indexPQCode, IVFPQ.java:357-386)."""
    monkeypatch.setenv("MMIDX_STATS", "1")
    rng = np.random.default_rng(5)
    d, m, ks, nlist, w, k = 64, 8, 256, 4, 4, 20
    S = d // m
    P = rng.normal(0, 5, size=(m, ks, S))
    Cq = np.zeros((nlist, d))
    Cq[:, 0] = [0.0, 6.0, 12.0, 18.0]  # a query near the origin ranks the lists 0, 1, 2, 3
    Q = rng.normal(0, 0.3, size=(6, d))
    Q = np.vstack([Q] * 120)            # > 592 queries: one CTA per query (no probe splitting)
    q0 = Q[0]

    def pick(l, nearest, count, width):
        r = (Cq[l] - q0).reshape(m, S)
        lut = ((r[:, None, :] - P) ** 2).sum(-1)                       # [m][ks]
        order = np.argsort(lut, axis=1)
        pool = order[:, :width] if nearest else order[:, -width:]
        return np.stack([pool[j, rng.integers(0, width, size=count)] for j in range(m)], axis=1).astype(np.uint8)

    codes = np.vstack([pick(0, False, 150, 16), pick(1, True, 5000, 6), pick(2, True, 3000, 6), pick(3, False, 40, 16)])
    lists = np.concatenate([np.full(150, 0), np.full(5000, 1), np.full(3000, 2), np.full(40, 3)]).astype(np.int32)
    sh = rng.permutation(len(lists))
    lists, codes = lists[sh], codes[sh]
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    ix.indexPQCodes(None, lists, codes)
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    ref = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w, threads=O.num_threads())
    assert_same(ix.searchBatch(k, Q), ref, "sweep overflow")
    st = ix.debugStats()
    assert st[1] >= len(Q) and st[3] == 0, f"the re-scan path did not run (stats {st})"
    assert np.isin(lists[ref[0][0]], [1, 2]).all()  # the scenario is what it claims: the winners sit in the later lists


def test_config4_geometry_large_batch():
    """m = 16, S = 8 (BASELINE configs[3] geometry, scaled down) with one CTA per query."""
    d, m, ks, nlist, w, n, nq, k = 128, 16, 256, 128, 32, 40000, 1200, 100
    ce = synth.mixture_centers(d, 256)
    X, Q = synth.mixture(n, d, synth.SEED_DB, ce), synth.mixture(nq, d, synth.SEED_Q, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=10000, iters=4, centers=ce)
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    ref = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w, threads=O.num_threads())
    assert_same(ix.searchBatch(k, Q), ref, "m=16 large batch")


# ---------------------------------------------------------------- fp32 range of the filters
@pytest.mark.parametrize("db_scale,q_scale", [(1e-25, 1e-25), (1e25, 1e25), (1.0, 1e20), (1.0, 1e-20), (1e-10, 1e-10)])
def test_fast_path_is_exact_outside_the_fp32_range(db_scale, q_scale):
    """Data far outside fp32's comfortable range: the fp32 filters (coarse and ADC) must not be trusted there.  An index
    outside the magnitude window takes the binary64 kernels, a query outside it is evaluated by the table-free exact
    kernel; inside the window the filter carries an absolute underflow slack.  Same bits as the oracle in every case."""
    d, m, ks, nlist, w, n, nq, k = 64, 8, 256, 32, 8, 12000, 48, 50
    ce = synth.mixture_centers(d, 64)
    X, Q = synth.mixture(n, d, synth.SEED_DB, ce) * db_scale, synth.mixture(nq, d, synth.SEED_Q, ce) * q_scale
    Q[::7] = synth.mixture(nq, d, 5, ce)[::7] * db_scale  # some queries at the database's own scale
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=4000, iters=3, centers=ce)
    Cq, P = Cq * db_scale, P * db_scale
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    ol, oc = O.ivfpq_encode(Cq, P, X, threads=8)
    assert (lists == ol).all() and (codes == oc).all()
    assert (ix.computeNearestCoarseIndices(Q) == O.coarse_topw(Cq, Q, w)).all()
    off, cc, ii = synth.csr_from_assignments(ol, oc, nlist)
    assert_same(ix.searchBatch(k, Q), O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w, threads=8), f"scale {db_scale:g}/{q_scale:g}")
    pq = M.PQ(d, n, m, ks)
    pq.loadProductQuantizer(P)
    _, pc = pq.indexVectors(None, X, return_codes=True)
    assert_same(pq.searchBatch(k, Q), O.pq_search(P, pc, Q, k, threads=8), f"flat scale {db_scale:g}/{q_scale:g}")


def test_results_do_not_depend_on_the_tie_rule_when_no_boundary_tie_exists():
    """The oracle's queue tie rule is restated from memory (LingPipe is un-vendored).  Property: a query
    whose k-th and (k+1)-th exact distances differ has ONE possible id set and distance list whatever the tie rule is --
    plain sort of the exact distances.  So for those queries the GPU result must equal a rule-free reference; the count
    of the others is what bench.py reports as ties_at_k_boundary."""
    d, m, ks, nlist, w, n, nq, k = 64, 8, 256, 32, 8, 15000, 200, 40
    ce = synth.mixture_centers(d, 64)
    X = synth.mixture(n, d, synth.SEED_DB, ce) + np.random.default_rng(0).normal(0, 0.3, (n, d))  # de-duplicated codes
    Q = synth.mixture(nq, d, synth.SEED_Q, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=5000, iters=3, centers=ce)
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    iids, dist, cnt, _ = ix.searchBatch(k, Q)
    i1, d1, c1, _ = ix.searchBatch(k + 1, Q)
    probes = ix.computeNearestCoarseIndices(Q)
    untied = 0
    for r in range(nq):
        # rule-free reference: exact ADC distance of every candidate (the library's own binary64 tables), plain sort
        cand, dd = [], []
        for l in probes[r]:
            idx = np.nonzero(lists == l)[0]
            lut = ix.computeLookupADC((Cq[l] - Q[r]).reshape(1, -1))[0]  # residual = centroid - query (IVFPQ.java:645)
            acc = np.zeros(len(idx))
            for j in range(m):  # l2distance += LUT[j][code_j], j ascending from 0.0 (IVFPQ.java:435-438)
                acc = acc + lut[j, codes[idx, j]]
            cand.append(idx)
            dd.append(acc)
        cand, dd = np.concatenate(cand), np.concatenate(dd)
        order = np.argsort(dd, kind="stable")
        if len(cand) > k and dd[order[k - 1]] == dd[order[k]]:
            continue  # a boundary tie: membership depends on the queue rule (covered by the oracle comparison)
        untied += 1
        top = order[:k]
        assert set(cand[top]) == set(iids[r, :cnt[r]]), r
        assert (np.sort(dd[top]) == dist[r, :cnt[r]]).all(), r
        assert c1[r] <= k or d1[r, k - 1] != d1[r, k]  # the k+1 search sees the same boundary
    assert untied >= nq * 0.9


def test_concurrent_searches_and_vlad_from_four_threads():
    """computeNearestNeighbors is unsynchronised in the reference (ASS.java:281) and VladAggregator.aggregate is
    called from pool threads on one shared aggregator (ImageVectorization.java:60,198): both must be re-entrant here."""
    import threading
    d, m, ks, nlist, w, n, k = 64, 8, 256, 32, 8, 20000, 30
    ce = synth.mixture_centers(d, 64)
    X = synth.mixture(n, d, synth.SEED_DB, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=4000, iters=3, centers=ce)
    ix = make_ivfpq(d, m, ks, nlist, w, Cq, P)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    desc, offs = synth.descriptors(40, 32)
    cb = synth.kmeans(desc[:4000], 16, 3, seed=1)
    agg = M.VladAggregator(cb)
    vref, _ = O.vlad(cb, desc, offs)
    sizes = [1, 33, 700, 5000]  # single query, small batch, one chunk, pipelined multi-chunk path
    Qs = [synth.mixture(s, d, 10 + t, ce) for t, s in enumerate(sizes)]
    refs = [O.ivfpq_search(Cq, P, off, cc, ii, q, k, w, threads=8) for q in Qs]
    ix.searchBatch(k, Qs[1])  # seal + tables before the threads start
    errors = []

    def worker(t):
        try:
            for rep in range(6):
                assert_same(ix.searchBatch(k, Qs[t]), refs[t], f"thread {t} rep {rep}")
                out, _ = agg.aggregateBatch((desc, offs))
                assert (out == vref.reshape(out.shape)).all(), f"vlad thread {t} rep {rep}"
        except BaseException as e:  # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    th = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


# ---------------------------------------------------------------- rows f2, f3, f4 on the device
def test_vlad_multi_vocabulary_and_normalizations():
    """VladAggregatorMultipleVocabularies (VAMV.java:84-101) and Normalization.java on the device vs the oracle: the L2 step is
    bit-identical (ordered sum of squares); power(0.5) is the correctly rounded root where Math.pow / libm pow may differ in
    the last bit, hence the 1e-12 bound on the normalised values."""
    rng = np.random.default_rng(2)
    D = 16
    codebooks = [rng.normal(size=(K, D)) for K in (8, 5, 12)]
    images = [rng.normal(size=(n, D)) for n in (30, 1, 0, 77, 400)]
    offsets = np.zeros(len(images) + 1, np.int64)
    offsets[1:] = np.cumsum([len(im) for im in images])
    desc = np.concatenate(images)
    mv = M.VladAggregatorMultipleVocabularies(codebooks)
    assert mv.getVectorLength() == (8 + 5 + 12) * D and mv.isNormalizationsOn()
    out = mv.aggregateBatch(images)
    ref = O.vlad_multi(codebooks, desc, offsets, normalize=True)
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-300)
    assert np.allclose(out[2], 1.0 / np.sqrt(mv.getVectorLength()))  # empty image: zero sub-VLADs are filled with ones (sic)
    assert np.allclose(mv.aggregate(images[0]), ref[0], rtol=1e-12, atol=0)
    mv.setNormalizationsOn(False)
    assert (mv.aggregateBatch(images) == O.vlad_multi(codebooks, desc, offsets, normalize=False)).all()  # plain concatenation
    # Normalization.java stand-alone, incl. a row longer than one staging chunk and the zero row
    for n in (1, 7, 4096, 4097, 32768):
        v = rng.normal(size=n) * rng.uniform(0.1, 100)
        assert (M.normalizeL2(v) == O.normalize_l2(v)).all(), n
    B = rng.normal(size=(9, 513))
    B[4] = 0
    nb = M.normalizeL2(B)
    for i in range(9):
        assert (nb[i] == O.normalize_l2(B[i])).all()
    assert (B[4] == 0).all()  # the mirror returns copies
    v = np.concatenate([rng.normal(size=5000) * 10, [0.0, -0.0, 1e-300, -1e300]])
    for a in (0.5, 0.3, 1.0):
        p, q = M.normalizePower(v, a), O.normalize_power(v, a)
        assert np.allclose(p, q, rtol=1e-12, atol=0) and (np.sign(p) == np.sign(v)).all()
    s = M.normalizeSSR(v[:5000])
    assert np.allclose(s, O.normalize_l2(O.normalize_power(v[:5000], 0.5)), rtol=1e-12, atol=0)


def test_pca_projection_vs_oracle():
    """PCA.sampleToEigenSpace (PCA.java:188-208): the device product adds a row's products for j ascending like the oracle, so
    the projection is bit-identical to it; 1e-4 relative is the bound claimed against the Java path (EJML un-vendored)."""
    from multimedia_indexing_b200.dimreduction import PCA
    rng = np.random.default_rng(0)
    for d, nc, n in ((24, 6, 50), (1000, 100, 333), (64, 64, 1)):
        A = rng.normal(size=(max(400, 2 * d), d)) * np.linspace(3, 0.2, d)
        mean = A.mean(0)
        _, sv, Vt = np.linalg.svd(A - mean, full_matrices=False)
        eig = sv ** 2 / (len(A) - 1)
        X = rng.normal(size=(n, d)) * np.linspace(3, 0.2, d)
        plain = PCA(nc, len(A), d)
        plain.loadPCAFromFile((mean, eig, Vt))
        Y = plain.sampleToEigenSpaceBatch(X)
        ref = O.pca_project(plain.V_t, mean, X)
        assert np.allclose(Y, ref, rtol=REL_TOL, atol=1e-12) and (Y == ref).all()
        assert (plain.sampleToEigenSpace(X[0]) == Y[0]).all()
        white = PCA(nc, len(A), d, doWhitening=True)
        white.loadPCAFromFile((mean, eig, Vt))
        Z = white.sampleToEigenSpaceBatch(X)
        zr = O.pca_project(white.V_t, mean, X, l2=True)
        assert (Z == zr).all() and np.allclose(np.linalg.norm(Z, axis=1), 1.0)


@pytest.mark.parametrize("kind", ["ivfpq", "pq"])
def test_random_rotation_with_a_supplied_matrix(kind):
    """TransformationType.RandomRotation (PQ.java:237-241,294-298; IVFPQ.java:319-323,420-424): the vector / the residual is
    multiplied by the d x d matrix (RandomRotation.java:44-49) before product quantization, at index and at search time.
    The matrix is supplied (EJML's generator is un-vendored); codes, ids and distances equal the oracle's, bit for bit."""
    d, m, ks, nlist, w, n, nq, k = 32, 8, 256, 16, 5, 6000, 40, 20
    rng = np.random.default_rng(5)
    R, _ = np.linalg.qr(rng.normal(size=(d, d)))
    ce = synth.mixture_centers(d, 64)
    X, Q = synth.mixture(n, d, synth.SEED_DB, ce), synth.mixture(nq, d, synth.SEED_Q, ce)
    v = rng.normal(size=d)
    assert np.allclose(O.apply_rotation(R, v), v @ R, rtol=1e-12)
    try:
        O.set_rotation(R)
        if kind == "ivfpq":
            Cq = synth.kmeans(X[:3000], nlist, 3, seed=1)
            res = np.stack([O.apply_rotation(R, Cq[l] - x) for l, x in zip(synth._assign(X[:3000], Cq), X[:3000])])
            P = synth.train_pq_on(res, m, ks, iters=3)
            with pytest.raises(M.MmidxError):
                M.IVFPQ(d, n, m, ks, M.TransformationType.RandomRotation, nlist)  # no matrix: not reproducible
            ix = M.IVFPQ(d, n, m, ks, M.TransformationType.RandomRotation, nlist, rotation=R)
            ix.loadCoarseQuantizer(Cq)
            ix.loadProductQuantizer(P)
            ix.setW(w)
            lists, codes = ix.indexVectors(None, X, return_codes=True)
            ol, oc = O.ivfpq_encode(Cq, P, X, threads=1)
            assert (lists == ol).all() and (codes == oc).all()
            off, cc, ii = synth.csr_from_assignments(ol, oc, nlist)
            assert_same(ix.searchBatch(k, Q), O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w, threads=1), "rotated ivfpq")
        else:
            P = synth.train_pq_on(np.stack([O.apply_rotation(R, x) for x in X[:3000]]), m, ks, iters=3)
            pq = M.PQ(d, n, m, ks, M.TransformationType.RandomRotation, rotation=R)
            pq.loadProductQuantizer(P)
            _, codes = pq.indexVectors(None, X, return_codes=True)
            oc = O.pq_encode(P, X, threads=1)
            assert (codes == oc).all()
            assert_same(pq.searchBatch(k, Q), O.pq_search(P, oc, Q, k, threads=1), "rotated pq")
    finally:
        O.set_rotation(None)


def test_filtered_argmins_resolve_ties_like_the_reference():
    """argmin_filter.cuh: nearest product centroid (PQ.java:411-429), nearest coarse centroid (IVFPQ.java:547-564) and nearest
    VLAD centroid (AFA.java:136-155) are decided by an fp32 filter when the runner-up is provably farther and in binary64
    otherwise.  Duplicated centroids and vectors that sit exactly on centroids / midway between two force the binary64 path:
    the lowest index among equal minima must win, as the reference's strict `<` does."""
    rng = np.random.default_rng(11)
    d, m, ks, nlist, n = 64, 8, 256, 40, 5000
    S = d // m
    ce = synth.mixture_centers(d, 32)
    X = synth.mixture(n, d, synth.SEED_DB, ce)
    Cq = synth.kmeans(X[:2000], nlist, 2, seed=3)
    Cq[7] = Cq[3]          # duplicated coarse centroids: list 3 must win over list 7
    Cq[39] = Cq[0]
    P = synth.train_pq_on(Cq[synth._assign(X[:2000], Cq)] - X[:2000], m, ks, iters=2)
    P[:, 200] = P[:, 5]    # duplicated product centroids in every sub-quantizer
    P[2, 17] = P[2, 16]
    X[:50] = Cq[rng.integers(0, nlist, 50)]                       # vectors ON coarse centroids (residual 0)
    X[50:100] = 0.5 * (Cq[rng.integers(0, nlist, 50)] + Cq[rng.integers(0, nlist, 50)])  # midway between two
    X[100:150] = Cq[3] - P[:, 5].reshape(-1)                      # residual exactly the duplicated product centroid
    ix = make_ivfpq(d, m, ks, nlist, 8, Cq, P)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    ol, oc = O.ivfpq_encode(Cq, P, X, threads=8)
    assert (lists == ol).all() and (codes == oc).all()
    assert not (lists == 7).any() and not (codes == 200).any()    # the later duplicate never wins
    pq = M.PQ(d, n, m, ks)
    pq.loadProductQuantizer(P)
    Y = X.copy()
    Y[:200] = np.tile(P[:, rng.integers(0, ks, 200)].transpose(1, 0, 2).reshape(200, d), 1)  # vectors made of product centroids
    _, pc = pq.indexVectors(None, Y, return_codes=True)
    assert (pc == O.pq_encode(P, Y, threads=8)).all()
    # VLAD assignment with duplicated codebook rows and descriptors on / between centroids
    K, D = 64, 32
    desc, offs = synth.descriptors(30, D)
    cb = synth.kmeans(desc[:5000], K, 2, seed=2)
    cb[40] = cb[9]
    desc[:64] = cb
    desc[64:100] = 0.5 * (cb[:36] + cb[1:37])
    agg = M.VladAggregator(cb)
    out, assign = agg.aggregateBatch((desc, offs), return_assign=True)
    ov, oa = O.vlad(cb, desc, offs, threads=8)
    assert (assign == oa).all() and (out == ov.reshape(out.shape)).all()
    assert not (assign == 40).any()
