"""Row f3 host logic (aggregation.py): Normalization.java and VladAggregatorMultipleVocabularies.java restated on the
host, checked against the oracle's C restatement.  CPU only: the per-vocabulary VLADs come from a stand-in aggregator
that calls the oracle, so no GPU kernel runs here (the GPU VLAD itself is covered by test_gpu_parity.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmidx_b200 as M  # noqa: E402  (loads libmmidx.so, no compute)
import pyoracle as O  # noqa: E402

REL_TOL = 1e-4  # north_star tolerance for floating point; the L2 step is required to be bit-identical


def test_l2_normalization_is_bit_identical_and_fills_zero_vectors_with_one():
    rng = np.random.default_rng(0)
    for n in (1, 7, 128, 8192, 32768):
        v = rng.normal(size=n) * rng.uniform(0.1, 100)
        assert (M.normalizeL2(v) == O.normalize_l2(v)).all()
    assert (M.normalizeL2(np.zeros(6)) == 1.0).all()  # Normalization.java:29-30 (sic)
    B = rng.normal(size=(9, 513))
    B[4] = 0
    out = M.normalizeL2(B)
    for i in range(9):
        assert (out[i] == O.normalize_l2(B[i])).all()
    assert (B[4] == 0).all()  # the input is not modified (the Java methods work in place; the mirror returns copies)


def test_power_and_ssr_normalization_within_tolerance():
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.normal(size=5000) * 10, [0.0, -0.0, 1e-300, -1e300]])
    for a in (0.5, 0.3, 1.0):
        p, q = M.normalizePower(v, a), O.normalize_power(v, a)
        assert np.allclose(p, q, rtol=REL_TOL, atol=0) and (np.sign(p) == np.sign(v)).all()
    s = M.normalizeSSR(v[:5000])
    ref = O.normalize_l2(O.normalize_power(v[:5000], 0.5))
    assert np.allclose(s, ref, rtol=REL_TOL, atol=0) and abs(np.linalg.norm(s) - 1.0) < 1e-12


class _OracleVlad:
    """stand-in for VladAggregator: same interface, VLAD from the oracle"""

    def __init__(self, codebook, device=-1):
        self.codebook = np.ascontiguousarray(codebook, dtype=np.float64)

    def getVectorLength(self):
        return self.codebook.size

    def aggregateBatch(self, images):
        offsets = np.zeros(len(images) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(im) for im in images])
        D = self.codebook.shape[1]
        desc = np.concatenate([np.asarray(im, dtype=np.float64).reshape(-1, D) for im in images])
        return O.vlad(self.codebook, desc, offsets)


def test_multiple_vocabularies_compose_like_the_reference():
    rng = np.random.default_rng(2)
    D = 16
    codebooks = [rng.normal(size=(K, D)) for K in (8, 5, 12)]
    images = [rng.normal(size=(n, D)) for n in (30, 1, 0, 77)]
    mv = M.VladAggregatorMultipleVocabularies(codebooks, aggregator=_OracleVlad)
    assert mv.getVectorLength() == (8 + 5 + 12) * D and mv.isNormalizationsOn()
    out = mv.aggregateBatch(images)
    for i, im in enumerate(images):
        subs = []
        for cb in codebooks:
            sub = _OracleVlad(cb).aggregateBatch([im])[0][0]
            subs.append(O.normalize_l2(O.normalize_power(sub, 0.5)))
        ref = O.normalize_l2(np.concatenate(subs))
        assert np.allclose(out[i], ref, rtol=REL_TOL, atol=1e-300), f"image {i}"
    # the empty image: every sub-VLAD is all zero -> filled with 1 by normalizeL2, then normalised again
    assert np.allclose(out[2], 1.0 / np.sqrt(mv.getVectorLength()))
    # one vocabulary: no second L2 pass; normalisations off: plain concatenation
    one = M.VladAggregatorMultipleVocabularies(codebooks[:1], aggregator=_OracleVlad)
    assert np.allclose(one.aggregate(images[0]), O.normalize_l2(O.normalize_power(_OracleVlad(codebooks[0]).aggregateBatch([images[0]])[0][0], 0.5)),
                       rtol=REL_TOL, atol=0)
    mv.setNormalizationsOn(False)
    raw = mv.aggregateBatch(images[:1])[0]
    assert (raw == np.concatenate([_OracleVlad(cb).aggregateBatch(images[:1])[0][0] for cb in codebooks])).all()
