"""Row f3, CPU side: the ORACLE's restatement of Normalization.java and VladAggregatorMultipleVocabularies.java against
hand-derived known answers and plain numpy formulas (the device path is compared with this oracle in
test_gpu_parity.py::test_vlad_multi_vocabulary_and_normalizations)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as O  # noqa: E402


def test_l2_known_answers():
    assert (O.normalize_l2(np.array([3.0, 4.0])) == np.array([3.0 / 5.0, 4.0 / 5.0])).all()
    assert (O.normalize_l2(np.zeros(6)) == 1.0).all()  # Normalization.java:29-30: a zero vector is FILLED WITH ONES (sic)
    rng = np.random.default_rng(0)
    for n in (1, 7, 128, 8192):
        v = rng.normal(size=n) * rng.uniform(0.1, 100)
        acc = 0.0
        for x in v.tolist():  # squares added for i ascending, each rounded (Normalization.java:23-25)
            acc += x * x
        assert (O.normalize_l2(v) == v / np.sqrt(acc)).all()


def test_power_known_answers():
    v = np.array([4.0, -9.0, 0.0, -0.0, 2.25])
    p = O.normalize_power(v, 0.5)  # signum(x) * pow(|x|, a), Normalization.java:74-79
    assert (p == np.array([2.0, -3.0, 0.0, 0.0, 1.5])).all()
    assert np.allclose(O.normalize_power(np.array([8.0, -27.0]), 1.0 / 3.0), [2.0, -3.0], rtol=1e-15)
    assert (O.normalize_power(v, 1.0) == v).all()


def test_multiple_vocabularies_compose_like_the_reference():
    rng = np.random.default_rng(2)
    D = 16
    codebooks = [rng.normal(size=(K, D)) for K in (8, 5, 12)]
    images = [rng.normal(size=(n, D)) for n in (30, 1, 0, 77)]
    offsets = np.zeros(len(images) + 1, np.int64)
    offsets[1:] = np.cumsum([len(im) for im in images])
    desc = np.concatenate(images)
    out = O.vlad_multi(codebooks, desc, offsets, normalize=True)
    L = (8 + 5 + 12) * D
    assert out.shape == (4, L)
    for i, im in enumerate(images):
        subs = []
        for cb in codebooks:
            sub = O.vlad(cb, im.reshape(-1, D), np.array([0, len(im)], np.int64))[0][0]
            sub = np.sign(sub) * np.sqrt(np.abs(sub))
            n2 = np.sqrt((sub * sub).sum())
            subs.append(sub / n2 if n2 > 0 else np.ones_like(sub))
        cat = np.concatenate(subs)
        assert np.allclose(out[i], cat / np.sqrt((cat * cat).sum()), rtol=1e-12, atol=1e-300), f"image {i}"
    # the empty image: every sub-VLAD is all zero -> filled with 1 by normalizeL2, then normalised again
    assert np.allclose(out[2], 1.0 / np.sqrt(L))
    # one vocabulary: no second L2 pass (VAMV.java:96-98); normalisations off: plain concatenation
    one = O.vlad_multi(codebooks[:1], desc, offsets, normalize=True)
    assert abs(np.linalg.norm(one[0]) - 1.0) < 1e-12
    raw = O.vlad_multi(codebooks, desc, offsets, normalize=False)
    assert (raw == np.concatenate([O.vlad(cb, desc, offsets)[0] for cb in codebooks], axis=1)).all()
