"""Row f4, CPU side: the reference's PCA text file format + whitening fold-in (host logic of dimreduction.py) and the
ORACLE's restatement of PCA.sampleToEigenSpace against numpy.  The device projection is compared with this oracle in
test_gpu_parity.py::test_pca_projection_vs_oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmidx_b200 as M  # noqa: E402,F401  (registers the package; loads libmmidx.so, no compute)
import pyoracle as O  # noqa: E402
from multimedia_indexing_b200.dimreduction import PCA  # noqa: E402


def _basis(rng, d, nc):
    A = rng.normal(size=(400, d)) * np.linspace(3, 0.2, d)
    mean = A.mean(0)
    _, s, Vt = np.linalg.svd(A - mean, full_matrices=False)
    return mean, (s ** 2 / (len(A) - 1))[:nc + 3], Vt[:nc + 3]


def test_oracle_projection_known_answers():
    rng = np.random.default_rng(0)
    d, nc = 24, 6
    mean, eig, Vt = _basis(rng, d, nc)
    X = rng.normal(size=(50, d)) * np.linspace(3, 0.2, d)
    Y = O.pca_project(Vt[:nc], mean, X)
    assert np.allclose(Y, (X - mean) @ Vt[:nc].T, rtol=1e-10, atol=1e-12)
    seq = []
    for i in range(nc):  # total += V_t[i][j] * sample[j], j ascending (explicit loop: builtin sum() is compensated)
        t = 0.0
        for pj in ((X[3] - mean) * Vt[i]).tolist():
            t += pj
        seq.append(t)
    assert (Y[3] == np.array(seq)).all()
    Z = O.pca_project(Vt[:nc] / np.sqrt(eig[:nc])[:, None], mean, X, l2=True)
    assert np.allclose(np.linalg.norm(Z, axis=1), 1.0)


def test_file_round_trip_and_whitening_fold_in(tmp_path):
    rng = np.random.default_rng(0)
    d, nc = 24, 6
    mean, eig, Vt = _basis(rng, d, nc)
    plain = PCA(nc, 400, d)
    plain.loadPCAFromFile((mean, eig, Vt))
    assert plain.V_t.shape == (nc, d) and (plain.V_t == Vt[:nc]).all() and (plain.means == mean).all()
    # text file: line 1 means, line 2 eigenvalues, then one eigenvector per line (PCA.java:219-247)
    f = str(tmp_path / "pca.txt")
    plain.savePCAToFile(f, eig)
    lines = open(f).read().splitlines()
    assert len(lines) == 2 + nc and len(lines[0].split(" ")) == d and len(lines[2].split(" ")) == d
    again = PCA(nc, 400, d)
    again.loadPCAFromFile(f)
    assert (again.V_t == plain.V_t).all() and (again.means == plain.means).all()  # repr() round-trips doubles exactly
    # whitening: rows scaled by eigenvalue^-0.5 at load time (PCA.java:283-310)
    white = PCA(nc, 400, d, doWhitening=True)
    white.loadPCAFromFile(f)
    assert np.allclose(white.V_t, Vt[:nc] / np.sqrt(eig[:nc])[:, None], rtol=1e-15)
    fewer = PCA(3, 400, d)  # fewer components than the file holds: the leading rows are used
    fewer.loadPCAFromFile(f)
    assert (fewer.V_t == Vt[:3]).all()


def test_error_behaviour():
    rng = np.random.default_rng(1)
    mean, eig, Vt = _basis(rng, 10, 4)
    p = PCA(4, 400, 10)
    with pytest.raises(RuntimeError):  # "PCA is not correctly initiallized!"
        p.sampleToEigenSpace(np.zeros(10))
    p.loadPCAFromFile((mean, eig, Vt))
    with pytest.raises(ValueError):  # "Unexpected vector length!"
        p.sampleToEigenSpace(np.zeros(9))
    with pytest.raises(ValueError):  # "Means line is wrong!"
        PCA(4, 400, 11).loadPCAFromFile((mean, eig, Vt))
    with pytest.raises(ValueError):  # not enough components
        PCA(20, 400, 10).loadPCAFromFile((mean, eig, Vt))
