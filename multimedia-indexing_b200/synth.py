"""Deterministic synthetic workloads for tests and bench.py. Not part of the search path.

Database / queries: SIFT-shaped mixture -- `ncenters` cluster centres uniform in [0,128)^d, a point is
clip(round(centre + 20*N(0,1)), 0, 255), stored as float64 but integer-valued.  Codebooks: plain Lloyd
k-means (numpy) from a seeded sample; the reference learns them offline with Weka (J/quantization/*, out of
scope) -- parity only needs both sides to consume the SAME codebooks.  Residuals use the reference's sign,
centroid - vector (IVFPQ.java:642-648, ResidualVectorComputation.java:34)."""
import numpy as np
from scipy import sparse

SEED_DB, SEED_Q, SEED_TRAIN, SEED_DESC, SEED_CENTERS = 1, 2, 3, 4, 5


def mixture_centers(d, ncenters=4096):
    return np.random.default_rng(SEED_CENTERS).uniform(0.0, 128.0, size=(ncenters, d))


def mixture(n, d, seed, centers=None, chunk=1 << 17):
    centers = mixture_centers(d) if centers is None else centers
    rng = np.random.default_rng(seed)
    out = np.empty((n, d), dtype=np.float64)
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        c = rng.integers(0, centers.shape[0], size=e - b)
        x = centers[c] + 20.0 * rng.standard_normal((e - b, d))
        out[b:e] = np.clip(np.rint(x), 0.0, 255.0)
    return out


def _assign(X, C):
    # argmin ||x - c||^2 via the expansion (training only; ties/rounding irrelevant here)
    c2 = (C * C).sum(1)[None, :]
    Ct2 = np.ascontiguousarray(-2.0 * C.T)
    out = np.empty(X.shape[0], dtype=np.int64)
    for b in range(0, X.shape[0], 4096):  # row blocks keep the [block][k] temporary cache-resident
        out[b:b + 4096] = np.argmin(X[b:b + 4096] @ Ct2 + c2, axis=1)
    return out


def kmeans(X, k, iters=20, seed=0):
    rng = np.random.default_rng(seed)
    C = X[rng.choice(X.shape[0], size=k, replace=X.shape[0] < k)].copy()
    if X.shape[0] < k:
        C += rng.standard_normal(C.shape) * 1e-3
    for _ in range(iters):
        a = _assign(X, C)
        cnt = np.bincount(a, minlength=k)
        onehot = sparse.csr_matrix((np.ones(X.shape[0]), (a, np.arange(X.shape[0]))), shape=(k, X.shape[0]))
        S = onehot @ X
        nz = cnt > 0
        C[nz] = S[nz] / cnt[nz, None]
        if (~nz).any():  # re-seed empty clusters
            C[~nz] = X[rng.choice(X.shape[0], size=int((~nz).sum()))]
    return C


def train_ivfpq(d, m, ks, nlist, ntrain=100_000, iters=20, centers=None):
    """(coarse[nlist][d], P[m][ks][d/m]) from the SEED_TRAIN sample; PQ learnt on residuals centroid - x."""
    T = mixture(ntrain, d, SEED_TRAIN, centers)
    Cq = kmeans(T, nlist, iters, seed=11)
    R = Cq[_assign(T, Cq)] - T
    return Cq, train_pq_on(R, m, ks, iters)


def train_pq_on(R, m, ks, iters=20):
    d = R.shape[1]
    S = d // m
    return np.stack([kmeans(np.ascontiguousarray(R[:, j * S:(j + 1) * S]), ks, iters, seed=100 + j) for j in range(m)])


def train_pq(d, m, ks, ntrain=100_000, iters=20, centers=None):
    return train_pq_on(mixture(ntrain, d, SEED_TRAIN, centers), m, ks, iters)


def descriptors(n_img, D=64, mean=1000, sd=200, seed=SEED_DESC, nmax=2000):
    """unit-L2-norm D-dim descriptors (SURF descriptors are L2-normalised); n_i ~ clip(round(N(mean, sd)), 1, nmax)"""
    rng = np.random.default_rng(seed)
    counts = np.clip(np.rint(rng.normal(mean, sd, size=n_img)), 1, nmax).astype(np.int64)
    offsets = np.zeros(n_img + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    X = rng.standard_normal((int(offsets[-1]), D))
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    return X, offsets


def csr_from_assignments(lists, codes, nlist):
    """Group (list id, code) pairs by list in insertion order: (list_off[nlist+1], codes_csr, iids_csr)."""
    order = np.argsort(lists, kind="stable")
    off = np.zeros(nlist + 1, dtype=np.int64)
    np.cumsum(np.bincount(lists, minlength=nlist), out=off[1:])
    return off, np.ascontiguousarray(codes[order]), order.astype(np.int32)
