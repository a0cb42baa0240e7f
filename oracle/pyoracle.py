"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see mmidx_oracle.h).
Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


if not os.path.exists(_LIB):
    build()
lib = C.CDLL(_LIB)
_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
for name, args, res in [
    ("orc_linear_search", [_vp, _i64, _i, _vp, _i, _vp, _vp], _i),
    ("orc_pq_lut", [_vp, _i, _i, _i, _vp, _vp], None),
    ("orc_pq_encode", [_vp, _i, _i, _i, _vp, _vp, _vp], None),
    ("orc_pq_search", [_vp, _i, _i, _i, _vp, _vp, _i64, _vp, _i, _vp, _vp], _i),
    ("orc_coarse_nearest", [_vp, _i, _i, _vp], _i),
    ("orc_coarse_topw", [_vp, _i, _i, _vp, _i, _vp], None),
    ("orc_ivfpq_encode", [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp], _i),
    ("orc_ivfpq_search", [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp], _i),
    ("orc_vlad", [_vp, _i, _i, _vp, _i64, _vp, _vp], None),
    ("orc_normalize_l2", [_vp, _i64], None),
    ("orc_normalize_power", [_vp, _i64, C.c_double], None),
    ("orc_random_permutation", [_i, _i, _vp], None),
    ("orc_ivfpq_search_batch", [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _i], None),
    ("orc_pq_search_batch", [_vp, _i, _i, _i, _vp, _vp, _i64, _vp, _i64, _i, _vp, _vp, _vp, _i], None),
    ("orc_linear_search_batch", [_vp, _i64, _i, _vp, _i64, _i, _vp, _vp, _vp, _i], None),
    ("orc_ivfpq_encode_batch", [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _i64, _vp, _vp, _i], None),
    ("orc_pq_encode_batch", [_vp, _i, _i, _i, _vp, _vp, _i64, _vp, _i], None),
    ("orc_vlad_batch", [_vp, _i, _i, _vp, _vp, _i64, _vp, _vp, _i], None),
    ("orc_num_threads", [], _i),
    ("orc_bpq_new", [_i], _vp),
    ("orc_bpq_free", [_vp], None),
    ("orc_bpq_offer", [_vp, _i, C.c_double], _i),
    ("orc_bpq_size", [_vp], _i),
    ("orc_bpq_last_distance", [_vp], C.c_double),
    ("orc_bpq_to_arrays", [_vp, _vp, _vp], None),
]:
    f = getattr(lib, name)
    f.argtypes, f.restype = args, res


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads():
    return lib.orc_num_threads()


class BPQ:
    """BoundedPriorityQueue<Result> (LingPipe 4.0.1) restatement."""

    def __init__(self, k):
        self.q = lib.orc_bpq_new(k)
        if not self.q:
            raise ValueError("BoundedPriorityQueue needs a positive max size")

    def offer(self, id, dist):
        return bool(lib.orc_bpq_offer(self.q, id, dist))

    def __len__(self):
        return lib.orc_bpq_size(self.q)

    def to_arrays(self):
        n = len(self)
        ids, d = np.empty(n, np.int32), np.empty(n, np.float64)
        lib.orc_bpq_to_arrays(self.q, _p(ids), _p(d))
        return ids, d

    def __del__(self):
        lib.orc_bpq_free(self.q)


def _out(nq, k):
    return np.empty((nq, k), np.int32), np.empty((nq, k), np.float64), np.empty(nq, np.int32)


def linear_search(X, Q, k, threads=1):
    X, Q = _f64(X), _f64(Q)
    ids, dist, cnt = _out(Q.shape[0], k)
    lib.orc_linear_search_batch(_p(X), X.shape[0], X.shape[1], _p(Q), Q.shape[0], k, _p(ids), _p(dist), _p(cnt), threads)
    return ids, dist, cnt


def pq_lut(P, v):
    P, v = _f64(P), _f64(v)
    m, ks, S = P.shape
    lut = np.empty((m, ks), np.float64)
    lib.orc_pq_lut(_p(P), m, ks, S, _p(v), _p(lut))
    return lut


def pq_encode(P, X, perm=None, threads=1):
    P, X = _f64(P), _f64(X)
    m, ks, S = P.shape
    perm = None if perm is None else np.ascontiguousarray(perm, np.int32)
    codes = np.empty((X.shape[0], m), np.int32)
    lib.orc_pq_encode_batch(_p(P), m, ks, S, _p(perm), _p(X), X.shape[0], _p(codes), threads)
    return codes


def _codes(codes, ks):
    return np.ascontiguousarray(codes, np.uint8 if ks <= 256 else np.uint16)


def pq_search(P, codes, Q, k, perm=None, threads=1):
    P, Q = _f64(P), _f64(Q)
    m, ks, S = P.shape
    codes = _codes(codes, ks)
    perm = None if perm is None else np.ascontiguousarray(perm, np.int32)
    ids, dist, cnt = _out(Q.shape[0], k)
    lib.orc_pq_search_batch(_p(P), m, ks, S, _p(perm), _p(codes), codes.shape[0], _p(Q), Q.shape[0], k, _p(ids), _p(dist),
                            _p(cnt), threads)
    return ids, dist, cnt


def coarse_topw(Cq, Q, w):
    Cq, Q = _f64(Cq), _f64(Q)
    out = np.empty((Q.shape[0], w), np.int32)
    for i in range(Q.shape[0]):
        lib.orc_coarse_topw(_p(Cq), Cq.shape[0], Cq.shape[1], _p(Q[i]), w, _p(out[i]))
    return out


def ivfpq_encode(Cq, P, X, perm=None, threads=1):
    Cq, P, X = _f64(Cq), _f64(P), _f64(X)
    m, ks, S = P.shape
    perm = None if perm is None else np.ascontiguousarray(perm, np.int32)
    lists = np.empty(X.shape[0], np.int32)
    codes = np.empty((X.shape[0], m), np.int32)
    lib.orc_ivfpq_encode_batch(_p(Cq), Cq.shape[0], Cq.shape[1], _p(P), m, ks, _p(perm), _p(X), X.shape[0], _p(lists),
                               _p(codes), threads)
    return lists, codes


def ivfpq_search(Cq, P, list_off, codes, iids, Q, k, w, perm=None, threads=1):
    Cq, P, Q = _f64(Cq), _f64(P), _f64(Q)
    m, ks, S = P.shape
    codes = _codes(codes, ks)
    list_off = np.ascontiguousarray(list_off, np.int64)
    iids = np.ascontiguousarray(iids, np.int32)
    perm = None if perm is None else np.ascontiguousarray(perm, np.int32)
    ids, dist, cnt = _out(Q.shape[0], k)
    lib.orc_ivfpq_search_batch(_p(Cq), Cq.shape[0], Cq.shape[1], _p(P), m, ks, _p(perm), _p(list_off), _p(codes), _p(iids),
                               _p(Q), Q.shape[0], k, w, _p(ids), _p(dist), _p(cnt), threads)
    return ids, dist, cnt


def vlad(codebook, desc, offsets, threads=1):
    codebook, desc = _f64(codebook), _f64(desc)
    offsets = np.ascontiguousarray(offsets, np.int64)
    K, D = codebook.shape
    n_img = offsets.shape[0] - 1
    out = np.empty((n_img, K * D), np.float64)
    assign = np.empty(desc.shape[0], np.int32)
    lib.orc_vlad_batch(_p(codebook), K, D, _p(desc), _p(offsets), n_img, _p(out), _p(assign), threads)
    return out, assign


def random_permutation(seed, dim):
    out = np.empty(dim, np.int32)
    lib.orc_random_permutation(seed, dim, _p(out))
    return out


def normalize_l2(v):
    v = _f64(v).copy()
    lib.orc_normalize_l2(_p(v), v.size)
    return v


def normalize_power(v, a):
    v = _f64(v).copy()
    lib.orc_normalize_power(_p(v), v.size, a)
    return v


_rot_keepalive = None


def set_rotation(R):
    """install (or clear with None) the RandomRotation matrix the encode / search functions apply (PQ.java:237-241)"""
    global _rot_keepalive
    if R is None:
        _rot_keepalive = None
        lib.orc_set_rotation(None, 0)
    else:
        _rot_keepalive = _f64(R)
        lib.orc_set_rotation(_p(_rot_keepalive), _rot_keepalive.shape[0])


def apply_rotation(R, v):
    R, v = _f64(R), _f64(v)
    out = np.empty_like(v)
    lib.orc_apply_rotation(_p(R), R.shape[0], _p(v), _p(out))
    return out


def pca_project(Vt, means, X, l2=False):
    """PCA.sampleToEigenSpace for every row of X"""
    Vt, means, X = _f64(Vt), _f64(means), _f64(X)
    nc, ss = Vt.shape
    out = np.empty((X.shape[0], nc), np.float64)
    for i in range(X.shape[0]):
        lib.orc_pca_project(_p(Vt), _p(means), nc, ss, _p(X[i]), 1 if l2 else 0, _p(out[i]))
    return out


def vlad_multi(codebooks, desc, offsets, normalize=True, threads=1):
    """VladAggregatorMultipleVocabularies.aggregate VAMV.java:84-101 over a batch of images"""
    subs = []
    for cb in codebooks:
        sub, _ = vlad(cb, desc, offsets, threads=threads)
        if normalize:
            sub = np.stack([normalize_l2(normalize_power(r, 0.5)) for r in sub])
        subs.append(sub)
    multi = np.concatenate(subs, axis=1)
    if normalize and len(codebooks) > 1:
        multi = np.stack([normalize_l2(r) for r in multi])
    return multi


def set_faithful_costs(on):
    """BASELINE.md variant (A): ivfpq_search pays the reference's per-candidate allocations (timing runs only)"""
    lib.orc_set_faithful_costs(1 if on else 0)
