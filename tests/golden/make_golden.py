"""Generates tests/golden/*.npz with tests/pyref.py (the pure-Python restatement of the Java path).
The reference cannot be imported or run here (Java, no JVM), so these are NOT reference outputs: they pin
the C oracle and the CUDA path to an independently written restatement.  Run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pyref  # noqa: E402


def dataset(rng, n, d, dup=0):
    X = np.clip(np.rint(rng.normal(64, 30, size=(n, d))), 0, 255)
    if dup:  # exact duplicates -> exact binary64 ties, exercising the BoundedPriorityQueue tie rules
        X[rng.integers(0, n, size=dup)] = X[rng.integers(0, n, size=dup)]
    return X


def main():
    rng = np.random.default_rng(20261017)
    # ---- Linear ----
    X = dataset(rng, 300, 8, dup=60)
    Q = np.vstack([dataset(rng, 6, 8), X[:3]])
    k = 7
    ids, dist = zip(*[pyref.linear_search(X.tolist(), q.tolist(), k) for q in Q])
    np.savez(os.path.join(HERE, "linear.npz"), X=X, Q=Q, k=k, ids=np.array(ids, np.int32), dist=np.array(dist))

    # ---- PQ / IVFPQ ----
    for name, d, m, ks, nlist, w, use_perm in [("ivfpq_a", 16, 4, 16, 12, 4, False), ("ivfpq_perm", 12, 3, 300, 6, 6, True)]:
        n, nq, k = 400, 8, 9
        X = dataset(rng, n, d, dup=80)
        Q = np.vstack([dataset(rng, nq - 2, d), X[5:7]])
        S = d // m
        Cq = X[rng.choice(n, nlist, replace=False)] + rng.normal(0, 1, size=(nlist, d))
        P = rng.normal(0, 25, size=(m, ks, S))
        perm = pyref.random_permutation(1, d) if use_perm else None
        enc = [pyref.ivfpq_encode(Cq.tolist(), P.tolist(), x.tolist(), perm) for x in X]
        lists = np.array([e[0] for e in enc], np.int32)
        codes = np.array([e[1] for e in enc], np.int32)
        inv = [[] for _ in range(nlist)]
        lc = [[] for _ in range(nlist)]
        for iid, (l, c) in enumerate(enc):
            inv[l].append(iid)
            lc[l].append(c)
        res = [pyref.ivfpq_search(Cq.tolist(), P.tolist(), inv, lc, q.tolist(), k, w, perm) for q in Q]
        cnt = np.array([len(r[0]) for r in res], np.int32)
        ids = np.full((len(Q), k), -1, np.int32)
        dist = np.full((len(Q), k), np.inf)
        for i, r in enumerate(res):
            ids[i, :cnt[i]] = r[0]
            dist[i, :cnt[i]] = r[1]
        probes = np.array([pyref.nearest_coarse_indices(Cq.tolist(), q.tolist(), w) for q in Q], np.int32)
        # flat PQ over the raw vectors with the same product quantizer
        pcodes = np.array([pyref.pq_encode(P.tolist(), x.tolist(), perm) for x in X], np.int32)
        pres = [pyref.pq_search(P.tolist(), pcodes.tolist(), q.tolist(), k, perm) for q in Q]
        lut0 = np.array(pyref.lookup_adc(P.tolist(), Q[0].tolist()))
        np.savez(os.path.join(HERE, name + ".npz"), X=X, Q=Q, k=k, w=w, Cq=Cq, P=P,
                 perm=np.array(perm if perm is not None else [], np.int32), lists=lists, codes=codes, ids=ids, dist=dist,
                 cnt=cnt, probes=probes, pq_codes=pcodes, pq_ids=np.array([r[0] for r in pres], np.int32),
                 pq_dist=np.array([r[1] for r in pres]), lut0=lut0)

    # ---- VLAD ----
    K, D = 5, 6
    cb = rng.normal(0, 1, size=(K, D))
    imgs = [rng.normal(0, 1, size=(n, D)) for n in (0, 1, 17, 40)]
    imgs[3][5] = imgs[3][4]
    out = np.array([pyref.vlad(cb.tolist(), im.tolist()) for im in imgs])
    offsets = np.cumsum([0] + [len(im) for im in imgs]).astype(np.int64)
    np.savez(os.path.join(HERE, "vlad.npz"), codebook=cb, desc=np.concatenate(imgs), offsets=offsets, out=out)

    # ---- RandomPermutation (java.util.Random known answers: new Random(1).nextInt(1000) == 985) ----
    np.savez(os.path.join(HERE, "perm.npz"), p10=np.array(pyref.random_permutation(1, 10), np.int32),
             p128=np.array(pyref.random_permutation(1, 128), np.int32),
             p1024_seed7=np.array(pyref.random_permutation(7, 1024), np.int32))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
