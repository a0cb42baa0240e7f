#!/usr/bin/env python
"""Measured coverage of the SURVEY.md section-8 rows that are NOT the headline bench line (bench.py): for every row the
B200 path through the C ABI with host buffers (wall clock, copies included), the CPU oracle on a bounded sample with
all host threads, and a bit-exact parity check of the two on that sample.  One JSON line per row.

  python profiles/bench_rows.py            (on a GPU box; ~2 minutes)

Rows: a8 Linear (configs[0]), a6 PQ flat scan (configs[1]), a2+a9+a10 IVFPQ encode, a9 PQ encode, a11+a12 VLAD
(configs[4], scaled to 2000 images), a1 coarse probes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmidx_b200 as M  # noqa: E402
from multimedia_indexing_b200 import synth  # noqa: E402
import pyoracle as O  # noqa: E402

T = O.num_threads()


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        out = fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def emit(row, what, unit, gpu_units, gpu_s, cpu_units, cpu_s, parity, extra=None):
    rec = {"row": row, "what": what, "unit": unit, "gpu": gpu_units / gpu_s, "gpu_s": gpu_s,
           "cpu": cpu_units / cpu_s, "cpu_cores": T, "cpu_sample": cpu_units, "speedup": (gpu_units / gpu_s) / (cpu_units / cpu_s),
           "bit_identical_on_sample": bool(parity)}
    if extra:
        rec.update(extra)
    print(json.dumps(rec), flush=True)


def same(res, ref):
    return all((np.asarray(a) == np.asarray(b)).all() for a, b in zip(res[:3], ref))


def main():
    # ---- a8 Linear: configs[0] (10k x 64, k = 10) ----
    d = 64
    ce = synth.mixture_centers(d, 256)
    X, Q = synth.mixture(10_000, d, synth.SEED_DB, ce), synth.mixture(10_000, d, synth.SEED_Q, ce)
    lin = M.Linear(d, 10_000)
    lin.indexVectors(None, X)
    gs, res = timed(lambda: lin.searchBatch(10, Q))
    t0 = time.perf_counter()
    ref = O.linear_search(X, Q[:2000], 10, threads=T)
    cs = time.perf_counter() - t0
    emit("a8", "Linear exact top-10, 10k x 64 (configs[0]), 10 000 queries/call", "queries/s", len(Q), gs, 2000, cs,
         same([r[:2000] for r in res[:3]], ref))

    # ---- shared 1M x 128 database and quantizers (same generator as bench.py, fewer k-means iterations) ----
    d, m, ks, nlist = 128, 8, 256, 1024
    ce = synth.mixture_centers(d)
    X = synth.mixture(1_000_000, d, synth.SEED_DB, ce)
    Q = synth.mixture(2_000, d, synth.SEED_Q, ce)
    Cq, Pr = synth.train_ivfpq(d, m, ks, nlist, ntrain=30_000, iters=5, centers=ce)
    Pf = synth.train_pq(d, m, ks, ntrain=30_000, iters=5, centers=ce)

    # ---- a9 PQ encode + a6 PQ flat ADC scan: configs[1] ----
    pq = M.PQ(d, 1_000_000, m, ks)
    pq.loadProductQuantizer(Pf)
    pq.encode(X[:100_000])  # warm-up (context, scratch pool)
    t0 = time.perf_counter()
    _, codes = pq.indexVectors(None, X, return_codes=True)
    gs = time.perf_counter() - t0
    t0 = time.perf_counter()
    oc = O.pq_encode(Pf, X[:100_000], threads=T)
    cs = time.perf_counter() - t0
    emit("a9", "PQ encode 1M x 128, m=8 ks=256 (host vectors in, codes out)", "vectors/s", len(X), gs, 100_000, cs,
         (codes[:100_000] == oc).all())
    gs, res = timed(lambda: pq.searchBatch(100, Q), reps=2)
    t0 = time.perf_counter()
    ref = O.pq_search(Pf, codes, Q[:64], 100, threads=T)
    cs = time.perf_counter() - t0
    emit("a6", "PQ flat ADC scan top-100 over 1M codes (configs[1]), 2000 queries/call", "queries/s", len(Q), gs, 64, cs,
         same([r[:64] for r in res[:3]], ref), {"algorithmic_GBps": len(Q) * 8e6 / gs / 1e9})
    del pq

    # ---- a2 + a3 + a9 + a10 IVFPQ indexing (coarse assign + residual + PQ encode) ----
    ix = M.IVFPQ(d, 1_000_000, m, ks, M.TransformationType.None_, nlist)
    ix.loadCoarseQuantizer(Cq)
    ix.loadProductQuantizer(Pr)
    ix.setW(32)
    ix.encode(X[:100_000])  # warm-up
    t0 = time.perf_counter()
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    gs = time.perf_counter() - t0
    t0 = time.perf_counter()
    ol, oc = O.ivfpq_encode(Cq, Pr, X[:100_000], threads=T)
    cs = time.perf_counter() - t0
    emit("a2+a9+a10", "IVFPQ indexVector: coarse assign + residual + PQ encode, 1M x 128, nlist=1024", "vectors/s", len(X), gs,
         100_000, cs, (lists[:100_000] == ol).all() and (codes[:100_000] == oc).all())
    # ---- a1 coarse probes ----
    gs, pr = timed(lambda: ix.computeNearestCoarseIndices(Q))
    t0 = time.perf_counter()
    opr = O.coarse_topw(Cq, Q[:500], 32)
    cs = time.perf_counter() - t0
    emit("a1", "coarse top-32 of 1024 centroids (fp32 filter + exact verification), 2000 queries/call", "queries/s", len(Q), gs,
         500, cs, (pr[:500] == opr).all(), {"cpu_cores": 1})
    del ix

    # ---- a11 + a12 VLAD: configs[4] geometry, 2000 images x ~1000 descriptors x 64, K = 128 ----
    desc, offsets = synth.descriptors(2000)
    cb = synth.kmeans(desc[:50_000], 128, iters=5, seed=7)
    agg = M.VladAggregator(cb)
    gs, (vl, asg) = timed(lambda: agg.aggregateBatch((desc, offsets), return_assign=True), reps=2)
    t0 = time.perf_counter()
    ov, oa = O.vlad(cb, desc[:offsets[200]], offsets[:201], threads=T)
    cs = time.perf_counter() - t0
    emit("a11+a12", "VLAD aggregate (K=128, D=64): 2000 images, ~1000 descriptors each", "descriptors/s", len(desc), gs,
         int(offsets[200]), cs, (vl[:200] == ov).all() and (asg[:offsets[200]] == oa).all(), {"images_per_s": 2000 / gs})


if __name__ == "__main__":
    main()
