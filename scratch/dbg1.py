import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'oracle')
import numpy as np, mmidx_b200 as M, pyoracle as O
from multimedia_indexing_b200 import synth
d, m, ks, nlist, w, n, nq, k = 128, 8, 256, 64, 16, 20000, 700, 100
ce = synth.mixture_centers(d, 64)
X, Q = synth.mixture(n, d, 1, ce), synth.mixture(nq, d, 2, ce)
Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=5000, iters=3, centers=ce)
ix = M.IVFPQ(d, n, m, ks, M.TransformationType.None_, nlist)
ix.loadCoarseQuantizer(Cq); ix.loadProductQuantizer(P); ix.setW(w)
lists, codes = ix.indexVectors(None, X, return_codes=True)
iids, dist, cnt, _ = ix.searchBatch(k, Q)
off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w, threads=16)
print('equal', (iids == oi).all(), (dist == od).all())
