import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'oracle')
import numpy as np, mmidx_b200 as M, pyoracle as O
from multimedia_indexing_b200 import synth
d, nlist, w, nq = 128, 1024, 32, 10000
ce = synth.mixture_centers(d)
Q = synth.mixture(nq, d, 2, ce)
Cq, P = synth.train_ivfpq(d, 8, 256, nlist, ntrain=20000, iters=3, centers=ce)
ix = M.IVFPQ(d, 10, 8, 256, M.TransformationType.None_, nlist)
ix.loadCoarseQuantizer(Cq); ix.loadProductQuantizer(P); ix.setW(w)
for rep in range(3):
    pr = ix.computeNearestCoarseIndices(Q)
    ref = O.coarse_topw(Cq, Q, w)
    print('probes equal', (pr == ref).all(), int((pr != ref).any(1).sum()))
