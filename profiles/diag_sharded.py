#!/usr/bin/env python
"""diagnostic (2 GPUs, torchrun): list-sharded vs replica results step by step, config-4-like geometry"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmidx_b200 as M
from multimedia_indexing_b200 import synth
from multimedia_indexing_b200.sharded import MultiIVFPQ, balanced_shard_map
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
d, m, ks = 128, 16, 256
nlist, w, n, nq, k = [int(x) for x in (sys.argv[1:6] + [8192, 64, 400000, 4000, 100][len(sys.argv) - 1:])]
timings = os.environ.get("DIAG_TIMINGS", "0") == "1"
ce = synth.mixture_centers(d, 512)
X, Q = synth.mixture(n, d, 1, ce), synth.mixture(nq, d, 2, ce)
Cq = synth.kmeans(X[:40000], nlist, 2, seed=1)
P = synth.train_pq_on(Cq[synth._assign(X[:20000], Cq)] - X[:20000], m, ks, iters=2)
def mk(S):
    mi = MultiIVFPQ(d, n, m, ks, M.TransformationType.None_, nlist, list_shards=S)
    mi.loadCoarseQuantizer(Cq); mi.loadProductQuantizer(P); mi.setW(w)
    return mi
m1 = mk(1)
lists, codes = m1.indexAll(X)
m1.connect(nq, k)
ms = mk(world)
ms.setShardMap(balanced_shard_map(np.bincount(lists, minlength=nlist), world))
ms.index.indexPQCodes(None, lists, codes)
ms.connect(nq, k)
per = (nq + world - 1) // world
dq1 = torch.from_numpy(np.ascontiguousarray(Q[rank * per:(rank + 1) * per])).cuda()
dQ = torch.from_numpy(Q).cuda()
if timings:
    m1.index.enableTimings(True); ms.index.enableTimings(True)
for step in range(8):
    a = m1.search(k, dq1, gather_all=True); torch.cuda.synchronize()
    ai, ad = a[0][:nq].cpu().numpy(), a[1][:nq].cpu().numpy()
    b = ms.search(k, dQ, gather_all=True); torch.cuda.synchronize()
    bi, bd = b[0][:nq].cpu().numpy(), b[1][:nq].cpu().numpy()
    bad = int(((ai != bi).any(1) | (ad != bd).any(1)).sum())
    if rank == 0:
        print(f"step {step}: mismatching queries {bad} / {nq}", flush=True)
    dist.barrier()
dist.barrier()
