# builds the bench workload on the CPU (oracle encode) and caches codes/lists + probes of 2000 queries for offline studies
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
X, Q, Cq, P, prefix = bench.load_workload()
O, off, cc, ii = bench.cpu_oracle_setup(X, Cq, P, prefix)
z = np.load(prefix + "_oracle_codes.npz")
lists, codes = z["lists"], z["codes"]
# coarse probes for 2000 queries (numpy, fp64 expansion is fine for a study)
q = Q[:2000]
d2 = (q*q).sum(1)[:,None] - 2*q@Cq.T + (Cq*Cq).sum(1)[None,:]
probes = np.argsort(d2, axis=1)[:, :32]
np.savez("/tmp/study.npz", lists=lists, codes=codes, probes=probes)
print("done", codes.shape, np.bincount(lists, minlength=1024)[probes].sum(1).mean())
