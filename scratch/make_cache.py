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
# inputs of scratch/sim_reorder.c (run it from /tmp/sim): codes in CSR order, list offsets, probe frequency per list
os.makedirs("/tmp/sim", exist_ok=True)
order = np.argsort(lists, kind="stable")
off = np.zeros(1025, np.int64)
off[1:] = np.cumsum(np.bincount(lists, minlength=1024))
codes.astype(np.uint8)[order].tofile("/tmp/sim/codes.bin")
off.tofile("/tmp/sim/off.bin")
np.bincount(probes.ravel(), minlength=1024).astype(np.int64).tofile("/tmp/sim/freq.bin")
print("done", codes.shape, np.bincount(lists, minlength=1024)[probes].sum(1).mean())
