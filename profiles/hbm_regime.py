#!/usr/bin/env python
"""HBM-streaming regime of the fused ADC scan kernel (SURVEY 8d caveat): a flat PQ index whose code
database does not fit the 126 MB L2 (N = 128 Mi codes x 8 B = 1 GiB), searched with ONE query per call so that every code byte
is read from HBM exactly once per call (with several concurrent queries the L2 serves the re-reads and DRAM traffic drops
below the algorithmic bytes).  Prints one JSON line; under ncu (profiles/run_r2_hbm.sh) the same launch gives
dram__bytes_read/write for roofline.traffic.  Codes are random bytes (the scan's cost does not depend on their values)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmidx_b200 as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 27)
ap.add_argument("--nq", type=int, default=1)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
lib, check = M._capi.lib, M._capi.check
d, m, ks, k = 128, 8, 256, 100
rng = np.random.default_rng(0)
P = rng.normal(0, 20, size=(m, ks, d // m))
pq = M.PQ(d, args.n, m, ks)
pq.loadProductQuantizer(P)
t0 = time.time()
for b in range(0, args.n, 1 << 24):
    nb = min(1 << 24, args.n - b)
    codes = rng.integers(0, 256, size=(nb, m), dtype=np.uint8)
    check(lib.mmidx_add_codes(pq._h, nb, None, C.c_void_p(codes.ctypes.data)))
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # capturable (the legacy default stream is not)
Q = torch.from_numpy(rng.normal(64, 20, size=(args.nq, d))).to(dev)
ii = torch.empty((args.nq, k), dtype=torch.int32, device=dev)
dd = torch.empty((args.nq, k), dtype=torch.float64, device=dev)
cc = torch.empty(args.nq, dtype=torch.int32, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def call():
    check(lib.mmidx_search_dev(pq._h, args.nq, p(Q), k, p(ii), p(dd), p(cc), st))


call()  # seals (pseudo lists + bank-conflict-aware order), builds the tables
torch.cuda.synchronize()
build_s = time.time() - t0
pq.enableTimings(True)
ts, scan = [], []
for _ in range(args.reps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    call()
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
    scan.append(pq.lastTimings()["scan_ms"])
alg = args.nq * args.n * m
peak = 6552.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ms, sms = float(np.median(ts)), float(np.median(scan))
print(json.dumps({"workload": f"flat PQ N={args.n} m=8 ks=256 top-100, {args.nq} query/call, L2 flushed", "codes_bytes": args.n * m,
                  "call_ms": ms, "scan_kernel_ms": sms, "algorithmic_bytes_per_call": alg,
                  "scan_GBps": alg / (sms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (sms * 1e-3) / 1e9 / peak, "peak_GBps": peak,
                  "launches": pq.lastLaunches(), "build_s": build_s}), flush=True)
