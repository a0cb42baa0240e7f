#!/usr/bin/env python
"""bench.py -- queries/sec + recall@100 of IVFPQ search (BASELINE.json): 1M x 128 synthetic SIFT-shaped
vectors, m=8, ks=256, nlist=1024, nprobe (w)=32, top-100.

  python bench.py --gpus N --steps K --warmup W            the B200 path (libmmidx.so, through the C ABI)
  python bench.py --impl reference --gpus N ...            the CPU arm: the oracle (C restatement of the Java
                                                            path; the reference itself is Java and there is no JVM
                                                            here) on all host cores, bounded query sample per step

A step = one pass of the search path over one batch of NQ queries.  `value` = device-resident queries/s
(CUDA events on the launching stream), `e2e` = the same through the host C-ABI call mmidx_search with pinned
HOST buffers (H2D of the queries and D2H of ids+distances inside the timed region).  N > 1: S list shards x R query
groups (multimedia-indexing_b200/sharded.py): the lists are sharded only until one shard's codes fit half of the L2
(S = 1 for this 12 MB index: every rank searches nq/N queries over the whole index and the results are all-gathered
over NCCL; MMIDX_LIST_SHARDS forces list sharding with the per-shard top-k exchange and device merge).  Total work is
fixed, so scaling is "strong"."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, M_SUB, KS, NLIST, W_PROBE, TOPK = 128, 8, 256, 1024, 32, 100
N_DB, NQ = 1_000_000, 10_000
NTRAIN, KM_ITERS = 50_000, 10
N_GT = 1000  # queries with exact ground truth for recall@100
WORKLOAD = "IVFPQ 1Mx128 nlist=1024 nprobe=32 m=8 ks=256 top-100 (BASELINE.json configs[2])"
CACHE = os.environ.get("MMIDX_BENCH_CACHE", "/tmp/mmidx_bench_cache")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_workload(need_db=True):
    """Deterministic inputs. Codebooks are cached on the box so both arms use the same bytes."""
    import mmidx_b200  # noqa: F401  registers the package (loads libmmidx.so; no compute)
    from multimedia_indexing_b200 import synth

    t0 = time.time()
    ce = synth.mixture_centers(D)
    key = f"d{D}_m{M_SUB}_ks{KS}_nl{NLIST}_nt{NTRAIN}_it{KM_ITERS}"
    os.makedirs(CACHE, exist_ok=True)
    fq = os.path.join(CACHE, key + "_quantizers.npz")
    if os.path.exists(fq):
        z = np.load(fq)
        Cq, P = z["Cq"], z["P"]
    else:
        Cq, P = synth.train_ivfpq(D, M_SUB, KS, NLIST, ntrain=NTRAIN, iters=KM_ITERS, centers=ce)
        tmp = fq + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, Cq=Cq, P=P)
        os.replace(tmp, fq)
    X = synth.mixture(N_DB, D, synth.SEED_DB, ce) if need_db else None
    Q = synth.mixture(NQ, D, synth.SEED_Q, ce)
    log(f"[bench] workload ready in {time.time() - t0:.1f}s")
    return X, Q, Cq, P, os.path.join(CACHE, key)


def recall_at_k(ids, gt):
    return float(np.mean([len(set(ids[r]) & set(gt[r])) / gt.shape[1] for r in range(gt.shape[0])]))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_setup(X, Cq, P, prefix, lists=None, codes=None):
    """CSR lists for the oracle. Encodes with the oracle itself unless assignments are given."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    from multimedia_indexing_b200 import synth

    if lists is None:
        f = prefix + "_oracle_codes.npz"
        if os.path.exists(f):
            z = np.load(f)
            lists, codes = z["lists"], z["codes"]
        else:
            t0 = time.time()
            lists, codes = O.ivfpq_encode(Cq, P, X, threads=O.num_threads())
            codes = codes.astype(np.uint8)
            log(f"[bench] oracle encoded {len(X)} vectors in {time.time() - t0:.1f}s on {O.num_threads()} threads")
            tmp = f + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, lists=lists, codes=codes)
            os.replace(tmp, f)
    off, cc, ii = synth.csr_from_assignments(lists, np.asarray(codes, dtype=np.uint8), NLIST)
    return O, off, cc, ii


def exact_gt_cpu(X, Qs, k):
    """exact top-k by squared L2 on the CPU. Inputs are integers <= 255 and d = 128, so every partial sum is an
    integer < 2^24: float32 matmul is exact here."""
    Xf = X.astype(np.float32)
    Qf = Qs.astype(np.float32)
    x2 = (Xf * Xf).sum(1)
    best_d = np.full((len(Qs), k), np.inf, np.float32)
    best_i = np.zeros((len(Qs), k), np.int64)
    for b in range(0, len(X), 131072):
        dd = x2[None, b:b + 131072] - 2.0 * (Qf @ Xf[b:b + 131072].T)
        idx = np.argpartition(dd, k - 1, axis=1)[:, :k]
        cd = np.concatenate([best_d, np.take_along_axis(dd, idx, 1)], 1)
        ci = np.concatenate([best_i, idx + b], 1)
        o = np.argsort(cd, axis=1, kind="stable")[:, :k]
        best_d, best_i = np.take_along_axis(cd, o, 1), np.take_along_axis(ci, o, 1)
    return best_i


def run_reference(args, rank):
    if rank != 0:
        return
    X, Q, Cq, P, prefix = load_workload()
    O, off, cc, ii = cpu_oracle_setup(X, Cq, P, prefix)
    cores = O.num_threads()
    # bounded sample per step: ~2 s of wall time on all cores
    t0 = time.perf_counter()
    O.ivfpq_search(Cq, P, off, cc, ii, Q[: 16 * cores], TOPK, W_PROBE, threads=cores)
    rate = 16 * cores / (time.perf_counter() - t0)
    sample = int(min(NQ, max(16 * cores, rate * 2.0)))
    for _ in range(args.warmup):
        O.ivfpq_search(Cq, P, off, cc, ii, Q[:sample], TOPK, W_PROBE, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids, dist, cnt = O.ivfpq_search(Cq, P, off, cc, ii, Q[:sample], TOPK, W_PROBE, threads=cores)
    dt = time.perf_counter() - t0
    qps = sample * args.steps / dt
    ngt = min(sample, 200)
    rec = recall_at_k(ids[:ngt], exact_gt_cpu(X, Q[:ngt], TOPK))
    desc = f"first {sample} of the {NQ} queries per step, query-level threads over a shared read-only index"
    print(json.dumps({
        "impl": "reference", "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nq_per_step": sample, "k": TOPK, "threads": cores},
        "recall_at_100": rec,
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": desc,
                         "note": "C restatement of the Java path (oracle/); the Java reference cannot run here (no JVM)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_gpu(args, rank, world, local_rank):
    import torch
    import mmidx_b200 as M
    from multimedia_indexing_b200 import _capi

    lib = _capi.lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    X, Q, Cq, P, prefix = load_workload()

    # ---- build the index (untimed): GPU coarse-assign + residual + PQ encode of the whole database ----
    t0 = time.time()
    if world > 1:
        from multimedia_indexing_b200.sharded import HybridIVFPQ
        # G = S list shards x R query groups (sharded.py).  Rule: shard the lists until one shard's codes + iids fit
        # half of the 126 MB L2 (a scan that stays L2-resident), replicate beyond that.  This 12 MB index gives S = 1;
        # BASELINE configs[3] (10M x 16 B + iids = 200 MB) gives S = 4.  MMIDX_LIST_SHARDS overrides.
        index_bytes = N_DB * (M_SUB + 4)
        S_auto = 1
        while index_bytes / S_auto > 64e6 and S_auto < world:
            S_auto *= 2
        sh = HybridIVFPQ(D, N_DB, M_SUB, KS, M.TransformationType.None_, NLIST,
                         int(os.environ.get("MMIDX_LIST_SHARDS", str(S_auto))))
        ix = sh.index
    else:
        sh = None
        ix = M.IVFPQ(D, N_DB, M_SUB, KS, M.TransformationType.None_, NLIST, device=local_rank)
    ix.loadCoarseQuantizer(Cq)
    ix.loadProductQuantizer(P)
    ix.setW(W_PROBE)
    lists, codes = sh.indexAll(X) if sh is not None else ix.indexVectors(None, X, return_codes=True)
    log(f"[bench] rank {rank}: indexed {ix.getLoadCounter()} vectors in {time.time() - t0:.1f}s")
    parity = {}
    foc = prefix + "_oracle_codes.npz"
    if rank == 0 and os.path.exists(foc):  # written by the reference arm on this box: full-size code parity
        z = np.load(foc)
        parity["codes_bit_exact_vs_oracle_1M"] = bool((z["lists"] == lists).all() and (z["codes"] == codes).all())

    dQ = torch.from_numpy(Q).to(dev)
    # ---- exact ground truth for recall (measurement infrastructure, torch) ----
    gt = None
    if rank == 0 and not args.profile:
        dX = torch.from_numpy(X).to(dev)
        x2 = (dX * dX).sum(1)
        gts = []
        for b in range(0, N_GT, 250):
            dd = x2[None, :] - 2.0 * (dQ[b:b + 250] @ dX.T)
            gts.append(torch.topk(dd, TOPK, dim=1, largest=False).indices.cpu().numpy())
        gt = np.concatenate(gts)
        del dX, x2, dd
        torch.cuda.empty_cache()

    d_iids = torch.empty((NQ, TOPK), dtype=torch.int32, device=dev)
    d_dist = torch.empty((NQ, TOPK), dtype=torch.float64, device=dev)
    d_cnt = torch.empty(NQ, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    stream = torch.cuda.current_stream()
    st = C.c_void_p(stream.cuda_stream)
    state = {}

    def step_dev():
        if sh is None:
            _capi.check(lib.mmidx_search_dev(ix._h, NQ, ptr(dQ), TOPK, ptr(d_iids), ptr(d_dist), ptr(d_cnt), st))
            return d_iids, d_dist, d_cnt
        out = sh.search(TOPK, dQ)  # includes the (normally empty) tie check, which reads 4 bytes back
        state["res"] = out
        return out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    nwarm = args.warmup if args.profile else max(args.warmup, 3)
    for _ in range(nwarm):
        step_dev()
    torch.cuda.synchronize()
    launches_per_step = ix.lastLaunches() + (2 if sh is not None and sh.S > 1 else 0)  # + the device merge kernels
    ix.enableTimings(True)
    stage_ms = np.zeros(5)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)  # evict L2 between timed steps (untimed)
        a.record(stream)
        res = step_dev()
        b.record(stream)
        b.synchronize()
        t = ix.lastTimings()
        stage_ms += [t["coarse_ms"], t["lut_ms"], t["scan_ms"], t["merge_ms"], t["total_ms"]]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    ix.enableTimings(False)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    if dist is not None:
        tt = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms = float(tt.item())
    ms_per_step = dev_ms / args.steps
    qps = NQ / (ms_per_step * 1e-3)
    r_iids = res[0].cpu().numpy()
    r_dist = res[1].cpu().numpy()

    if sh is not None and sh.sharded is not None and os.environ.get("MMIDX_SHARD_TIMING"):
        log(f"[bench] rank {rank} phases ms: " + json.dumps({k: round(v, 3) for k, v in sh.sharded.last_phase_ms.items()}))
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "value": qps, "ms_per_step": ms_per_step,
                              "stage_ms_per_step": (stage_ms / args.steps).tolist(),
                              "gpu_launches": launches_per_step * args.steps}), flush=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- end-to-end through the host C-ABI call, pinned host buffers ----
    hQ = torch.from_numpy(Q).pin_memory()
    h_iids = torch.empty((NQ, TOPK), dtype=torch.int32).pin_memory()
    h_dist = torch.empty((NQ, TOPK), dtype=torch.float64).pin_memory()
    h_cnt = torch.empty(NQ, dtype=torch.int32).pin_memory()
    h2d, d2h = hQ.numel() * 8, h_iids.numel() * 4 + h_dist.numel() * 8 + h_cnt.numel() * 4

    def step_e2e():
        if sh is None:
            _capi.check(lib.mmidx_search(ix._h, NQ, ptr(hQ), TOPK, ptr(h_iids), ptr(h_dist), ptr(h_cnt)))
        else:
            dq = hQ.to(dev, non_blocking=True)
            i2, d2, c2 = sh.search(TOPK, dq)
            if rank == 0:
                h_iids.copy_(i2, non_blocking=True)
                h_dist.copy_(d2, non_blocking=True)
                h_cnt.copy_(c2, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(3):
        step_e2e()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        step_e2e()
        e2e_s += time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_qps = NQ * args.steps / e2e_s
    if rank == 0:
        parity["e2e_equals_device_path"] = bool((h_iids.numpy() == r_iids).all() and (h_dist.numpy() == r_dist).all())

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    rec = recall_at_k(r_iids[:N_GT], gt)
    # ---- roofline of the ADC scan kernel: algorithmic bytes = sum over probed lists of len*(m + 4) ----
    scan_bytes = ix.scanBytes(Q) if sh is None else None
    if sh is not None:  # this rank's share: its query group's slice, the probed lists it stores
        q0, q1, _ = sh.slice_of(NQ)
        probes = ix.computeNearestCoarseIndices(Q[q0:q1])
        ls = ix.listSizes()
        scan_bytes = int(ls[probes].sum()) * (M_SUB + 4)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    scan_ms = stage_ms[2] / args.steps
    achieved = scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))["dram_bytes_per_step"]
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": "k_ivfpq_scan_fast", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_step": scan_bytes, "kernel_ms_per_step": scan_ms,
            "note": "algorithmic bytes = sum over probed lists of len*(m+4); the 12 MB code database is L2-resident "
                    "after first touch, the kernel is bound by shared-memory gathers + issue, see DESIGN.md 6"}

    # ---- CPU baseline: the oracle on a bounded query sample, all host cores ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        O, off, cc, ii = cpu_oracle_setup(X, Cq, P, prefix, lists, codes)
        cores = O.num_threads()
        t0 = time.perf_counter()
        O.ivfpq_search(Cq, P, off, cc, ii, Q[: 16 * cores], TOPK, W_PROBE, threads=cores)
        rate = 16 * cores / (time.perf_counter() - t0)
        sample = int(min(NQ, max(16 * cores, rate * 12.0)))
        t0 = time.perf_counter()
        oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q[:sample], TOPK, W_PROBE, threads=cores)
        dt = time.perf_counter() - t0
        cpu = {"value": sample / dt, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} of the {NQ} queries, one pass, query-level threads ({dt:.1f}s)",
               "recall_at_100": recall_at_k(oi[:min(sample, N_GT)], gt[:min(sample, N_GT)])}
        parity["sample_queries"] = sample
        parity["ids_equal_to_oracle"] = bool((oi == r_iids[:sample]).all())
        parity["dist_bit_equal_to_oracle"] = bool((od == r_dist[:sample]).all())
        parity["max_rel_dist_err"] = float(np.max(np.abs(od - r_dist[:sample]) / np.maximum(od, 1e-300)))
    elif world > 1 and not args.no_cpu_baseline:
        # multi-GPU lines: the merged result of a bounded query sample against the oracle (rank 0, untimed)
        try:
            O, off, cc, ii = cpu_oracle_setup(X, Cq, P, prefix, lists, codes)
            ns = 256
            oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q[:ns], TOPK, W_PROBE, threads=O.num_threads())
            parity["sample_queries"] = ns
            parity["ids_equal_to_oracle"] = bool((oi == r_iids[:ns]).all())
            parity["dist_bit_equal_to_oracle"] = bool((od == r_dist[:ns]).all())
        except Exception as e:  # never let the checker break a scaling line
            parity["oracle_sample_error"] = repr(e)[:200]

    out = {
        "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": nwarm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nq_per_step": NQ, "k": TOPK, "l2": "flushed between timed steps (256 MiB write)",
                   "sharding": "none" if world == 1 else
                   f"{sh.S} list shards (balanced list->shard map, per-shard top-k exchanged over NCCL, device merge) x "
                   f"{sh.R} query groups"},
        "recall_at_100": rec,
        "stage_ms_per_step": {"coarse": stage_ms[0] / args.steps, "prep": stage_ms[1] / args.steps,
                              "scan": stage_ms[2] / args.steps, "merge_ties": stage_ms[3] / args.steps,
                              "whole_call": stage_ms[4] / args.steps},
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps, "api": "mmidx_search (host pointers, pinned)"},
        "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "parity": parity,
        "wall_s_timed_region": wall,
    }
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="for runs under ncu: exact warm-up count, no e2e / CPU legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
