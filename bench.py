#!/usr/bin/env python
"""bench.py -- queries/sec + recall@100 of IVFPQ search (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W              the B200 path (libmmidx.so, through the C ABI)
  python bench.py --impl reference --gpus N ...              the CPU arm: the oracle (C restatement of the Java path; the
                                                             reference itself is Java and there is no JVM here) on all
                                                             host cores, a bounded query sample per step
  python bench.py --config 4 --gpus 8                        BASELINE configs[3]: 10M x 128, nlist=8192, w=64, m=16,
                                                             list-sharded over the GPUs (database generated on device)

Default workload = BASELINE configs[2]: IVFPQ 1M x 128 synthetic SIFT-shaped vectors, m=8, ks=256, nlist=1024,
nprobe (w)=32, top-100, 10 000 queries per GPU and step.  A step = one pass of the search path over one batch.
`value` = device-resident queries/s (CUDA events on the launching stream), `e2e` = the same through the host C-ABI call
with pinned HOST buffers (H2D of the queries and D2H of ids + distances inside the timed region).

N > 1 (one process per GPU, torchrun): per-GPU work is fixed ("scaling": "weak"): every rank serves its own 10 000-query
batch against its copy of the 12 MB index (S = 1 list shard x R = N groups) and stores its result rows into the exchange
windows of all ranks over NVLink (multimedia-indexing_b200/csrc/comm.cuh; no NCCL call on the data path -- torch's NCCL
group only carries the set-up handles, the per-step barrier and the max-over-ranks reduction of the timings).  The same
line reports beside it `list_sharded` (S = N: the north_star layout, per-shard queues merged by the slice owners, same
job batch) and `strong_scaling_10k` (10 000 queries in total, split over the ranks).

Timing protocol: W >= 6 warm-up steps (the third call with a given shape is captured into a CUDA graph, once per window
parity); every timed step is preceded by an L2 flush (256 MiB write, untimed) and, for N > 1, a barrier followed by a
common start instant on the node's monotonic clock (`Ctx.aligned_start`: the ranks leave a barrier with a skew of host
wake-up latencies, which would otherwise be measured as waiting for peers' rows inside the step); K steps are timed
per round and rounds repeat until the timed region holds >= 0.5 s; per step the MAX over ranks is taken; mean and median
over all timed steps are reported (`value` uses the mean)."""
import argparse
import ctypes as C
import gc
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CACHE = os.environ.get("MMIDX_BENCH_CACHE", "/tmp/mmidx_bench_cache")
N_GT = 1000  # queries with exact ground truth for recall@100
MIN_TIMED_S = 0.5

WL = {
    3: dict(cfg=3, D=128, M=8, KS=256, NLIST=1024, W=32, K=100, N_DB=1_000_000, NQ=10_000, NTRAIN=100_000, ITERS=20,
            name="IVFPQ 1Mx128 nlist=1024 nprobe=32 m=8 ks=256 top-100 (BASELINE.json configs[2])"),
    4: dict(cfg=4, D=128, M=16, KS=256, NLIST=8192, W=64, K=100, N_DB=10_000_000, NQ=10_000, NTRAIN=262_144, ITERS=10,
            name="IVFPQ 10Mx128 nlist=8192 nprobe=64 m=16 ks=256 top-100 (BASELINE.json configs[3])"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_synth():
    """synth.py by path: the CPU arm must not load libmmidx.so"""
    name = "mmidx_synth_standalone"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "multimedia-indexing_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    return O


def cached(path, make):
    """npz cache on the box so both arms (and all ranks) use the same bytes"""
    os.makedirs(CACHE, exist_ok=True)
    f = os.path.join(CACHE, path)
    if os.path.exists(f):
        z = np.load(f)
        return {k: z[k] for k in z.files}
    out = make()
    tmp = f + f".{os.getpid()}.tmp.npz"
    np.savez(tmp, **out)
    os.replace(tmp, f)
    return out


def quantizers(wl, synth, centers):
    key = f"q_d{wl['D']}_m{wl['M']}_ks{wl['KS']}_nl{wl['NLIST']}_nt{wl['NTRAIN']}_it{wl['ITERS']}.npz"

    def make():
        t0 = time.time()
        Cq, P = synth.train_ivfpq(wl["D"], wl["M"], wl["KS"], wl["NLIST"], ntrain=wl["NTRAIN"], iters=wl["ITERS"], centers=centers)
        log(f"[bench] trained quantizers in {time.time() - t0:.1f}s")
        return dict(Cq=Cq, P=P)

    z = cached(key, make)
    return z["Cq"], z["P"]


def queries_of_group(wl, synth, centers, g):
    """group 0 searches the standard query set (SEED_Q); further groups get fresh noise from the same mixture"""
    return synth.mixture(wl["NQ"], wl["D"], synth.SEED_Q + 7919 * g, centers)


def recall_at_k(ids, gt):
    return float(np.mean([len(set(ids[r]) & set(gt[r])) / gt.shape[1] for r in range(gt.shape[0])]))


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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md); rank 0 only"""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu, enabled=True):
        self.gpu, self.rows, self.proc, self.enabled = gpu, [], None, enabled

    def start(self):
        if not self.enabled:
            return
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
        if not self.enabled:
            return None
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


# =====================================================================================================================
# CPU arm
# =====================================================================================================================
def oracle_csr(wl, synth, X, Cq, P, lists=None, codes=None):
    O = oracle()
    if lists is None:
        key = f"codes_d{wl['D']}_m{wl['M']}_nl{wl['NLIST']}_n{wl['N_DB']}_nt{wl['NTRAIN']}_it{wl['ITERS']}.npz"

        def make():
            t0 = time.time()
            l, c = O.ivfpq_encode(Cq, P, X, threads=O.num_threads())
            log(f"[bench] oracle encoded {len(X)} vectors in {time.time() - t0:.1f}s on {O.num_threads()} threads")
            return dict(lists=l, codes=c.astype(np.uint8))

        z = cached(key, make)
        lists, codes = z["lists"], z["codes"]
    off, cc, ii = synth.csr_from_assignments(lists, np.asarray(codes, dtype=np.uint8), wl["NLIST"])
    return O, off, cc, ii, lists, codes


def cpu_rate(O, wl, Cq, P, off, cc, ii, Q, threads, seconds):
    """queries/s of the oracle on a sample sized to take about `seconds`; returns (qps, sample, dt, result)"""
    probe = min(len(Q), 16 * threads)
    t0 = time.perf_counter()
    O.ivfpq_search(Cq, P, off, cc, ii, Q[:probe], wl["K"], wl["W"], threads=threads)
    rate = probe / (time.perf_counter() - t0)
    sample = int(min(len(Q), max(probe, rate * seconds)))
    t0 = time.perf_counter()
    res = O.ivfpq_search(Cq, P, off, cc, ii, Q[:sample], wl["K"], wl["W"], threads=threads)
    dt = time.perf_counter() - t0
    return sample / dt, sample, dt, res


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = WL[args.config]
    if wl["cfg"] == 4:
        print(json.dumps({"impl": "reference", "unavailable": "config 4 is a GPU-only scaling run; the CPU arm is timed on the default workload"}), flush=True)
        return
    synth = load_synth()
    ce = synth.mixture_centers(wl["D"])
    Cq, P = quantizers(wl, synth, ce)
    X = synth.mixture(wl["N_DB"], wl["D"], synth.SEED_DB, ce)
    Q = queries_of_group(wl, synth, ce, 0)
    O, off, cc, ii, _, _ = oracle_csr(wl, synth, X, Cq, P)
    cores = O.num_threads()
    # bounded sample per step: ~2 s of wall time on all cores
    _, sample, _, _ = cpu_rate(O, wl, Cq, P, off, cc, ii, Q, cores, 2.0)
    for _ in range(args.warmup):
        O.ivfpq_search(Cq, P, off, cc, ii, Q[:sample], wl["K"], wl["W"], threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids, dist, cnt = O.ivfpq_search(Cq, P, off, cc, ii, Q[:sample], wl["K"], wl["W"], threads=cores)
    dt = time.perf_counter() - t0
    qps = sample * args.steps / dt
    ngt = min(sample, 200)
    rec = recall_at_k(ids[:ngt], exact_gt_cpu(X, Q[:ngt], wl["K"]))
    t1_qps, t1_sample, t1_dt, _ = cpu_rate(O, wl, Cq, P, off, cc, ii, Q, 1, 3.0)
    desc = f"first {sample} of the {wl['NQ']} queries per step, query-level threads over a shared read-only index"
    print(json.dumps({
        "impl": "reference", "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "nq_per_step": wl["NQ"] * args.gpus, "k": wl["K"]},
        "recall_at_100": rec,
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": desc,
                         "single_thread": {"value": t1_qps, "sample": f"first {t1_sample} queries, 1 thread ({t1_dt:.1f}s): how every reference driver issues queries (Example.java:101-112)"},
                         "note": "C restatement of the Java path (oracle/), binary heap queue, no per-offer allocation: the strongest CPU "
                                 "form of the reference's algorithm; the Java reference itself cannot run here (no JVM)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# =====================================================================================================================
# B200 arm
# =====================================================================================================================
class Ctx:
    """torch / torch.distributed plumbing of one rank"""

    def __init__(self, rank, world, local_rank):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        # a side stream is made torch's current one: the legacy default stream cannot be captured into a CUDA graph
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.st = C.c_void_p(self.stream.cuda_stream)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def aligned_start(self, lead_s=250e-6):
        """N > 1, after the barrier: all ranks agree on one instant of the node's monotonic clock (the latest rank's
        now + lead) and spin until then, so that the timed step starts together on every rank.  Without it the skew with
        which the ranks leave the barrier (tens to hundreds of microseconds of host wake-up) is measured as time the
        early ranks spend waiting for the late ranks' rows inside the step."""
        if self.dist is None:
            return
        t = self.torch.tensor([time.perf_counter() + lead_s], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        t0 = float(t.item())
        while time.perf_counter() < t0:
            pass

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def timed(self, step, steps, warmup, after=None, wall=False, min_s=MIN_TIMED_S, max_rounds=40):
        """The timing protocol of the module docstring.  wall: host clock around a synchronous call (e2e) instead of
        CUDA events.  Returns per-step times in ms (max over ranks), all rounds."""
        torch = self.torch
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        out = []
        gc.collect()
        gc.disable()  # a collection between the start event and the launch would be timed on this rank and, through the max, on all
        try:
            return self._timed_rounds(step, steps, after, wall, min_s, max_rounds, out)
        finally:
            gc.enable()

    def _timed_rounds(self, step, steps, after, wall, min_s, max_rounds, out):
        torch = self.torch
        for _ in range(max_rounds):
            ts = []
            for _ in range(steps):
                self.flush.fill_(1)  # evict L2 between timed steps (untimed)
                self.barrier()
                self.aligned_start()
                if wall:
                    t0 = time.perf_counter()
                    step()
                    ts.append(1e3 * (time.perf_counter() - t0))
                else:
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(self.stream)
                    step()
                    b.record(self.stream)
                    b.synchronize()
                    ts.append(a.elapsed_time(b))
                if after:
                    after()
            out.extend(self.max_over_ranks(ts).tolist())
            if sum(out) * 1e-3 >= min_s:
                break
        return out

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


MULTI_STAGES = ("coarse", "prep", "scan", "after_scan", "whole_call", "exchange_points", "merge", "tie_pass")


def summarize(times_ms, queries_per_step):
    mean, med = float(np.mean(times_ms)), float(np.median(times_ms))
    pct = {f"p{q}": float(np.percentile(times_ms, q)) for q in (10, 90, 99)}
    pct["max"] = float(np.max(times_ms))
    return {"value": queries_per_step / (mean * 1e-3), "ms_per_step": mean, "median_ms_per_step": med,
            "value_at_median": queries_per_step / (med * 1e-3), "timed_steps": len(times_ms),
            "timed_region_s": float(np.sum(times_ms)) * 1e-3, "step_ms_percentiles": pct}


def ptr(t):
    return C.c_void_p(t.data_ptr())


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in p:
            return p["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def tie_counts(ctx, M, ix, wl, Cq, dQ):
    """How many queries have an exact binary64 tie AT the k-th (w-th) boundary: only those depend on the LingPipe tie rule
    that the oracle restates from memory.  k+1 (w+1) results, then dist[k-1] == dist[k]."""
    torch, lib = ctx.torch, M._capi.lib
    nq, k, w = dQ.shape[0], wl["K"], wl["W"]
    ii = torch.empty((nq, k + 1), dtype=torch.int32, device=ctx.dev)
    dd = torch.empty((nq, k + 1), dtype=torch.float64, device=ctx.dev)
    cc = torch.empty(nq, dtype=torch.int32, device=ctx.dev)
    M._capi.check(lib.mmidx_search_dev(ix._h, nq, ptr(dQ), k + 1, ptr(ii), ptr(dd), ptr(cc), ctx.st))
    torch.cuda.synchronize()
    at_k = int(((dd[:, k - 1] == dd[:, k]) & (cc > k)).sum().item())
    inside = int(((dd[:, 1:k] == dd[:, :k - 1]).any(dim=1)).sum().item())
    # coarse stage: exact distances to the centroids = a Linear index over the centroids (same squared terms, bit for bit)
    lin = M.Linear(wl["D"], wl["NLIST"], device=ctx.dev.index)
    lin.indexVectors(None, Cq)
    li = torch.empty((nq, w + 1), dtype=torch.int32, device=ctx.dev)
    ld = torch.empty((nq, w + 1), dtype=torch.float64, device=ctx.dev)
    lc = torch.empty(nq, dtype=torch.int32, device=ctx.dev)
    M._capi.check(lib.mmidx_search_dev(lin._h, nq, ptr(dQ), w + 1, ptr(li), ptr(ld), ptr(lc), ctx.st))
    torch.cuda.synchronize()
    coarse = int((ld[:, w - 1] == ld[:, w]).sum().item())
    lin.close()
    return {"queries": nq, "ties_at_k_boundary": at_k, "coarse_ties_at_w_boundary": coarse,
            "queries_with_equal_distances_inside_top_k": inside,
            "note": "0 boundary ties => the id sets and distances of this run do not depend on any queue tie rule"}


def small_batch(ctx, M, ix, wl, Q, cpu):
    """latency of the reference's own call shape: a few queries per computeNearestNeighbors call
    (AbstractSearchStructure.java:281-291), host buffers in and out through mmidx_search"""
    torch, lib = ctx.torch, M._capi.lib
    k = wl["K"]
    out = {}
    for nq, calls in ((1, 400), (32, 200), (1024, 60)):
        hQ = torch.from_numpy(Q[:nq].copy()).pin_memory()
        hi = torch.empty((nq, k), dtype=torch.int32).pin_memory()
        hd = torch.empty((nq, k), dtype=torch.float64).pin_memory()
        hc = torch.empty(nq, dtype=torch.int32).pin_memory()

        def call():
            M._capi.check(lib.mmidx_search(ix._h, nq, ptr(hQ), k, ptr(hi), ptr(hd), ptr(hc)))

        for _ in range(8):
            call()
        ts = []
        for _ in range(calls):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        med = float(np.median(ts))
        row = {"us_per_call": 1e6 * med, "queries_per_s": nq / med, "launches_per_call": ix.lastLaunches()}
        if cpu is not None:
            O, Cq, P, off, cc, ii = cpu
            for name, th in (("cpu_1_thread", 1), ("cpu_all_threads", O.num_threads())):
                reps = 5 if nq >= 32 else 20
                t0 = time.perf_counter()
                for _ in range(reps):
                    oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q[:nq], k, wl["W"], threads=min(th, nq))
                dt = (time.perf_counter() - t0) / reps
                row[name] = {"us_per_call": 1e6 * dt, "queries_per_s": nq / dt}
            row["equal_to_oracle"] = bool((hi.numpy() == oi).all() and (hd.numpy() == od).all())
        out[str(nq)] = row
    return out


def other_rows(ctx, M, wl, synth, X, Q, ce, have_cpu):
    """SURVEY 8 rows next to the headline (BASELINE configs[0], [1], [4]): device-resident timing, bounded to seconds"""
    torch, lib = ctx.torch, M._capi.lib
    O = oracle() if have_cpu else None
    rows = {}
    sm_count = torch.cuda.get_device_properties(ctx.dev).multi_processor_count
    fp64_peak = sm_count * 64 * 1.965e9  # non-fused binary64 operations/s (64 lanes per SM), nominal

    def timeit(fn, min_s=0.3):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        while sum(ts) < min_s * 1e3 and len(ts) < 400:
            ctx.flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(ctx.stream)
            fn()
            b.record(ctx.stream)
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))

    def dev_search(ix, dQ, k):
        nq = dQ.shape[0]
        ii = torch.empty((nq, k), dtype=torch.int32, device=ctx.dev)
        dd = torch.empty((nq, k), dtype=torch.float64, device=ctx.dev)
        cc = torch.empty(nq, dtype=torch.int32, device=ctx.dev)
        return (lambda: M._capi.check(lib.mmidx_search_dev(ix._h, nq, ptr(dQ), k, ptr(ii), ptr(dd), ptr(cc), ctx.st))), ii, dd, cc

    # ---- configs[0]: Linear, 10k x 64, k = 10 ----
    try:
        d1, n1, k1 = 64, 10_000, 10
        c1 = synth.mixture_centers(d1, 256)
        X1, Q1 = synth.mixture(n1, d1, synth.SEED_DB, c1), synth.mixture(10_000, d1, synth.SEED_Q, c1)
        lin = M.Linear(d1, n1, device=ctx.dev.index)
        lin.indexVectors(None, X1)
        dQ1 = torch.from_numpy(Q1).to(ctx.dev)
        fn, ii, dd, cc = dev_search(lin, dQ1, k1)
        ms = timeit(fn)
        row = {"workload": "Linear 10k x 64, top-10, 10 000 queries/step (BASELINE configs[0])", "queries_per_s": 1e4 / (ms * 1e-3), "ms_per_step": ms,
               "fp64_ops_per_s": 3.0 * n1 * d1 * 1e4 / (ms * 1e-3), "fp64_frac_of_nominal": 3.0 * n1 * d1 * 1e4 / (ms * 1e-3) / fp64_peak}
        if O is not None:
            ns = 1000
            t0 = time.perf_counter()
            oi, od, oc = O.linear_search(X1, Q1[:ns], k1, threads=O.num_threads())
            row["cpu_queries_per_s"] = ns / (time.perf_counter() - t0)
            row["parity_queries"] = ns
            row["equal_to_oracle"] = bool((ii[:ns].cpu().numpy() == oi).all() and (dd[:ns].cpu().numpy() == od).all())
        lin.close()
        rows["linear"] = row
    except Exception as e:  # a row never breaks the headline
        rows["linear"] = {"error": repr(e)[:300]}

    # ---- configs[1]: flat PQ over the 1M database, m = 8, top-100 ----
    try:
        z = cached(f"pq_d{wl['D']}_m{wl['M']}_ks{wl['KS']}.npz", lambda: dict(P=synth.train_pq(wl["D"], wl["M"], wl["KS"], ntrain=50_000, iters=10, centers=ce)))
        Pf = z["P"]
        pq = M.PQ(wl["D"], wl["N_DB"], wl["M"], wl["KS"], M.TransformationType.None_, device=ctx.dev.index)
        pq.loadProductQuantizer(Pf)
        dX = torch.from_numpy(X).to(ctx.dev)
        dcodes = torch.empty((wl["N_DB"], wl["M"]), dtype=torch.uint8, device=ctx.dev)
        warm = M.PQ(wl["D"], 20_000, wl["M"], wl["KS"], M.TransformationType.None_, device=ctx.dev.index)  # loads the encode kernels
        warm.loadProductQuantizer(Pf)
        M._capi.check(lib.mmidx_add_dev(warm._h, 20_000, ptr(dX), None, ptr(dcodes)))
        warm.close()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        M._capi.check(lib.mmidx_add_dev(pq._h, wl["N_DB"], ptr(dX), None, ptr(dcodes)))
        torch.cuda.synchronize()
        enc_s = time.perf_counter() - t0
        nq2 = 2000
        dQ2 = torch.from_numpy(Q[:nq2].copy()).to(ctx.dev)
        fn, ii, dd, cc = dev_search(pq, dQ2, wl["K"])
        ms = timeit(fn)
        pk, _ = peaks()
        gbs = nq2 * wl["N_DB"] * wl["M"] / (ms * 1e-3) / 1e9
        row = {"workload": "PQ 1Mx128 m=8 ks=256 top-100, 2000 queries/step (BASELINE configs[1])", "queries_per_s": nq2 / (ms * 1e-3), "ms_per_step": ms,
               "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / pk, "regime": "codes L2-resident (8 MB)",
               "pq_encode_vectors_per_s": wl["N_DB"] / enc_s,
               "pq_encode_fp64_ops_per_s": 3.0 * wl["M"] * wl["KS"] * (wl["D"] // wl["M"]) * wl["N_DB"] / enc_s}
        if O is not None:
            ns = 1000
            codes = dcodes.cpu().numpy()
            t0 = time.perf_counter()
            oi, od, oc = O.pq_search(Pf, codes, Q[:ns], wl["K"], threads=O.num_threads())
            row["cpu_queries_per_s"] = ns / (time.perf_counter() - t0)
            row["parity_queries"] = ns
            row["equal_to_oracle"] = bool((ii[:ns].cpu().numpy() == oi).all() and (dd[:ns].cpu().numpy() == od).all())
            ne = 20_000
            oc2 = O.pq_encode(Pf, X[:ne], threads=O.num_threads())
            row["codes_equal_to_oracle_first_20000"] = bool((np.asarray(oc2) == codes[:ne]).all())
        pq.close()
        # IVFPQ.indexVectorInternal (coarse assign + residual + encode, IVFPQ.java:309-355) over the same device-resident vectors
        try:
            Cq3, P3 = quantizers(wl, synth, ce)
            iv = M.IVFPQ(wl["D"], wl["N_DB"], wl["M"], wl["KS"], M.TransformationType.None_, wl["NLIST"], device=ctx.dev.index)
            iv.loadCoarseQuantizer(Cq3)
            iv.loadProductQuantizer(P3)
            dl = torch.empty(wl["N_DB"], dtype=torch.int32, device=ctx.dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            M._capi.check(lib.mmidx_add_dev(iv._h, wl["N_DB"], ptr(dX), ptr(dl), ptr(dcodes)))
            torch.cuda.synchronize()
            row["ivfpq_index_vectors_per_s"] = wl["N_DB"] / (time.perf_counter() - t0)
            iv.close()
            del dl
        except Exception as e:
            row["ivfpq_index_error"] = repr(e)[:200]
        del dX, dcodes
        rows["pq_flat"] = row
    except Exception as e:
        rows["pq_flat"] = {"error": repr(e)[:300]}

    # ---- configs[4]: VLAD aggregate, 100k images x ~1000 SURF(64-d) descriptors, codebook 128 ----
    try:
        K5, D5, chunk_img, passes = 128, 64, 2000, 50
        desc, offs = synth.descriptors(chunk_img, D5)
        cb = cached(f"vlad_cb_k{K5}_d{D5}.npz", lambda: dict(cb=synth.kmeans(synth.descriptors(200, D5, seed=77)[0], K5, 10, seed=5)))["cb"]
        dcb, ddesc = torch.from_numpy(cb).to(ctx.dev), torch.from_numpy(desc).to(ctx.dev)
        doff = torch.from_numpy(offs).to(ctx.dev)
        dout = torch.empty((chunk_img, K5 * D5), dtype=torch.float64, device=ctx.dev)
        dasg = torch.empty(desc.shape[0], dtype=torch.int32, device=ctx.dev)

        def vlad():
            M._capi.check(lib.mmidx_vlad_dev(ptr(dcb), K5, D5, chunk_img, ptr(doff), desc.shape[0], ptr(ddesc), ptr(dout), ptr(dasg), ctx.st))

        for _ in range(2):
            vlad()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ctx.stream)
        for _ in range(passes):
            vlad()
        b.record(ctx.stream)
        b.synchronize()
        s = a.elapsed_time(b) * 1e-3
        nd = desc.shape[0] * passes
        row = {"workload": f"VLAD K=128, {chunk_img * passes} images x ~1000 SURF-like 64-d descriptors ({passes} passes over a {chunk_img}-image chunk resident in HBM; BASELINE configs[4])",
               "descriptors_per_s": nd / s, "images_per_s": chunk_img * passes / s, "seconds": s,
               "fp64_ops_per_s": 3.0 * K5 * D5 * nd / s, "fp64_frac_of_nominal": 3.0 * K5 * D5 * nd / s / fp64_peak}
        if O is not None:
            ni = 100
            t0 = time.perf_counter()
            ov, oa = O.vlad(cb, desc[:offs[ni]], offs[:ni + 1], threads=O.num_threads())
            row["cpu_descriptors_per_s"] = offs[ni] / (time.perf_counter() - t0)
            row["parity_images"] = ni
            row["equal_to_oracle"] = bool((dout[:ni].cpu().numpy() == ov.reshape(ni, -1)).all() and (dasg[:offs[ni]].cpu().numpy() == oa).all())
        rows["vlad"] = row
    except Exception as e:
        rows["vlad"] = {"error": repr(e)[:300]}
    return rows


def run_gpu(args, rank, world, local_rank):
    wl = WL[args.config]
    ctx = Ctx(rank, world, local_rank)
    torch = ctx.torch
    import mmidx_b200 as M
    from multimedia_indexing_b200 import synth
    from multimedia_indexing_b200.sharded import MultiIVFPQ, balanced_shard_map

    if wl["cfg"] == 4:
        run_gpu_cfg4(args, ctx, M, synth, wl)
        return
    lib = M._capi.lib
    NQ, K, W, D = wl["NQ"], wl["K"], wl["W"], wl["D"]
    ce = synth.mixture_centers(D)
    if world > 1:  # rank 0 trains (or finds the cache the CPU arm left), everybody loads the same file
        if rank == 0:
            quantizers(wl, synth, ce)
        ctx.barrier()
    Cq, P = quantizers(wl, synth, ce)
    X = synth.mixture(wl["N_DB"], D, synth.SEED_DB, ce)
    G = world
    group = rank  # main layout: S = 1, every rank is a group
    Qg = queries_of_group(wl, synth, ce, group)
    extras = not args.profile and not args.quick

    # ---- build the index (untimed): GPU coarse-assign + residual + PQ encode of the whole database ----
    t0 = time.time()
    if world > 1:
        mi = MultiIVFPQ(D, wl["N_DB"], wl["M"], wl["KS"], M.TransformationType.None_, wl["NLIST"], list_shards=1)
        ix = mi.index
    else:
        mi = None
        ix = M.IVFPQ(D, wl["N_DB"], wl["M"], wl["KS"], M.TransformationType.None_, wl["NLIST"], device=local_rank)
    ix.loadCoarseQuantizer(Cq)
    ix.loadProductQuantizer(P)
    ix.setW(W)
    lists, codes = ix.indexVectors(None, X, return_codes=True)
    log(f"[bench] rank {rank}: indexed {ix.getLoadCounter()} vectors in {time.time() - t0:.1f}s")
    if mi is not None:
        mi.connect(max_gq=NQ, k_max=K + 1)

    dQ = torch.from_numpy(Qg).to(ctx.dev)
    d_iids = torch.empty((NQ, K), dtype=torch.int32, device=ctx.dev)
    d_dist = torch.empty((NQ, K), dtype=torch.float64, device=ctx.dev)
    d_cnt = torch.empty(NQ, dtype=torch.int32, device=ctx.dev)
    state = {}

    def step_dev():
        if mi is None:
            M._capi.check(lib.mmidx_search_dev(ix._h, NQ, ptr(dQ), K, ptr(d_iids), ptr(d_dist), ptr(d_cnt), ctx.st))
            state["res"] = (d_iids, d_dist, d_cnt)
        else:
            state["res"] = mi.search(K, dQ, gather_all=True)[:3]

    # ---- device-resident timing (stage events on: they are part of what is timed) ----
    nwarm = args.warmup if args.profile else max(args.warmup, 6)
    ix.enableTimings(True)
    stage = []
    clocks = ClockSampler(local_rank, enabled=(rank == 0))
    ctx.barrier()
    clocks.start()
    wall0 = time.perf_counter()
    times = ctx.timed(step_dev, args.steps, nwarm, after=lambda: stage.append(list(ix.lastTimings().values())),
                      min_s=0.0 if args.profile else MIN_TIMED_S)
    wall = time.perf_counter() - wall0
    clk = clocks.stop()
    ix.enableTimings(False)
    launches_per_step = ix.lastLaunches()
    main = summarize(times, NQ * G)
    stage_ms = np.mean(np.asarray(stage), axis=0)
    res = state["res"]
    r_iids, r_dist = res[0].cpu().numpy(), res[1].cpu().numpy()  # job-wide rows (all groups) when world > 1

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, **main, "stage_ms_per_step": stage_ms.tolist(),
                              "gpu_launches": launches_per_step * len(times)}), flush=True)
        ctx.close()
        return

    # ---- end-to-end through the host C-ABI call, pinned host buffers; every rank moves its own batch ----
    hQ = torch.from_numpy(Qg).pin_memory()
    h_iids = torch.empty((NQ, K), dtype=torch.int32).pin_memory()
    h_dist = torch.empty((NQ, K), dtype=torch.float64).pin_memory()
    h_cnt = torch.empty(NQ, dtype=torch.int32).pin_memory()
    h2d, d2h = hQ.numel() * 8, h_iids.numel() * 4 + h_dist.numel() * 8 + h_cnt.numel() * 4

    def step_e2e():
        M._capi.check(lib.mmidx_search(ix._h, NQ, ptr(hQ), K, ptr(h_iids), ptr(h_dist), ptr(h_cnt)))

    e2e = summarize(ctx.timed(step_e2e, args.steps, 3, wall=True), NQ * G)
    parity = {}
    my_rows = slice(group * NQ, (group + 1) * NQ) if world > 1 else slice(0, NQ)
    parity["e2e_equals_device_path"] = bool((h_iids.numpy() == r_iids[my_rows]).all() and (h_dist.numpy() == r_dist[my_rows]).all())

    # ---- extra layouts on the same job batch (N > 1) ----
    strong = sharded = None
    if world > 1 and not args.quick:
        # (a) strong scaling: the standard 10 000 queries in total, NQ / G per rank
        per = (NQ + G - 1) // G
        Q0 = queries_of_group(wl, synth, ce, 0)
        Qs = np.concatenate([Q0, np.repeat(Q0[:1], per * G - NQ, axis=0)])[group * per:(group + 1) * per]
        dQs = torch.from_numpy(np.ascontiguousarray(Qs)).to(ctx.dev)
        st8 = {}

        def step_strong():
            st8["res"] = mi.search(K, dQs, gather_all=True)[:3]

        strong = summarize(ctx.timed(step_strong, args.steps, 6), NQ)
        strong["nq_per_step"] = NQ
        s_iids, s_dist = st8["res"][0].cpu().numpy(), st8["res"][1].cpu().numpy()
        # (b) list sharding, S = G: the north_star layout, same job batch as the headline
        t0 = time.time()
        ms = MultiIVFPQ(D, wl["N_DB"], wl["M"], wl["KS"], M.TransformationType.None_, wl["NLIST"], list_shards=G)
        ms.loadCoarseQuantizer(Cq)
        ms.loadProductQuantizer(P)
        ms.setW(W)
        ms.setShardMap(balanced_shard_map(np.bincount(lists, minlength=wl["NLIST"]), G))
        ms.index.indexPQCodes(None, lists, codes)
        ms.connect(max_gq=NQ * G, k_max=K)
        Qall = np.concatenate([queries_of_group(wl, synth, ce, g) for g in range(G)])
        dQall = torch.from_numpy(Qall).to(ctx.dev)
        log(f"[bench] rank {rank}: list-sharded index ({int(ms.listSizes().sum())} of {wl['N_DB']} vectors here) in {time.time() - t0:.1f}s")
        sh8 = {}

        def step_sharded():
            sh8["res"] = ms.search(K, dQall, gather_all=True)[:3]

        ms.index.enableTimings(True)
        sh_stage = []
        sharded = summarize(ctx.timed(step_sharded, args.steps, 6, after=lambda: sh_stage.append(list(ms.index.lastTimingsMulti().values()))), NQ * G)
        ms.index.enableTimings(False)
        sharded.update(S=G, R=1, nq_per_step=NQ * G, launches_per_step=ms.index.lastLaunches(),
                       stage_ms_per_step=dict(zip(MULTI_STAGES, np.mean(np.asarray(sh_stage), axis=0).tolist())),
                       vectors_on_rank0=int(ms.listSizes().sum()))
        l_iids, l_dist = sh8["res"][0].cpu().numpy(), sh8["res"][1].cpu().numpy()
        hQall = torch.from_numpy(Qall).pin_memory()
        sl = (NQ * G + G - 1) // G
        hs_i = torch.empty((sl, K), dtype=torch.int32).pin_memory()
        hs_d = torch.empty((sl, K), dtype=torch.float64).pin_memory()
        hs_c = torch.empty(sl, dtype=torch.int32).pin_memory()
        sharded["e2e"] = summarize(ctx.timed(lambda: ms.search_host(K, hQall, hs_i, hs_d, hs_c), args.steps, 3, wall=True), NQ * G)
        sharded["e2e"].update(h2d_bytes_per_step=hQall.numel() * 8, d2h_bytes_per_step=sl * K * 12 + sl * 4,
                              note="every shard needs the whole job batch of queries; each rank reads back its slice of the results")
        sharded["parity"] = {"equals_replica_layout_all_rows": bool((l_iids == r_iids).all() and (l_dist == r_dist).all())}
        ms.close()

    if rank != 0:
        ctx.close()
        return

    # ---- rank 0: parity against the oracle, recall, roofline, CPU baseline, extras ----
    gt = None
    try:
        dX = torch.from_numpy(X).to(ctx.dev)
        x2 = (dX * dX).sum(1)
        gts = []
        dQ0 = torch.from_numpy(queries_of_group(wl, synth, ce, 0)[:N_GT]).to(ctx.dev)
        for b in range(0, N_GT, 250):
            dd = x2[None, :] - 2.0 * (dQ0[b:b + 250] @ dX.T)
            gts.append(torch.topk(dd, K, dim=1, largest=False).indices.cpu().numpy())
        gt = np.concatenate(gts)
        del dX, x2, dd
        torch.cuda.empty_cache()
    except Exception as e:
        log(f"[bench] ground truth failed: {e!r}")
    rec = recall_at_k(r_iids[:N_GT], gt) if gt is not None else None

    scan_bytes = ix.scanBytes(Qg)
    peak, peak_src = peaks()
    scan_ms = float(stage_ms[2])
    achieved = scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else None
    traffic = None
    if world == 1:  # the committed ncu capture is of this exact launch (N = 1, 10 000 queries); not transferable to other N
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))["dram_bytes_per_step"]
        except Exception:
            pass
    roof = {"bound": "hbm", "kernel": "k_ivfpq_scan_fast", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": scan_bytes, "kernel_ms_per_launch": scan_ms,
            "regime": "L2-resident: the 12 MB code database stays in the 126 MB L2, so DRAM traffic (ncu, one launch on one GPU) is a few "
                      "percent of the algorithmic bytes and the kernel's physical limits are the shared-memory pipe and issue slots; the "
                      "HBM-streaming regime of the same kernel is profiles/r2_hbm_regime.* (DESIGN.md 6)"}

    cpu = None
    O = None
    if not args.no_cpu_baseline:
        O, off, cc, ii, _, _ = oracle_csr(wl, synth, X, Cq, P, lists, codes)
        cores = O.num_threads()
        qps, sample, dt, (oi, od, oc) = cpu_rate(O, wl, Cq, P, off, cc, ii, Qg, cores, 12.0 if world == 1 else 2.0)
        cpu = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} of the {NQ} queries of rank 0, one pass, query-level threads ({dt:.1f}s)"}
        if gt is not None:
            cpu["recall_at_100"] = recall_at_k(oi[:min(sample, N_GT)], gt[:min(sample, N_GT)])
        if world == 1:
            t1_qps, t1_sample, t1_dt, _ = cpu_rate(O, wl, Cq, P, off, cc, ii, Qg, 1, 3.0)
            cpu["single_thread"] = {"value": t1_qps, "sample": f"first {t1_sample} queries, 1 thread ({t1_dt:.1f}s)"}
            O.set_faithful_costs(True)  # BASELINE.md variant (A): the reference's allocations per candidate
            fa_qps, fa_sample, fa_dt, (fi, fd, fc) = cpu_rate(O, wl, Cq, P, off, cc, ii, Qg, cores, 3.0)
            O.set_faithful_costs(False)
            cpu["allocation_faithful"] = {"value": fa_qps, "cores": cores, "same_results": bool((fi == oi[:fa_sample]).all() and (fd == od[:fa_sample]).all()),
                                          "sample": f"first {fa_sample} queries ({fa_dt:.1f}s): per candidate a code copy, a Result and, when accepted, a queue "
                                                    "entry are heap-allocated as in IVFPQ.java:434,443-445"}
        parity["sample_queries"] = sample
        parity["ids_equal_to_oracle"] = bool((oi == r_iids[:sample]).all())
        parity["dist_bit_equal_to_oracle"] = bool((od == r_dist[:sample]).all())
        parity["max_rel_dist_err"] = float(np.max(np.abs(od - r_dist[:sample]) / np.maximum(od, 1e-300)))
        foc = os.path.join(CACHE, f"codes_d{wl['D']}_m{wl['M']}_nl{wl['NLIST']}_n{wl['N_DB']}_nt{wl['NTRAIN']}_it{wl['ITERS']}.npz")
        if os.path.exists(foc):  # written by the reference arm on this box: full-size code parity
            z = np.load(foc)
            parity["codes_bit_exact_vs_oracle_1M"] = bool((z["lists"] == lists).all() and (z["codes"] == codes).all())
        if world > 1:  # rows of the other groups (their own query sets) and the other layouts, 256 queries each
            ns = 256
            okg = True
            for g in range(1, G):
                Qo = queries_of_group(wl, synth, ce, g)[:ns]
                gi, gd, gc = O.ivfpq_search(Cq, P, off, cc, ii, Qo, K, W, threads=cores)
                okg &= bool((gi == r_iids[g * NQ:g * NQ + ns]).all() and (gd == r_dist[g * NQ:g * NQ + ns]).all())
            parity["other_groups_256_each_equal_to_oracle"] = okg
            if strong is not None:
                per = (NQ + G - 1) // G
                Q0 = queries_of_group(wl, synth, ce, 0)
                gi, gd, gc = O.ivfpq_search(Cq, P, off, cc, ii, Q0[:ns], K, W, threads=cores)
                strong["parity_256_equal_to_oracle"] = bool((gi == s_iids[:ns]).all() and (gd == s_dist[:ns]).all())
            if sharded is not None:
                sharded["parity"]["first_256_equal_to_oracle"] = bool((oi[:ns] == l_iids[:ns]).all() and (od[:ns] == l_dist[:ns]).all())

    out = {
        "metric": "queries/sec", "value": main["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": nwarm, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "nq_per_step": NQ * G, "k": K},
        "layout": "1 GPU" if world == 1 else f"1 list shard x {G} groups: every rank searches its own {NQ}-query batch over its copy of the index "
                  "and stores its rows into all ranks' exchange windows (NVLink P2P, no collective call)",
        "l2": "flushed between timed steps (256 MiB write)",
        "median_ms_per_step": main["median_ms_per_step"], "value_at_median": main["value_at_median"],
        "step_ms_percentiles": main.get("step_ms_percentiles"),
        "timed_steps": main["timed_steps"], "timed_region_s": main["timed_region_s"],
        "training": f"{wl['NTRAIN']} points, {wl['ITERS']} Lloyd iterations (SURVEY 8d)",
        "recall_at_100": rec,
        "stage_ms_per_step": dict(zip(("coarse", "prep", "scan", "merge_ties_exchange", "whole_call"), stage_ms.tolist())),
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e["value"], "unit": "queries/s", "h2d_bytes_per_step": h2d * G, "d2h_bytes_per_step": d2h * G,
                "ms_per_step": e2e["ms_per_step"], "median_ms_per_step": e2e["median_ms_per_step"], "timed_steps": e2e["timed_steps"],
                "api": "mmidx_search (host pointers, pinned), one call per rank and step on its own batch"},
        "gpu_launches": launches_per_step * main["timed_steps"], "launches_per_step": launches_per_step,
        "clocks": clk, "parity": parity, "wall_s_timed_region": wall,
    }
    if strong is not None:
        out["strong_scaling_10k"] = strong
    if sharded is not None:
        out["list_sharded"] = sharded
    if extras:
        t0 = time.time()
        try:
            out["ties"] = tie_counts(ctx, M, ix, wl, Cq, dQ)
        except Exception as e:
            out["ties"] = {"error": repr(e)[:300]}
        if world == 1:
            try:
                out["small_batch"] = small_batch(ctx, M, ix, wl, Qg, (O, Cq, P, off, cc, ii) if O is not None else None)
            except Exception as e:
                out["small_batch"] = {"error": repr(e)[:300]}
            out["rows"] = other_rows(ctx, M, wl, synth, X, Qg, ce, O is not None)
        log(f"[bench] extras in {time.time() - t0:.1f}s")
    print(json.dumps(out), flush=True)
    ctx.close()


# =====================================================================================================================
# BASELINE configs[3]: 10M x 128, nlist = 8192, w = 64, m = 16, list-sharded across the GPUs
# =====================================================================================================================
def gen_on_device(torch, dev, centers_dev, n, seed, chunk=500_000):
    """the SIFT-shaped mixture of synth.mixture, generated on the device (identical on every rank: same seeds)"""
    for b in range(0, n, chunk):
        nb = min(chunk, n - b)
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 1_000_003 + b)
        c = torch.randint(0, centers_dev.shape[0], (nb,), generator=g, device=dev)
        x = centers_dev[c] + 20.0 * torch.randn((nb, centers_dev.shape[1]), generator=g, device=dev, dtype=torch.float64)
        yield b, torch.clamp(torch.round(x), 0.0, 255.0)


def kmeans_dev(torch, X, k, iters, seed):
    g = torch.Generator(device=X.device)
    g.manual_seed(seed)
    C_ = X[torch.randperm(X.shape[0], generator=g, device=X.device)[:k]].clone()
    for _ in range(iters):
        a = torch.empty(X.shape[0], dtype=torch.int64, device=X.device)
        c2 = (C_ * C_).sum(1)
        for b in range(0, X.shape[0], 32768):
            a[b:b + 32768] = torch.argmin(c2[None, :] - 2.0 * (X[b:b + 32768] @ C_.T), dim=1)
        cnt = torch.bincount(a, minlength=k).to(X.dtype)
        S_ = torch.zeros_like(C_).index_add_(0, a, X)
        nz = cnt > 0
        C_[nz] = S_[nz] / cnt[nz, None]
    return C_


def run_gpu_cfg4(args, ctx, M, synth, wl):
    from multimedia_indexing_b200.sharded import MultiIVFPQ, balanced_shard_map
    torch, lib = ctx.torch, M._capi.lib
    rank, world = ctx.rank, ctx.world
    NQ, K, W, D, N_DB = wl["NQ"], wl["K"], wl["W"], wl["D"], (args.n_db or wl["N_DB"])
    ce = synth.mixture_centers(D)
    dce = torch.from_numpy(ce).to(ctx.dev)

    def make_q():
        t0 = time.time()
        T = torch.cat([x for _, x in gen_on_device(torch, ctx.dev, dce, wl["NTRAIN"], synth.SEED_TRAIN)])
        Cq = kmeans_dev(torch, T, wl["NLIST"], wl["ITERS"], 11)
        c2 = (Cq * Cq).sum(1)
        a = torch.cat([torch.argmin(c2[None, :] - 2.0 * (T[b:b + 32768] @ Cq.T), dim=1) for b in range(0, T.shape[0], 32768)])
        R_ = Cq[a] - T  # the reference's residual sign (IVFPQ.java:645)
        S_ = D // wl["M"]
        P = torch.stack([kmeans_dev(torch, R_[:, j * S_:(j + 1) * S_].contiguous(), wl["KS"], wl["ITERS"], 100 + j) for j in range(wl["M"])])
        log(f"[bench] trained config-4 quantizers on the device in {time.time() - t0:.1f}s")
        return dict(Cq=Cq.cpu().numpy(), P=P.cpu().numpy())

    key = f"q4_d{D}_m{wl['M']}_nl{wl['NLIST']}_nt{wl['NTRAIN']}_it{wl['ITERS']}.npz"
    if world > 1:  # rank 0 trains, everybody loads the same file
        if rank == 0:
            cached(key, make_q)
        ctx.barrier()
    z = cached(key, make_q)
    Cq, P = z["Cq"], z["P"]
    Q = np.concatenate([x.cpu().numpy() for _, x in gen_on_device(torch, ctx.dev, dce, NQ, synth.SEED_Q)])
    dQ = torch.from_numpy(Q).to(ctx.dev)

    def build(S):
        t0 = time.time()
        mi = MultiIVFPQ(D, N_DB, wl["M"], wl["KS"], M.TransformationType.None_, wl["NLIST"], list_shards=S)
        mi.loadCoarseQuantizer(Cq)
        mi.loadProductQuantizer(P)
        mi.setW(W)
        return mi, t0

    # ---- replica layout (S = 1): every rank holds all 10M codes; also yields the codes for the sharded build ----
    mi1, t0 = build(1)
    d_lists = torch.empty(N_DB, dtype=torch.int32, device=ctx.dev)
    d_codes = torch.empty((N_DB, wl["M"]), dtype=torch.uint8, device=ctx.dev)
    keep_x = {}
    t_add = 0.0
    for b, x in gen_on_device(torch, ctx.dev, dce, N_DB, synth.SEED_DB):
        torch.cuda.synchronize()
        ta = time.perf_counter()
        mi1.indexDev(x, d_lists[b:b + x.shape[0]], d_codes[b:b + x.shape[0]])
        t_add += time.perf_counter() - ta
        if b == 0:
            keep_x[0] = x[:20_000].cpu().numpy()
    torch.cuda.synchronize()
    index_rate = N_DB / t_add  # IVFPQ.indexVectorInternal (coarse assign + residual + PQ encode + append), vectors already in HBM
    lists, codes = d_lists.cpu().numpy(), d_codes.cpu().numpy()
    del d_lists, d_codes
    # every rank generated and encoded the database itself: the copies must be identical
    chk = torch.tensor([float(np.bitwise_xor.reduce(codes.view(np.uint64).ravel()) % (1 << 52)), float(lists.astype(np.int64).sum())],
                       dtype=torch.float64, device=ctx.dev)
    lo, hi = chk.clone(), chk.clone()
    if ctx.dist is not None:
        ctx.dist.all_reduce(lo, op=ctx.dist.ReduceOp.MIN)
        ctx.dist.all_reduce(hi, op=ctx.dist.ReduceOp.MAX)
    same_db = bool((lo == hi).all().item())
    log(f"[bench] rank {rank}: indexed {N_DB} vectors (device-generated) in {time.time() - t0:.1f}s, {t_add:.2f}s of it in mmidx_add_dev ({index_rate / 1e6:.1f} M vectors/s)")
    per = (NQ + world - 1) // world
    mi1.connect(max_gq=max(per, NQ if world == 1 else per), k_max=K)
    results = {}

    def measure(mi, dq, name, queries_per_step):
        st = {}

        def step():
            st["res"] = mi.search(K, dq, gather_all=True)[:3]

        mi.index.enableTimings(True)
        stg = []
        s = summarize(ctx.timed(step, args.steps, args.warmup if args.profile else 6, after=lambda: stg.append(list(mi.index.lastTimingsMulti().values())),
                                min_s=0.0 if args.profile else MIN_TIMED_S), queries_per_step)
        mi.index.enableTimings(False)
        s["stage_ms_per_step"] = dict(zip(MULTI_STAGES, np.mean(np.asarray(stg), axis=0).tolist()))
        s["launches_per_step"] = mi.index.lastLaunches()
        results[name] = (s, st["res"][0].cpu().numpy(), st["res"][1].cpu().numpy())
        return s

    Qp = np.concatenate([Q, np.repeat(Q[:1], per * world - NQ, axis=0)])
    dq1 = torch.from_numpy(np.ascontiguousarray(Qp[rank * per:(rank + 1) * per])).to(ctx.dev)
    clocks = ClockSampler(ctx.dev.index, enabled=(rank == 0))
    clocks.start()
    rep = measure(mi1, dq1, "replicas", NQ)
    rep.update(S=1, R=world)
    if args.profile:  # under ncu: the replica-layout step only
        if rank == 0:
            print(json.dumps({"profile_run": True, "config": 4, **{k: rep[k] for k in ("value", "ms_per_step", "stage_ms_per_step")}}), flush=True)
        clocks.stop()
        ctx.close()
        return
    scan_bytes_total = None
    if rank == 0:
        ls = np.bincount(lists, minlength=wl["NLIST"])
        probes = mi1.index.computeNearestCoarseIndices(Q)
        scan_bytes_total = int(ls[probes].sum()) * (wl["M"] + 4)
    # ---- list sharding (S = world) ----
    shd = None
    if world > 1:
        ms, t0 = build(world)
        ms.setShardMap(balanced_shard_map(np.bincount(lists, minlength=wl["NLIST"]), world))
        for b in range(0, N_DB, 2_000_000):  # straight through the C ABI: no host-side id map for 10M entries
            lb, cb = np.ascontiguousarray(lists[b:b + 2_000_000]), np.ascontiguousarray(codes[b:b + 2_000_000])
            M._capi.check(lib.mmidx_add_codes(ms.index._h, lb.shape[0], C.c_void_p(lb.ctypes.data), C.c_void_p(cb.ctypes.data)))
        ms.connect(max_gq=NQ, k_max=K)
        log(f"[bench] rank {rank}: list-sharded index ({int(ms.listSizes().sum())} vectors here) in {time.time() - t0:.1f}s")
        shd = measure(ms, dQ, "sharded", NQ)
        shd.update(S=world, R=1, vectors_on_rank0=int(ms.listSizes().sum()))
        hQ = torch.from_numpy(Q).pin_memory()
        sl = (NQ + world - 1) // world
        hi, hd, hc = (torch.empty((sl, K), dtype=torch.int32).pin_memory(), torch.empty((sl, K), dtype=torch.float64).pin_memory(),
                      torch.empty(sl, dtype=torch.int32).pin_memory())
        e2e = summarize(ctx.timed(lambda: ms.search_host(K, hQ, hi, hd, hc), args.steps, 3, wall=True), NQ)
        e2e.update(h2d_bytes_per_step=hQ.numel() * 8 * world, d2h_bytes_per_step=(sl * K * 12 + sl * 4) * world)
    else:
        hQ = torch.from_numpy(Q).pin_memory()
        hi, hd, hc = (torch.empty((NQ, K), dtype=torch.int32).pin_memory(), torch.empty((NQ, K), dtype=torch.float64).pin_memory(),
                      torch.empty(NQ, dtype=torch.int32).pin_memory())
        e2e = summarize(ctx.timed(lambda: M._capi.check(lib.mmidx_search(mi1.index._h, NQ, ptr(hQ), K, ptr(hi), ptr(hd), ptr(hc))),
                                  args.steps, 3, wall=True), NQ)
        e2e.update(h2d_bytes_per_step=hQ.numel() * 8, d2h_bytes_per_step=NQ * K * 12 + NQ * 4)
    clk = clocks.stop()
    if rank != 0:
        ctx.close()
        return
    head_name = "sharded" if world > 1 else "replicas"
    head, h_iids, h_dist = results[head_name]
    parity = {"database_identical_on_all_ranks": same_db}
    if not args.no_cpu_baseline:
        O = oracle()
        off, cc, ii = synth.csr_from_assignments(lists, codes, wl["NLIST"])
        ns = 256
        oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q[:ns], K, W, threads=O.num_threads())
        for name, (_, ri, rd) in results.items():
            parity[f"{name}_first_{ns}_equal_to_oracle"] = bool((ri[:ns] == oi).all() and (rd[:ns] == od).all())
        ne = 5000
        ol, ocodes = O.ivfpq_encode(Cq, P, keep_x[0][:ne], threads=O.num_threads())
        parity[f"codes_first_{ne}_equal_to_oracle"] = bool((ol == lists[:ne]).all() and (np.asarray(ocodes) == codes[:ne]).all())
    if world > 1:
        si, sd = results["sharded"][1][:NQ], results["sharded"][2][:NQ]
        ri, rd = results["replicas"][1][:NQ], results["replicas"][2][:NQ]
        parity["sharded_equals_replicas_all_rows"] = bool((si == ri).all() and (sd == rd).all())
        bad = np.nonzero((si != ri).any(axis=1) | (sd != rd).any(axis=1))[0]
        if len(bad):  # diagnostics for a failing run
            q = int(bad[0])
            parity["mismatching_queries"] = int(len(bad))
            parity["first_mismatch"] = {"query": q, "same_id_set": bool(set(si[q]) == set(ri[q])), "same_dist_multiset": bool((np.sort(sd[q]) == np.sort(rd[q])).all()),
                                        "first_col": int(np.nonzero((si[q] != ri[q]) | (sd[q] != rd[q]))[0][0]),
                                        "sharded": [int(x) for x in si[q][:6]], "replicas": [int(x) for x in ri[q][:6]]}
    peak, peak_src = peaks()
    scan_ms = head["stage_ms_per_step"]["scan"]
    per_gpu_bytes = scan_bytes_total / world if world > 1 else scan_bytes_total
    achieved = per_gpu_bytes / (scan_ms * 1e-3) / 1e9
    traffic4 = None  # DRAM bytes of one launch of the scan on ONE GPU holding the whole 10 M database (ncu, profiles/)
    if world == 1 and N_DB == wl["N_DB"]:
        try:
            traffic4 = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic_cfg4.json")))["dram_bytes_per_step"]
        except Exception:
            traffic4 = None
    out = {
        "metric": "queries/sec", "value": head["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": 6, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (generated on the device)",
        "config": {"workload": wl["name"] if N_DB == wl["N_DB"] else wl["name"] + f" [reduced to {N_DB} vectors]", "nq_per_step": NQ, "k": K},
        "layout": f"{world} list shards x 1 group (per-shard queues stored into the slice owners' windows, device merge)" if world > 1 else "1 GPU",
        "median_ms_per_step": head["median_ms_per_step"], "timed_steps": head["timed_steps"], "timed_region_s": head["timed_region_s"],
        "step_ms_percentiles": head.get("step_ms_percentiles"),
        "stage_ms_per_step": head["stage_ms_per_step"],
        "roofline": {"bound": "hbm", "kernel": "k_ivfpq_scan_fast<2048,16>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic4, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": per_gpu_bytes, "kernel_ms_per_launch": scan_ms,
                     "note": "rank 0's launch; bytes = the job's probed-list bytes / shards (balanced map)"},
        "list_sharded": shd, "replicas": rep, "index_vectors_per_s": index_rate,
        "e2e": {"value": e2e["value"], "unit": "queries/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                "d2h_bytes_per_step": e2e["d2h_bytes_per_step"], "ms_per_step": e2e["ms_per_step"]},
        "gpu_launches": head["launches_per_step"] * head["timed_steps"], "clocks": clk, "parity": parity, "cpu_baseline": None,
    }
    print(json.dumps(out), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[3, 4], help="3: BASELINE configs[2] (default, the metric's config); 4: configs[3]")
    ap.add_argument("--n-db", type=int, default=0, help="config 4 only: reduced database size for quick runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline numbers only: no extra layouts, rows, latency table")
    ap.add_argument("--profile", action="store_true", help="for runs under ncu: exact warm-up count, no e2e / CPU legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
