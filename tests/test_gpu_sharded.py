"""Multi-GPU parity (needs >= 2 B200: gpurun --gpus 2): the whole step inside libmmidx (mmidx_search_multi_dev) --
IVF lists sharded over the ranks, result rows stored straight into the peers' exchange windows over NVLink, epoch-flag
exchange points, device merge and the cross-shard ordered tie pass -- against the UNSHARDED CPU oracle
(the reference keeps one queue for all probed lists, IVFPQ.java:409,445)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _check(mi, O, synth, Cq, P, lists, codes, nlist, Q, k, w, gather, steps=1):
    """Q: the JOB-wide batch, split evenly over the R groups (every rank of a job steps with the same gq, so the last
    group's batch is padded with copies of query 0).  Compares with the unsharded oracle."""
    import torch.distributed as dist
    nq = Q.shape[0]
    per = (nq + mi.R - 1) // mi.R
    g0 = mi.group * per
    Qg = np.concatenate([Q[g0:g0 + per], np.repeat(Q[:1], max(0, g0 + per - nq), axis=0)])[:per]
    dQ = torch.from_numpy(np.ascontiguousarray(Qg)).cuda()
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w, threads=8)
    gqp = ((per + mi.S - 1) // mi.S) * mi.S
    ok = True
    for _ in range(steps):  # several steps: window parity, epochs, CUDA-graph capture on the third
        iids, dd, cnt, row0, nrows = mi.search(k, dQ, gather_all=gather)
        torch.cuda.synchronize()
        if gather:
            spans = [(g * gqp, g * per, min(nq, (g + 1) * per)) for g in range(mi.R)]
        else:  # only this rank's rows are defined
            a0 = g0 + (row0 - mi.group * gqp)
            spans = [(row0, a0, min(nq, a0 + nrows))]
        for r0, a0, a1 in spans:
            if a1 <= a0:
                continue
            ok &= bool((iids[r0:r0 + a1 - a0].cpu().numpy() == oi[a0:a1]).all())
            ok &= bool((dd[r0:r0 + a1 - a0].cpu().numpy() == od[a0:a1]).all())
            ok &= bool((cnt[r0:r0 + a1 - a0].cpu().numpy() == oc[a0:a1]).all())
        dist.barrier()
    return ok


def _worker(rank, world, port, q):
    try:
        import sys
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        torch.cuda.set_stream(torch.cuda.Stream())  # a capturable stream: steps 3+ of every case replay CUDA graphs
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (root, os.path.join(root, "oracle")):
            sys.path.insert(0, p)
        import mmidx_b200 as M
        import pyoracle as O
        from multimedia_indexing_b200 import synth
        from multimedia_indexing_b200.sharded import MultiIVFPQ

        ok = True
        msgs = []
        # ---- list sharding, S = world: plain data, then heavy duplication (exact ties cut ACROSS shards) ----
        for case in ("plain", "ties", "exact_kernels", "m16_short_lists"):
            d, m, ks, nlist, w, k, n, nq = 64, 8, 256, 64, 16, 100, 30000, 701  # odd nq: ragged last slice
            if case == "m16_short_lists":  # configs[3] geometry: m = 16, many probes over short (and some empty) lists
                d, m, nlist, w, n, nq = 128, 16, 512, 64, 60000, 1000
            ce = synth.mixture_centers(d, 128)
            if case == "ties":
                base = synth.mixture(60, d, 1, ce)
                X = base[np.random.default_rng(1).integers(0, 60, size=n)]
                Q = base[:40] + 1.0
                nq, k = 40, 25
            else:
                X, Q = synth.mixture(n, d, 1, ce), synth.mixture(nq, d, 2, ce)
            if case == "exact_kernels":
                ks, m = 64, 4  # outside the fused kernel's geometry: binary64 ADC-table kernels
            Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=5000, iters=3, centers=ce)
            mi = MultiIVFPQ(d, n, m, ks, M.TransformationType.None_, nlist, list_shards=world)
            mi.loadCoarseQuantizer(Cq)
            mi.loadProductQuantizer(P)
            mi.setW(w)
            lists, codes = mi.indexAll(X, balanced=(case != "plain"))
            mi.connect(max_gq=1024, k_max=128)
            same = _check(mi, O, synth, Cq, P, lists, codes, nlist, Q, k, w, gather=True, steps=7)
            same &= _check(mi, O, synth, Cq, P, lists, codes, nlist, Q[: nq // 2], k, w, gather=False, steps=2)
            ok &= bool(same)
            stored = int(mi.listSizes().sum())
            ok &= stored < n  # really sharded
            msgs.append(f"S={world} {case}: equal={bool(same)} local_vectors={stored}")
            # host-buffer entry point (e2e path): this rank's rows of the batch
            hi = torch.empty((1024, k), dtype=torch.int32).pin_memory()
            hd = torch.empty((1024, k), dtype=torch.float64).pin_memory()
            hc = torch.empty(1024, dtype=torch.int32).pin_memory()
            fq, nr = mi.search_host(k, np.ascontiguousarray(Q), hi, hd, hc)
            off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
            oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w, threads=8)
            sameh = (hi[:nr].numpy() == oi[fq:fq + nr]).all() and (hd[:nr].numpy() == od[fq:fq + nr]).all() and \
                (hc[:nr].numpy() == oc[fq:fq + nr]).all()
            ok &= bool(sameh)
            msgs.append(f"S={world} {case} host path: equal={bool(sameh)} rows={nr}")
            dist.barrier()
            mi.close()
        # ---- replica groups, S = 1 x R = world: every group searches its own slice, rows gathered everywhere ----
        d, m, ks, nlist, w, k, n, nq = 64, 8, 256, 64, 16, 100, 30000, 1501
        ce = synth.mixture_centers(d, 128)
        X, Q = synth.mixture(n, d, 1, ce), synth.mixture(nq, d, 2, ce)
        Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=5000, iters=3, centers=ce)
        mi = MultiIVFPQ(d, n, m, ks, M.TransformationType.None_, nlist, list_shards=1)
        mi.loadCoarseQuantizer(Cq)
        mi.loadProductQuantizer(P)
        mi.setW(w)
        lists, codes = mi.indexAll(X)
        mi.connect(max_gq=1024, k_max=100)
        same = _check(mi, O, synth, Cq, P, lists, codes, nlist, Q, k, w, gather=True, steps=7)
        ok &= bool(same)
        msgs.append(f"S=1 x R={world}: equal={bool(same)}")
        dist.barrier()
        mi.close()
        if world >= 4 and world % 2 == 0:  # hybrid: 2 list shards x world/2 groups
            mi = MultiIVFPQ(d, n, m, ks, M.TransformationType.None_, nlist, list_shards=2)
            mi.loadCoarseQuantizer(Cq)
            mi.loadProductQuantizer(P)
            mi.setW(w)
            lists, codes = mi.indexAll(X)
            mi.connect(max_gq=1024, k_max=100)
            same = _check(mi, O, synth, Cq, P, lists, codes, nlist, Q, k, w, gather=True, steps=7)
            same &= _check(mi, O, synth, Cq, P, lists, codes, nlist, Q, k, w, gather=False, steps=2)
            ok &= bool(same)
            msgs.append(f"S=2 x R={world // 2}: equal={bool(same)}")
            dist.barrier()
            mi.close()
        q.put((rank, ok, msgs))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, f"{e}\n{traceback.format_exc()}"))


def _run(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] is True for r in res), res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(1200)
def test_multi_ivfpq_two_gpus():
    _run(2)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
@pytest.mark.timeout(1200)
def test_multi_ivfpq_four_gpus_hybrid():
    """4 ranks: the S = 4 and S = 1 cases above plus the hybrid layout, 2 list shards x 2 groups"""
    _run(4)


@pytest.mark.timeout(600)
def test_multi_step_on_one_gpu():
    """world == 1 goes through the same entry points (window, epochs, graph capture) with no peers"""
    import sys
    import mmidx_b200 as M
    import pyoracle as O
    from multimedia_indexing_b200 import synth
    from multimedia_indexing_b200.sharded import MultiIVFPQ

    d, m, ks, nlist, w, k, n, nq = 32, 8, 256, 32, 8, 10, 8000, 300
    ce = synth.mixture_centers(d, 64)
    X, Q = synth.mixture(n, d, 1, ce), synth.mixture(nq, d, 2, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=4000, iters=3, centers=ce)
    mi = MultiIVFPQ(d, n, m, ks, M.TransformationType.None_, nlist, list_shards=1)
    mi.loadCoarseQuantizer(Cq)
    mi.loadProductQuantizer(P)
    mi.setW(w)
    lists, codes = mi.indexAll(X)
    mi.connect(max_gq=512, k_max=16)
    torch.cuda.set_stream(torch.cuda.Stream())  # a capturable stream (the legacy default stream is not): steps 3+ replay graphs
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w)
    dQ = torch.from_numpy(Q).cuda()
    for step in range(8):
        iids, dd, cnt, row0, nrows = mi.search(k, dQ, gather_all=True)
        torch.cuda.synchronize()
        assert (row0, nrows) == (0, nq)
        assert (iids[:nq].cpu().numpy() == oi).all() and (dd[:nq].cpu().numpy() == od).all() and (cnt[:nq].cpu().numpy() == oc).all(), step
    mi.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_one_process_one_thread_per_gpu(monkeypatch):
    """The embedding of a single JVM: ONE process, one host thread per GPU.  mmidx_comm_attach maps a same-process peer's window
    by plain peer access instead of CUDA IPC; the two threads then step through mmidx_search_multi (host buffers) together.
    Two list shards, result == the unsharded oracle."""
    import ctypes as C
    import threading

    import mmidx_b200 as M
    import pyoracle as O
    from multimedia_indexing_b200 import _capi, synth
    from multimedia_indexing_b200._capi import check, lib

    monkeypatch.setenv("MMIDX_COMM_TIMEOUT_S", "30")
    d, m, ks, nlist, w, k, n, nq, S = 32, 8, 256, 48, 12, 10, 12000, 333, 2
    ce = synth.mixture_centers(d, 64)
    X, Q = synth.mixture(n, d, 1, ce), synth.mixture(nq, d, 2, ce)
    Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=4000, iters=3, centers=ce)
    idx, handles = [], []
    for r in range(S):
        ix = M.IVFPQ(d, n, m, ks, M.TransformationType.None_, nlist, device=r, shard_rank=r, shard_count=S)
        ix.loadCoarseQuantizer(Cq)
        ix.loadProductQuantizer(P)
        ix.setW(w)
        lists, codes = ix.indexVectors(None, X, return_codes=True)  # every shard is offered every vector, keeps its lists
        h = (C.c_ubyte * _capi.COMM_HANDLE_BYTES)()
        check(lib.mmidx_comm_create(ix._h, r, S, S, 512, 16, h))
        idx.append(ix)
        handles.append(bytes(h))
    for ix in idx:
        check(lib.mmidx_comm_attach(ix._h, C.c_char_p(b"".join(handles))))
    off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
    oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w)
    sl = (nq + S - 1) // S
    out = [None] * S
    err = [None] * S

    def rank_thread(r):
        try:
            for step in range(5):  # window parity, epochs, graph capture on the third step
                ids = np.full((sl, k), -7, dtype=np.int32)
                dd = np.zeros((sl, k))
                cn = np.zeros(sl, dtype=np.int32)
                fq, nr = C.c_int64(), C.c_int64()
                check(lib.mmidx_search_multi(idx[r]._h, nq, C.c_void_p(Q.ctypes.data), k, C.c_void_p(ids.ctypes.data),
                                             C.c_void_p(dd.ctypes.data), C.c_void_p(cn.ctypes.data), C.byref(fq), C.byref(nr)))
                out[r] = (fq.value, nr.value, ids, dd, cn)
        except Exception as e:  # noqa: BLE001
            err[r] = repr(e)

    ts = [threading.Thread(target=rank_thread, args=(r,)) for r in range(S)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert err == [None] * S, err
    covered = 0
    for r in range(S):
        fq, nr, ids, dd, cn = out[r]
        assert (ids[:nr] == oi[fq:fq + nr]).all() and (dd[:nr] == od[fq:fq + nr]).all() and (cn[:nr] == oc[fq:fq + nr]).all(), r
        covered += nr
    assert covered == nq
    for ix in idx:
        check(lib.mmidx_comm_destroy(ix._h))
        ix.close()
