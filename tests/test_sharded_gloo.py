"""World-size-2 CPU test (gloo) of the multi-GPU host logic: the list-sharding rules and the row routing of
multimedia_indexing_b200.sharded (the host mirror of csrc/comm.cuh: which exchange window a shard's partial queue of
query q is stored into, and where) plus the cross-shard merge order.  Each rank plays a list shard with the CPU oracle
standing in for the kernels (tests may use the oracle) and gloo standing in for the NVLink stores; the merged result
must equal the unsharded oracle (the reference keeps ONE queue for all probed lists, IVFPQ.java:409,445)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (root, os.path.join(root, "oracle")):
            sys.path.insert(0, p)
        import mmidx_b200  # noqa: F401
        import pyoracle as O
        from multimedia_indexing_b200 import sharded, synth

        d, m, ks, nlist, w, k, n, nq = 16, 4, 32, 12, 5, 10, 3000, 24
        ce = synth.mixture_centers(d, 32)
        X, Q = synth.mixture(n, d, 1, ce) + np.random.default_rng(9).normal(0, 0.01, (n, d)), synth.mixture(nq, d, 2, ce)
        Cq, P = synth.train_ivfpq(d, m, ks, nlist, ntrain=2000, iters=3, centers=ce)
        lists, codes = O.ivfpq_encode(Cq, P, X)
        off, cc, ii = synth.csr_from_assignments(lists, codes, nlist)
        # this rank's shard: same CSR with the foreign lists emptied (iids stay global)
        own = sharded.owner_of_list(np.arange(nlist), world) == rank
        keep = own[lists[ii]]
        soff = np.zeros(nlist + 1, np.int64)
        np.cumsum(np.where(own, np.diff(off), 0), out=soff[1:])
        li, ld, lc = O.ivfpq_search(Cq, P, soff, cc[keep], ii[keep], Q, k, w)
        # offer sequence numbers: probe rank * 2^32 + position in list
        probes = O.coarse_topw(Cq, Q, w)
        pos_in_list = np.empty(n, np.int64)
        for l in range(nlist):
            pos_in_list[ii[off[l]:off[l + 1]]] = np.arange(off[l + 1] - off[l])
        seq = np.zeros((nq, k), np.int64)
        for r in range(nq):
            rank_of = {int(l): p for p, l in enumerate(probes[r])}
            for c in range(lc[r]):
                seq[r, c] = (rank_of[int(lists[li[r, c]])] << 32) + pos_in_list[li[r, c]]
        # "store" every partial row into the window of the slice owner, at the row route_row() names (PeerSink ROUTE)
        S, sl = world, sharded.slice_len(nq, world)
        sendbuf = [torch.zeros((sl, 3 * k + 1), dtype=torch.float64) for _ in range(S)]  # iid | dist | seq | cnt
        for qq in range(nq):
            owner, row = sharded.route_row(qq, rank, nq, S)
            assert row == rank * sl + (qq - owner * sl) and 0 <= owner < S
            r = sendbuf[owner][row - rank * sl]
            r[0:k] = torch.from_numpy(li[qq].astype(np.float64))
            r[k:2 * k] = torch.from_numpy(ld[qq])
            r[2 * k:3 * k] = torch.from_numpy(seq[qq].astype(np.float64))  # < 2^53: exact
            r[3 * k] = float(lc[qq])
        recv = [torch.zeros((sl, 3 * k + 1), dtype=torch.float64) for _ in range(S)]
        works = [dist.isend(sendbuf[t], t) for t in range(S) if t != rank] + []
        recv[rank] = sendbuf[rank]
        for t in range(S):
            if t != rank:
                dist.recv(recv[t], t)
        for wk in works:
            wk.wait()
        window = torch.cat(recv)  # [S][sl] rows: shard t's partial of slice query ql at row t*sl + ql
        # merge my slice in BoundedPriorityQueue order: ascending distance, later-offered (larger seq) first among ties
        oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w)
        ok = True
        fin = torch.full((sl, 2 * k + 1), -1.0, dtype=torch.float64)
        for ql in range(sl):
            qq = rank * sl + ql
            if qq >= nq:
                break
            ent = []
            for t in range(S):
                r = window[t * sl + ql]
                for c in range(int(r[3 * k])):
                    ent.append((float(r[k + c]), -int(r[2 * k + c]), int(r[c])))
            ent.sort()
            ent = ent[:k]
            ok &= [e[2] for e in ent] == oi[qq, :oc[qq]].tolist() and [e[0] for e in ent] == od[qq, :oc[qq]].tolist()
            ok &= len(ent) == oc[qq]
            fin[ql, :len(ent)] = torch.tensor([float(e[2]) for e in ent], dtype=torch.float64)
            fin[ql, k:k + len(ent)] = torch.tensor([e[0] for e in ent], dtype=torch.float64)
            fin[ql, 2 * k] = len(ent)
        # final rows land at final_row() of the job-wide arrays on every rank (PeerSink BCAST)
        allfin = [torch.zeros_like(fin) for _ in range(S)]
        dist.all_gather(allfin, fin)
        job = torch.cat(allfin)
        for qq in range(nq):
            row = sharded.final_row(qq, 0, nq, S)
            n_ = int(job[row, 2 * k])
            ok &= n_ == oc[qq] and job[row, :n_].to(torch.int64).tolist() == oi[qq, :n_].tolist()
        # the balanced map is a valid ownership too
        bm = sharded.balanced_shard_map(np.bincount(lists, minlength=nlist), world)
        ok &= bm.min() >= 0 and bm.max() < world and len(bm) == nlist
        # every list is owned by exactly one rank
        owners = torch.from_numpy(own.astype(np.int32))
        dist.all_reduce(owners)
        ok &= bool((owners == 1).all())
        q.put((rank, bool(ok)))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, f"{e}\n{traceback.format_exc()}"))


@pytest.mark.timeout(300)
def test_sharded_merge_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res
