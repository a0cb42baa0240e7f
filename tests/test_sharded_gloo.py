"""World-size-2 CPU test (gloo) of the multi-GPU host logic: the list-sharding rule, the packed all-gather
layout of multimedia_indexing_b200.sharded, and the cross-shard merge order.  Each rank plays a shard with the
CPU oracle standing in for the kernels (tests may use the oracle), then the merged result must equal the
unsharded oracle (the reference keeps ONE queue for all probed lists, IVFPQ.java:409,445)."""
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
        # pack exactly as ShardedIVFPQ.search_dev does and all-gather
        offs, sizes = sharded.packed_layout(nq, k)
        local = torch.zeros(offs[-1], dtype=torch.uint8)
        fields = dict(iids=li, dist=ld, seq=seq, tie=np.full(nq, -1.0), cnt=lc)
        for i, (name, dt, _) in enumerate(sharded.PACK_FIELDS):
            raw = torch.from_numpy(np.ascontiguousarray(fields[name])).to(dt).contiguous().view(torch.uint8).view(-1)
            local[offs[i]:offs[i] + sizes[i]] = raw
        parts = sharded.gather_partials(local, world, nq, k)
        # merge in BoundedPriorityQueue order: ascending distance, later-offered (larger seq) first among ties
        oi, od, oc = O.ivfpq_search(Cq, P, off, cc, ii, Q, k, w)
        ok = True
        for r in range(nq):
            ent = []
            for s in range(world):
                for c in range(int(parts["cnt"][s, r])):
                    ent.append((float(parts["dist"][s, r, c]), -int(parts["seq"][s, r, c]), int(parts["iids"][s, r, c])))
            ent.sort()
            ent = ent[:k]
            ok &= [e[2] for e in ent] == oi[r, :oc[r]].tolist() and [e[0] for e in ent] == od[r, :oc[r]].tolist()
            ok &= len(ent) == oc[r]
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
