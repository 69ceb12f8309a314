"""World-size-2 (and 3, ragged) gloo tests of the sharding / gather logic on CPU: the path has no per-step
communication, so the only collective is the final all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dgdm_b200 import distributed as D


def test_shard_ranges_cover_and_balance():
    for n in (1, 5, 64, 1024, 1000):
        for w in (1, 2, 3, 4, 8):
            r = [D.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sz = [b - a for a, b in r]
            assert max(sz) - min(sz) <= 1 and sz == D.shard_sizes(n, w)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_obj, B, P, k, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(0)
        scores = torch.from_numpy(rs.randint(0, 5, size=(n_obj, B)).astype(np.float32))    # ties on purpose
        designs = torch.from_numpy(rs.randn(n_obj, B, P, 1).astype(np.float32))
        order = np.argsort(-scores.numpy(), axis=-1, kind="stable")[:, :k]
        lo, hi = D.shard_range(n_obj, world, rank)
        local = {"scores": scores[lo:hi], "designs": designs[lo:hi], "best_ids": torch.from_numpy(order[lo:hi].copy()),
                 "best_scores": torch.from_numpy(np.take_along_axis(scores.numpy(), order, 1)[lo:hi].copy())}
        g = D.gather_per_object_results(local, n_obj)
        ok = (torch.equal(g["scores"], scores) and torch.equal(g["designs"], designs)
              and g["best_ids"].numpy().tolist() == order.tolist())
        # multi-object mode: candidates sharded
        sc = scores[0]
        lo, hi = D.shard_range(B, world, rank)
        sel = lambda s, kk: (torch.from_numpy(np.argsort(-s.numpy(), axis=-1, kind="stable")[:, :kk].copy()),
                             torch.from_numpy(-np.sort(-s.numpy(), axis=-1, kind="stable")[:, :kk].copy()))
        m = D.gather_multi_object_results({"scores": sc[lo:hi], "designs": designs[0, lo:hi]}, B, k, sel)
        ok = ok and torch.equal(m["scores"], sc) and torch.equal(m["designs"], designs[0]) \
            and m["best_ids"].tolist() == np.argsort(-sc.numpy(), kind="stable")[:k].tolist()
        # per-object mode with fewer objects than ranks: candidates sharded, every rank holds all objects
        c = D.gather_per_object_candidate_shards({"scores": scores[:, lo:hi].contiguous(), "designs": designs[:, lo:hi].contiguous()},
                                                 B, k, sel)
        ok = ok and torch.equal(c["scores"], scores) and torch.equal(c["designs"], designs) \
            and c["best_ids"].numpy().tolist() == order.tolist()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_plan_per_object():
    assert D.plan_per_object(1024, 8) == "objects" and D.plan_per_object(8, 8) == "objects"
    assert D.plan_per_object(5, 8) == "candidates" and D.plan_per_object(1, 2) == "candidates"
    assert D.plan_per_object(5, 1) == "objects"


@pytest.mark.parametrize("world,n_obj", [(2, 6), (3, 7), (2, 1)])
def test_gather_results_gloo(world, n_obj):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_obj, 5, 14, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
