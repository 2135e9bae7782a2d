"""World-size-2 run of the row-sharded search plumbing on CPU (gloo): partition arithmetic, the
packed single all-gather, post-filter-after-global-merge staging. The oracle stands in for
the CUDA kernels here (test infrastructure only); the GPU tests cover the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flat_search as fs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _np_merge(dist_v, idx_v, grp_v, k_out, exclude_group, filter_mode, stride):
    from motionrag_b200.store import SearchResult
    G, nq, k = dist_v.shape
    od = torch.full((nq, k_out), float("inf"))
    oi = torch.full((nq, k_out), -1, dtype=torch.int64)
    og = torch.full((nq, k_out), -1, dtype=torch.int32)
    for q in range(nq):
        flat = [(float(dist_v[g, q, j]), g * k + j, int(idx_v[g, q, j]), int(grp_v[g, q, j]))
                for g in range(G) for j in range(k) if int(idx_v[g, q, j]) >= 0]
        flat.sort(key=lambda t: (t[0], t[1]))
        ex = -1 if exclude_group is None else int(exclude_group[q])
        if filter_mode == "post":
            flat = [t for t in flat[:k_out] if ex < 0 or t[3] != ex]
        elif filter_mode == "pre":
            flat = [t for t in flat if ex < 0 or t[3] != ex]
        for j, t in enumerate(flat[:k_out]):
            od[q, j], oi[q, j], og[q, j] = t[0], t[2], t[3]
    return SearchResult(od, oi, og)


def _worker(rank, world, port, db, q, groups, excl, k, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from motionrag_b200.parallel import ShardedRetriever, shard_range
    rps, lo, hi = shard_range(db.shape[0], world, rank)
    shard, shard_groups = db[lo:hi], groups[lo:hi]

    def local_search(queries, k_, metric, path, refine, exclude_group, filter_mode, index_base, out):
        ex = None if exclude_group is None else exclude_group.numpy()
        d, i = fs.flat_search(shard, queries.numpy(), k_, metric, shard_groups, ex, prefilter=(filter_mode == "pre"))
        out.distance.copy_(torch.from_numpy(d))
        out.index.copy_(torch.from_numpy(np.where(i >= 0, i + index_base, -1)))
        out.group.copy_(torch.from_numpy(np.where(i >= 0, shard_groups[np.clip(i, 0, None)], -1).astype(np.int32)))
        assert index_base == lo
        return out

    sr = ShardedRetriever(None, rank, world, rps, local_search=local_search, merge=_np_merge,
                          device=torch.device("cpu"))
    out = {}
    for mode in ("none", "post", "pre"):
        r = sr.search(torch.from_numpy(q), k, exclude_group=None if mode == "none" else torch.from_numpy(excl),
                      filter_mode=mode if mode != "none" else "post")
        out[mode] = (r.distance.numpy().copy(), r.index.numpy().copy())
    results[rank] = out
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharded_search_equals_single_table():
    rng = np.random.default_rng(0)
    n, dim, nq, k = 1001, 64, 9, 12
    db = fs.normalise_rows(rng.standard_normal((n, dim)).astype(np.float32))
    db[700] = db[20]                                  # cross-shard exact tie
    src = rng.integers(0, n, nq)
    src[0] = 20
    q = (db[src] * rng.uniform(5, 15, (nq, 1))).astype(np.float32)
    groups = (np.arange(n) // 3).astype(np.int32)
    excl = groups[src].astype(np.int32)
    excl[3] = -1
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), db, q, groups, excl, k, results), nprocs=2, join=True)
    assert set(results.keys()) == {0, 1}
    want = {"none": fs.flat_search(db, q, k), "post": fs.flat_search(db, q, k, "l2", groups, excl),
            "pre": fs.flat_search(db, q, k, "l2", groups, excl, prefilter=True)}
    for mode, (wd, wi) in want.items():
        for rank in (0, 1):
            gd, gi = results[rank][mode]
            np.testing.assert_array_equal(gi, wi, err_msg=f"{mode} rank {rank}")
            np.testing.assert_allclose(gd, wd, rtol=1e-6)
    assert results[0]["none"][1][0, :2].tolist() == [20, 700]
