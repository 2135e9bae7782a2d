"""Worker for tests/test_gpu_multi.py: run under torch.distributed.run with >= 2 ranks (NCCL).

Checks, on every rank: (1) row-sharded search == single-table oracle search of the same rows, for the
streaming and tensor paths, with and without the filters, over both transports (NCCL all-gather + merge
kernel; exchange fused into the last search kernel over peer memory), incl. the exactness margin of the
global result, batches beyond the single-phase limit, an empty shard, the host-buffer (graph) entry, the
reference-facing RAGDatabase over a sharded table with its fp32 re-check, and a peer that never shows up
(bounded wait -> error, then re-synchronisation); (2) the gather kernel reading a row-sharded feature
table through CUDA-IPC peer pointers == the restatement.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import motionrag_b200 as m  # noqa: E402
from motionrag_b200 import synthetic  # noqa: E402
from oracle import cama_context as cc  # noqa: E402
from oracle import compare, flat_search as fs  # noqa: E402


def check_retriever(retr, db, q, groups, excl, k, dev):
    for nq, path in ((1, "stream_f32"), (1, "auto"), (4, "stream_bf16"), (300, "tensor_bf16"), (64, "auto")):
        for filt in (None, "post", "pre"):
            ex = None if filt is None else torch.from_numpy(excl[:nq]).to(dev)
            r = retr.search(torch.from_numpy(q[:nq]).to(dev), k, path=path, exclude_group=ex, filter_mode=filt or "post",
                            certify=True)
            rd, ri = fs.flat_search(db, q[:nq], k, "l2", groups if filt else None, excl[:nq] if filt else None,
                                    prefilter=(filt == "pre"))
            rep = compare.check_retrieval(r.distance.cpu().numpy(), r.index.cpu().numpy(), rd, ri, db, q[:nq])
            assert rep["index_mismatches"] == rep["near_tie_positions"], (nq, path, filt, rep)
            gi = r.index.cpu().numpy()
            assert np.all(r.group.cpu().numpy()[gi >= 0] == groups[gi[gi >= 0]])
            mg = r.margin.cpu().numpy()
            assert mg.shape == (nq,) and not np.isnan(mg).any() and (mg > -1e-6).all(), (nq, path, filt, mg[:4])


def bf16_round(a):
    return torch.from_numpy(a).bfloat16().float().numpy()


def check_sharded_margin(retr, db, q, k, rps, world, dev):
    """margin of the GLOBAL result = (exact q.d of the k-th hit - max over shards of the scan score of that
    shard's weakest re-ranked row) / |q| — recomputed here from the definition (stream_bf16: 32 re-ranked)."""
    nq = 3
    r = retr.search(torch.from_numpy(q[:nq]).to(dev), k, path="stream_bf16", certify=True)
    scan = q[:nq].astype(np.float64) @ bf16_round(db).T.astype(np.float64)
    true = q[:nq].astype(np.float64) @ db.T.astype(np.float64)
    idx = r.index.cpu().numpy()
    for j in range(nq):
        weakest = max(np.sort(scan[j, g * rps:(g + 1) * rps])[::-1][31] for g in range(world)
                      if db[g * rps:(g + 1) * rps].shape[0] >= 32)
        want = (true[j, idx[j, k - 1]] - weakest) / np.linalg.norm(q[j])
        got = float(r.margin[j])
        assert abs(got - want) < 2e-4, (j, got, want)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, dim, k = 30_007, 768, 12
    db = synthetic.database(n, dim, "clustered", seed=5, device="cpu").numpy()
    rng = np.random.default_rng(1)
    src = rng.integers(0, n, 300)
    q = synthetic.queries_from_rows(torch.from_numpy(db[src]), seed=6).numpy()
    groups = (np.arange(n) // 3).astype(np.int32)
    excl = groups[src].astype(np.int32)
    rps, lo, hi = m.shard_range(n, world, rank)
    shard = m.EmbeddingStore(dim, hi - lo, dev)
    shard.append(db[lo:hi], normalise=False)
    shard.set_groups(groups[lo:hi])
    xchg = m.PeerExchange(rank, world, dev, nq_cap=4096, k_cap=32)
    for retr in (m.ShardedRetriever(shard, rank, world, rps),                    # NCCL all-gather + merge kernel
                 m.ShardedRetriever(shard, rank, world, rps, exchange=xchg)):     # exchange fused into the last kernel
        check_retriever(retr, db, q, groups, excl, k, dev)
    retr = m.ShardedRetriever(shard, rank, world, rps, exchange=xchg)
    check_sharded_margin(retr, db, q, k, rps, world, dev)
    for rep in range(40):    # slot reuse / epoch ordering under back-to-back calls (single-launch form)
        path = ("stream_f32", "auto")[rep % 2]
        nqr = 1 if rep % 2 else 3
        r = retr.search(torch.from_numpy(q[rep:rep + nqr]).to(dev), k, path=path)
        rd, ri = fs.flat_search(db, q[rep:rep + nqr], k)
        compare.check_retrieval(r.distance.cpu().numpy(), r.index.cpu().numpy(), rd, ri, db, q[rep:rep + nqr])
    # pipelined single-query searches: 300 launches enqueued without a host sync in between, so the scan of call
    # i+1 runs under the tail (select, re-rank, peer exchange) of call i on every rank; all results checked after
    qd = torch.from_numpy(q).to(dev)
    exd = torch.from_numpy(excl).to(dev)
    pending = [retr.search(qd[j:j + 1], k, exclude_group=exd[j:j + 1], certify=True) for j in range(300)]
    torch.cuda.synchronize()
    rd, ri = fs.flat_search(db, q[:300], k, "l2", groups, excl[:300])
    got_d = torch.cat([r.distance for r in pending]).cpu().numpy()
    got_i = torch.cat([r.index for r in pending]).cpu().numpy()
    repo = compare.check_retrieval(got_d, got_i, rd, ri, db, q[:300])
    assert repo["index_mismatches"] == repo["near_tie_positions"], repo
    assert not bool(torch.isnan(torch.cat([r.margin for r in pending])).any())
    del pending
    # every rank must hold the identical answer
    r = retr.search(torch.from_numpy(q[:64]).to(dev), k)
    ref = r.index.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, r.index)

    # a batch beyond the single-phase limit: publish in K3, wait + merge in a second kernel (4096 queries, 3 epochs)
    qbig = np.tile(q, (14, 1))[:4096] * np.linspace(0.5, 2.0, 4096, dtype=np.float32)[:, None]
    exbig = np.tile(excl, 14)[:4096]
    sample = np.r_[0:200, 1900:2100, 3896:4096]           # head, middle and tail of the batch against the oracle
    rd, ri = fs.flat_search(db, qbig[sample], k, "l2", groups, exbig[sample])
    for rep in range(3):
        r = retr.search(torch.from_numpy(qbig).to(dev), k, exclude_group=torch.from_numpy(exbig).to(dev), certify=True)
        repo = compare.check_retrieval(r.distance.cpu().numpy()[sample], r.index.cpu().numpy()[sample], rd, ri, db, qbig[sample])
        assert repo["index_mismatches"] == repo["near_tie_positions"]
        assert not bool(torch.isnan(r.margin).any()) and bool((r.index[:, 0] >= 0).all())
        ref = r.index.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, r.index)
    # soak: 100 more epochs of the two-phase exchange back to back (slot reuse, flags, epochs); the answer of every
    # epoch must be the first one's, bit for bit
    first_i, first_d = r.index.clone(), r.distance.clone()
    qbig_d, exbig_d = torch.from_numpy(qbig).to(dev), torch.from_numpy(exbig).to(dev)
    for rep in range(100):
        r = retr.search(qbig_d, k, exclude_group=exbig_d)
        assert torch.equal(r.index, first_i) and torch.equal(r.distance, first_d), rep
    shard.poll_error()

    # host buffers in / out: one captured graph per rank with the exchange inside; epochs keep advancing
    for rep in range(12):
        nqr = (1, 1, 3, 64)[rep % 4]
        d, i, g, mg = retr.search_host(q[rep:rep + nqr], k, exclude_group=excl[rep:rep + nqr], certify=True)
        rd, ri = fs.flat_search(db, q[rep:rep + nqr], k, "l2", groups, excl[rep:rep + nqr])
        repo = compare.check_retrieval(d, i, rd, ri, db, q[rep:rep + nqr])
        assert repo["index_mismatches"] == repo["near_tie_positions"] and not np.isnan(mg).any()

    # the reference-facing class over the sharded table: same records on every rank, == the oracle class
    cols = {"video": np.array([f"video_{j // 3}" for j in range(n)]), "start_sec": np.arange(n, dtype=np.float64)}
    rdb = m.RAGDatabase.from_store(shard, cols, retriever=retr)
    ora = fs.OracleRAGDatabase({**cols, "text_embedding": db})
    for j in range(6):
        kw = dict(top_k=k, where=f'video != "{cols["video"][src[j]]}"', select=["video", "start_sec"])
        got, want = rdb.text_search(q[j], **kw), ora.text_search(q[j], **kw)
        assert [r["video"] for r in got] == [r["video"] for r in want], j
        assert np.allclose([r["_distance"] for r in got], [r["_distance"] for r in want], rtol=1e-4)
    batch = rdb.search_batch(q[:200], top_k=k, where=[f'video != "{cols["video"][s_]}"' for s_ in src[:200]],
                             select=["video", "start_sec"])
    rd, ri = fs.flat_search(db, q[:200], k, "l2", groups, excl[:200])
    got_i = np.full((200, k), -1, dtype=np.int64)
    got_d = np.full((200, k), np.inf, dtype=np.float32)
    for j, rr in enumerate(batch):
        got_i[j, :len(rr)] = [int(r["start_sec"]) for r in rr]           # start_sec holds the row number here
        got_d[j, :len(rr)] = [r["_distance"] for r in rr]
        assert all(r["video"] == cols["video"][int(r["start_sec"])] for r in rr)
    repo = compare.check_retrieval(got_d, got_i, rd, ri, db, q[:200])
    assert repo["index_mismatches"] == repo["near_tie_positions"]
    assert rdb.fp32_rechecks == 0

    # uncertifiable queries on a sharded table are re-run on the fp32 rows of every shard
    rng2 = np.random.default_rng(7)
    n2 = 4000
    base = fs.normalise_rows(rng2.standard_normal((1, dim)).astype(np.float32))[0]
    db2 = fs.normalise_rows(rng2.standard_normal((n2, dim)).astype(np.float32))
    twins = rng2.choice(n2, 200, replace=False)
    db2[twins] = fs.normalise_rows(base[None] + 1e-2 / np.sqrt(dim) * rng2.standard_normal((200, dim)).astype(np.float32))
    rps2, lo2, hi2 = m.shard_range(n2, world, rank)
    shard2 = m.EmbeddingStore(dim, max(hi2 - lo2, 1), dev)
    shard2.append(db2[lo2:hi2], normalise=False)
    retr2 = m.ShardedRetriever(shard2, rank, world, rps2, exchange=xchg)
    rdb2 = m.RAGDatabase.from_store(shard2, {"video": np.array([f"v{j}" for j in range(n2)])}, retriever=retr2)
    plain_row = int(np.setdiff1d(np.arange(n2), twins)[5])            # an ordinary row: certified by its margin
    q2 = np.stack([base * 9, db2[plain_row] * 4]).astype(np.float32)
    for qq in (q2[:1], q2):
        res = rdb2.search_batch(qq, top_k=12, select=["video"])
        rd, ri = fs.flat_search(db2, qq, 12)
        got_i = np.array([[int(r["video"][1:]) for r in rr] for rr in res])
        compare.check_retrieval(np.array([[r["_distance"] for r in rr] for rr in res]), got_i, rd, ri, db2, qq)
        assert len(set(ri[0].tolist()) & set(got_i[0].tolist())) >= 10
    # (with many shards a shard may hold fewer twins than it re-ranks: then everything relevant WAS re-scored
    # in fp32 and the certificate rightly passes; the 16-candidate tensor path always has to re-run here)
    assert rdb2.fp32_rechecks == 2 if world <= 2 else rdb2.fp32_rechecks >= 1, rdb2.fp32_rechecks
    shard2.close()

    # an EMPTY last shard still takes part in the exchange (uneven partition)
    if world >= 2:
        n3 = 5000
        rps3 = -(-n3 // (world - 1))
        lo3, hi3 = min(n3, rank * rps3), min(n3, (rank + 1) * rps3)
        shard3 = m.EmbeddingStore(dim, max(hi3 - lo3, 1), dev)
        if hi3 > lo3:
            shard3.append(db[lo3:hi3], normalise=False)
        shard3.set_groups(groups[lo3:hi3])
        retr3 = m.ShardedRetriever(shard3, rank, world, rps3, exchange=xchg)
        assert (len(shard3) == 0) == (rank == world - 1)
        for nq3, path in ((1, "auto"), (3, "stream_f32"), (150, "auto")):
            r = retr3.search(torch.from_numpy(q[:nq3]).to(dev), k, path=path, exclude_group=torch.from_numpy(excl[:nq3]).to(dev),
                             certify=True)
            rd, ri = fs.flat_search(db[:n3], q[:nq3], k, "l2", groups[:n3], excl[:nq3])
            compare.check_retrieval(r.distance.cpu().numpy(), r.index.cpu().numpy(), rd, ri, db[:n3], q[:nq3])
        shard3.close()

    # a peer that never publishes: the wait is bounded, the error is readable, the ranks re-synchronise
    torch.cuda.synchronize()
    dist.barrier()
    if rank != 0:
        r = shard.search(torch.from_numpy(q[:2]).to(dev), k, path="stream_f32", index_base=rank * rps,
                         exchange=xchg.next(timeout_ms=300))
        torch.cuda.synchronize()
        assert bool((r.index == -1).all())
        try:
            shard.poll_error()
        except m.MragError as e:
            assert "timed out" in str(e)
        else:
            raise AssertionError("a missing peer must surface as an error")
        shard.poll_error()          # cleared by the first poll
    xchg.reset()
    r = retr.search(torch.from_numpy(q[:5]).to(dev), k, certify=True)
    rd, ri = fs.flat_search(db, q[:5], k)
    compare.check_retrieval(r.distance.cpu().numpy(), r.index.cpu().numpy(), rd, ri, db, q[:5])
    shard.poll_error()

    # row-sharded feature table read through peer pointers
    L, C, K = 25, 1024, 9
    rows_per = 96
    full = synthetic.features(rows_per * world, L, C, torch.bfloat16, seed=7, device="cpu")
    block = m.alloc_feature_block(rows_per, L, C, torch.bfloat16, dev)
    block.copy_(full[rank * rows_per:(rank + 1) * rows_per])
    torch.cuda.synchronize()
    table = m.open_peer_tables(m.FeatureTable(block, rows_per_shard=rows_per, shard_rank=rank, n_shards=world))
    assert table.complete
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, rows_per * world, (5, K), generator=g)
    idx[1, 3] = -1
    sos = (torch.randn(1, L, C, generator=g) / 32).bfloat16()
    un = torch.randn(L, C, generator=g).bfloat16()
    cond = torch.randn(5, (K + 1) * L, C, generator=g).bfloat16()
    ctx = m.MotionContext(table, sos, un, pe_max_length=256)
    x = ctx.build(idx.to(dev), cond.to(dev))
    want = cc.context_restatement(cc.gather_restatement(full, idx, un), sos, cc.sinusoid_table(256, C), cond.clone())
    assert torch.equal(x.cpu(), want)
    dist.barrier()
    if rank == 0:
        print("MGPU_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
