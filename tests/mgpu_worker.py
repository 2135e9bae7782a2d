"""Worker for tests/test_gpu_multi.py: run under torch.distributed.run with >= 2 ranks (NCCL).

Checks, on every rank: (1) row-sharded search over NCCL == single-table search of the same rows,
for the streaming and tensor paths, with and without the post-filter; (2) the gather kernel
reading a row-sharded feature table through CUDA-IPC peer pointers == the restatement.
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
    for nq, path in ((1, "stream_f32"), (4, "stream_bf16"), (300, "tensor_bf16"), (64, "auto")):
        for filt in (None, "post", "pre"):
            ex = None if filt is None else torch.from_numpy(excl[:nq]).to(dev)
            r = retr.search(torch.from_numpy(q[:nq]).to(dev), k, path=path, exclude_group=ex, filter_mode=filt or "post")
            rd, ri = fs.flat_search(db, q[:nq], k, "l2", groups if filt else None, excl[:nq] if filt else None,
                                    prefilter=(filt == "pre"))
            compare.check_retrieval(r.distance.cpu().numpy(), r.index.cpu().numpy(), rd, ri, db, q[:nq])
            gi = r.index.cpu().numpy()
            assert np.all(r.group.cpu().numpy()[gi >= 0] == groups[gi[gi >= 0]])


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
    xchg = m.PeerExchange(rank, world, dev, nq_cap=512, k_cap=32)
    for retr in (m.ShardedRetriever(shard, rank, world, rps),                    # NCCL all-gather + merge kernel
                 m.ShardedRetriever(shard, rank, world, rps, exchange=xchg)):     # exchange fused into K3 over peer memory
        check_retriever(retr, db, q, groups, excl, k, dev)
    retr = m.ShardedRetriever(shard, rank, world, rps, exchange=xchg)
    for rep in range(40):    # slot reuse / epoch ordering under back-to-back calls
        r = retr.search(torch.from_numpy(q[rep:rep + 3]).to(dev), k, path="stream_f32")
        rd, ri = fs.flat_search(db, q[rep:rep + 3], k)
        compare.check_retrieval(r.distance.cpu().numpy(), r.index.cpu().numpy(), rd, ri, db, q[rep:rep + 3])
    # every rank must hold the identical answer
    r = retr.search(torch.from_numpy(q[:64]).to(dev), k)
    ref = r.index.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, r.index)

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
