"""ORACLE (test infrastructure) — the timed CPU arm: the reference's retrieval call pattern
restated with torch fp32 on the host cores (all threads torch/MKL will use).

`search_loop` mirrors what the reference executes today: ONE flat scan per query, one call per
annotation (src/data/datamodule.py:257-262 -> src/data/rag.py:54), squared-L2 over fp32 rows,
top-k, post-filter, record list. `search_batched` is the best case a CPU engine could do
(one sgemm for the whole batch). Only bench.py's cpu_baseline / --impl reference legs and
tests/ may import this.
"""
from __future__ import annotations

import torch


def search_one(db: torch.Tensor, db_sq: torch.Tensor, q: torch.Tensor, k: int,
               row_group: torch.Tensor | None = None, exclude: int = -1):
    """db [N, D] fp32, db_sq [N] = |d|^2, q [D] -> (distance [<=k], index [<=k]) ascending."""
    d = (q @ q) + db_sq - 2.0 * (db @ q)
    dist, idx = torch.topk(d, min(k, d.numel()), largest=False, sorted=True)
    if exclude >= 0 and row_group is not None:
        keep = row_group[idx] != exclude
        dist, idx = dist[keep], idx[keep]
    return dist.clamp_min_(0), idx


def search_loop(db, db_sq, queries, k, row_group=None, exclude=None, videos=None):
    out = []
    for i in range(queries.shape[0]):
        dist, idx = search_one(db, db_sq, queries[i], k, row_group, -1 if exclude is None else int(exclude[i]))
        if videos is not None:   # the record conversion the reference pays per call (rag.py:29-30)
            out.append([{"video": videos[j], "_distance": float(x)} for x, j in zip(dist.tolist(), idx.tolist())])
        else:
            out.append((dist, idx))
    return out


def search_batched(db, db_sq, queries, k, chunk: int = 256):
    outs_d, outs_i = [], []
    for s in range(0, queries.shape[0], chunk):
        q = queries[s:s + chunk]
        d = (q * q).sum(-1, keepdim=True) + db_sq[None] - 2.0 * (q @ db.T)
        dist, idx = torch.topk(d, min(k, d.shape[1]), largest=False, sorted=True)
        outs_d.append(dist.clamp_min_(0))
        outs_i.append(idx)
    return torch.cat(outs_d), torch.cat(outs_i)
