"""ORACLE (test infrastructure) — what the reference's retrieval becomes on tables of more than 1 M rows.

tools/build_rag_database.py:51-52 builds `table.create_index(metric='dot', vector_column_name='text_embedding')`
iff `len(table) > 1_000_000`, and every search passes `nprobes=50, refine_factor=30` (src/data/rag.py:37,54).
LanceDB 0.14's defaults make that an IVF-PQ index (256 partitions, 96 sub-vectors of 8 dims, 8-bit codes), so the
reference's answers on the 10 M-row config are APPROXIMATE: probe the `nprobes` partitions whose centroids score
best, rank their rows by the PQ-approximated dot product (asymmetric distance computation: exact query against
quantised row), take the best `k * refine_factor`, re-score those exactly, return the best k.

LanceDB itself is not installable here (parity unpinned, see flat_search.py), and its k-means seeds are its own, so
index-identical answers cannot be the bar for this regime. This module restates the published algorithm in numpy so
that the comparison the product IS held to can be defined and tested (oracle/compare.py::check_recall):
the exact scan must DOMINATE the approximate answer rank by rank, agree on the exact distance of every row both
return, and the approximate answer's recall against it is reported.
"""
from __future__ import annotations

import numpy as np


def _kmeans(x: np.ndarray, k: int, iters: int, rng, spherical: bool = False) -> np.ndarray:
    """Plain Lloyd iterations; returns [k, d] centroids (empty clusters re-seeded from random points)."""
    c = x[rng.choice(len(x), size=k, replace=len(x) < k)].copy()
    for _ in range(iters):
        d = (x * x).sum(-1)[:, None] - 2.0 * (x @ c.T) + (c * c).sum(-1)[None]
        a = d.argmin(-1)
        for j in range(k):
            m = a == j
            c[j] = x[m].mean(0) if m.any() else x[rng.integers(len(x))]
        if spherical:
            c /= np.maximum(np.linalg.norm(c, axis=-1, keepdims=True), 1e-12)
    return c.astype(np.float32)


class IvfPqIndex:
    """IVF-PQ over fp32 rows with the dot metric (`_distance = 1 - q.d`, as flat_search's "dot")."""

    def __init__(self, rows: np.ndarray, num_partitions: int = 256, num_sub_vectors: int = 96, num_bits: int = 8,
                 sample: int = 8192, iters: int = 8, seed: int = 0):
        rows = np.asarray(rows, dtype=np.float32)
        n, dim = rows.shape
        if dim % num_sub_vectors:
            raise ValueError("dim must be a multiple of num_sub_vectors")
        rng = np.random.default_rng(seed)
        train = rows[rng.choice(n, size=min(sample, n), replace=False)]
        self.rows, self.m, self.ds = rows, num_sub_vectors, dim // num_sub_vectors
        self.centroids = _kmeans(train, num_partitions, iters, rng)
        d = (rows * rows).sum(-1)[:, None] - 2.0 * (rows @ self.centroids.T) + (self.centroids ** 2).sum(-1)[None]
        self.part = d.argmin(-1)
        # product quantiser: one codebook of 2^bits centroids per sub-vector, trained on the sample
        kc = 1 << num_bits
        tr = train.reshape(len(train), self.m, self.ds)
        self.codebooks = np.stack([_kmeans(np.ascontiguousarray(tr[:, j]), kc, iters, rng) for j in range(self.m)])
        sub = rows.reshape(n, self.m, self.ds)
        self.codes = np.empty((n, self.m), dtype=np.uint16)
        for j in range(self.m):
            cb = self.codebooks[j]
            dj = (sub[:, j] ** 2).sum(-1)[:, None] - 2.0 * (sub[:, j] @ cb.T) + (cb * cb).sum(-1)[None]
            self.codes[:, j] = dj.argmin(-1)
        self.lists = [np.nonzero(self.part == p)[0] for p in range(num_partitions)]

    def search(self, q: np.ndarray, k: int, nprobes: int = 50, refine_factor: int | None = 30):
        """-> (distance f32 [<=k], row i64 [<=k]) ascending `_distance = 1 - q.d`, ties by row."""
        q = np.asarray(q, dtype=np.float32)
        probe = np.argsort(-(self.centroids @ q), kind="stable")[:nprobes]
        cand = np.concatenate([self.lists[p] for p in probe]) if len(probe) else np.empty(0, dtype=np.int64)
        if cand.size == 0:
            return np.empty(0, np.float32), np.empty(0, np.int64)
        lut = np.einsum("mcd,md->mc", self.codebooks, q.reshape(self.m, self.ds))       # [m, 2^bits] partial dots
        approx = lut[np.arange(self.m)[None, :], self.codes[cand]].sum(-1)
        keep = k * refine_factor if refine_factor else k
        order = np.lexsort((cand, -approx))[:keep]
        cand = cand[order]
        if refine_factor:
            exact = 1.0 - (self.rows[cand].astype(np.float64) @ q.astype(np.float64))
            o = np.lexsort((cand, exact))[:k]
            return exact[o].astype(np.float32), cand[o].astype(np.int64)
        return (1.0 - approx[order][:k]).astype(np.float32), cand[:k].astype(np.int64)
