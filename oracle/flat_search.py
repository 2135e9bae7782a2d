"""ORACLE (test infrastructure, never shipped or measured as product) — CPU restatement of the
reference's retrieval: LanceDB 0.14 flat (un-indexed) vector search as driven by
`RAGDatabase.vector_search` (reference src/data/rag.py:36-61) from
`VideoDataModule.prepare_annotations` (src/data/datamodule.py:231-236, 257-265).

PARITY UNPINNED: the arithmetic lives in the un-vendored third-party wheel `lancedb==0.14.0`
(reference requirements.txt:18; Rust `lance` flat KNN), which is not installed here and
cannot be fetched, and the reference's own tests (tests/test_read_video.py) never touch
retrieval, so no golden vector of the reference pins this file. It restates LanceDB's
published semantics and is cross-checked in tests/ against two independent implementations
(scikit-learn brute-force kNN and a pure-Python float64 loop).

Restated semantics (each switchable where the LanceDB default could not be verified):
  * no index  -> exact scan of every row; `nprobes` / `refine_factor` are ignored
    (rag.py:54 passes them unconditionally; they only matter for IVF-PQ tables).
  * metric: LanceDB default is squared L2, `_distance = sum((q - d)^2)`; "cosine" gives
    `1 - cos(q, d)`; "dot" gives `1 - q.d`.  rag.py:54 never sets one -> "l2".
  * results come back in ascending `_distance`; at most `top_k` rows (`.limit(k)`).
  * `.where(expr)` without `prefilter=True` is a POST-filter: the k nearest rows are found
    first and the rows failing the predicate are dropped, so fewer than k rows can come
    back — which is why datamodule.py:234 asks for `ref_video_num + 3` and dataset.py:296
    slices `[:ref_video_num]`.  `prefilter=True` is available for completeness.
  * tie order is unspecified in LanceDB; this oracle (and the product) define it as lowest
    row index first.
  * database vectors are fp32 and L2-normalised (tools/build_rag_database.py:31-37), query
    vectors are fp32 and NOT normalised (datamodule.py:300-302).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import re

import numpy as np

METRICS = ("l2", "cosine", "dot")


def normalise_rows(x: np.ndarray) -> np.ndarray:
    """fp32 x / max(|x|, 1e-12) per row — how the store normalises on upload."""
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(-1, keepdims=True)).astype(np.float32)
    return (x / np.maximum(n, np.float32(1e-12))).astype(np.float32)


def distances(db: np.ndarray, queries: np.ndarray, metric: str = "l2",
              accumulate: str = "f64") -> np.ndarray:
    """[Q, N] `_distance` matrix, returned as float32.

    accumulate="f64": products and sums in float64, one final rounding to fp32 (the tightest
    fp32 reference; default). accumulate="f32": fp32 BLAS, as a CPU engine would run it.
    """
    if metric not in METRICS:
        raise ValueError(f"Invalid metric: {metric}")
    wt = np.float64 if accumulate == "f64" else np.float32
    d = np.asarray(db, dtype=np.float32).astype(wt)
    q = np.asarray(queries, dtype=np.float32).astype(wt)
    dot = q @ d.T
    if metric == "dot":
        out = 1.0 - dot
    else:
        qq = (q * q).sum(-1)[:, None]
        dd = (d * d).sum(-1)[None, :]
        if metric == "l2":
            out = np.maximum(qq + dd - 2.0 * dot, 0.0)
        else:
            out = 1.0 - dot / np.maximum(np.sqrt(qq) * np.sqrt(dd), 1e-30)
    return out.astype(np.float32)


def distances_direct_l2(db: np.ndarray, queries: np.ndarray) -> np.ndarray:
    """sum((q - d)^2) evaluated literally in float64 (small inputs; no expansion)."""
    d = np.asarray(db, dtype=np.float64)
    q = np.asarray(queries, dtype=np.float64)
    return ((q[:, None, :] - d[None, :, :]) ** 2).sum(-1).astype(np.float32)


def topk_rows(dist_row: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k smallest distances, ascending, ties -> lowest index."""
    n = dist_row.shape[0]
    k = min(k, n)
    if k <= 0:
        return np.empty(0, dtype=np.int64)
    if n > 4 * k:
        kth = np.partition(dist_row, k - 1)[k - 1]
        cand = np.nonzero(dist_row <= kth)[0]           # ascending index
    else:
        cand = np.arange(n)
    order = np.argsort(dist_row[cand], kind="stable")   # stable: equal distances keep index order
    return cand[order][:k].astype(np.int64)


def flat_search(db: np.ndarray, queries: np.ndarray, k: int, metric: str = "l2",
                row_group: np.ndarray | None = None, exclude_group: np.ndarray | None = None,
                prefilter: bool = False, accumulate: str = "f64", chunk: int = 64):
    """Flat kNN for a query batch.

    row_group [N] int / exclude_group [Q] int (-1 = no filter for that query) express the
    only predicate the reference ever issues, `video != "<own video>"` (datamodule.py:235).
    Returns (distance f32 [Q,k], index i64 [Q,k]); unused slots are (+inf, -1).
    """
    db = np.asarray(db, dtype=np.float32)
    queries = np.atleast_2d(np.asarray(queries, dtype=np.float32))
    nq = queries.shape[0]
    out_d = np.full((nq, k), np.inf, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for s in range(0, nq, chunk):
        dm = distances(db, queries[s:s + chunk], metric, accumulate)
        for j in range(dm.shape[0]):
            qi = s + j
            row = dm[j]
            ex = -1 if exclude_group is None else int(exclude_group[qi])
            if ex >= 0 and row_group is not None and prefilter:
                keep = np.nonzero(row_group != ex)[0]
                idx = keep[topk_rows(row[keep], k)]
            else:
                idx = topk_rows(row, k)
                if ex >= 0 and row_group is not None:
                    idx = idx[row_group[idx] != ex]     # post-filter: may leave < k rows
            out_i[qi, :idx.size] = idx
            out_d[qi, :idx.size] = row[idx]
    return out_d, out_i


# --- the reference's record interface -------------------------------------------------------
_WHERE = re.compile(r'^\s*(\w+)\s*!=\s*(["\'])((?:(?!\2).)*)\2\s*$')


def parse_where(where: str | None):
    """The one predicate shape the reference produces: `<column> != "<value>"`."""
    if where is None:
        return None
    m = _WHERE.match(where)
    if not m:
        raise ValueError(f"unsupported where clause: {where!r}")
    return m.group(1), m.group(3)


def where_mask(columns: dict, where: str) -> np.ndarray:
    """bool [N]: rows passing an SQL predicate (src/data/rag.py:56-57 hands the string to LanceDB's
    `.where`). Evaluated by SQLite over the scalar columns — an SQL engine independent of the product's
    own parser (motionrag_b200/where.py). A double-quoted token that names no column is a string
    literal there too, which is how the reference's `video != "<name>"` reads."""
    import sqlite3
    names = [c for c, v in columns.items() if np.asarray(v).ndim == 1]
    n = len(np.asarray(columns[names[0]]))
    con = sqlite3.connect(":memory:")
    con.execute("create table t(%s)" % ", ".join(f'"{c}"' for c in names))
    cells = [np.asarray(columns[c]).tolist() for c in names]
    con.executemany("insert into t values(%s)" % ",".join("?" * len(names)), zip(*cells))
    try:
        hits = [r[0] - 1 for r in con.execute(f"select rowid from t where {where}")]
    except sqlite3.Error as e:
        raise ValueError(f"unsupported where clause: {where!r} ({e})") from e
    mask = np.zeros(n, dtype=bool)
    mask[hits] = True
    return mask


class OracleRAGDatabase:
    """Restatement of `RAGDatabase` (src/data/rag.py:11-80) over an in-memory table:
    `columns` is a dict of equal-length sequences that must contain the vector column."""

    def __init__(self, columns: dict, vector_column: str = "text_embedding",
                 metric: str = "l2", prefilter: bool = False):
        self.columns = columns
        self.vector_column = vector_column
        self.metric = metric
        self.prefilter = prefilter
        self.db = np.asarray(columns[vector_column], dtype=np.float32)

    @staticmethod
    def format_result(records: list[dict], format: str = "dict"):
        if format in ("dict", "list"):
            return records
        if format == "pandas":
            import pandas as pd
            return pd.DataFrame.from_records(records)
        if format == "pyarrow":
            import pyarrow as pa
            return pa.Table.from_pylist(records)
        raise ValueError(f"Invalid format: {format}")

    def vector_search(self, vector, vector_column_name=None, top_k=10, table=None, where=None,
                      select=None, nprobes=50, refine_factor=30, output_format="dict"):
        row_group = exclude = None
        if where is not None:
            row_group = (~where_mask(self.columns, where)).astype(np.int64)  # group 1 = rows failing the predicate
            exclude = np.array([1])
        column = vector_column_name or self.vector_column
        mat = self.db if column == self.vector_column else np.asarray(self.columns[column], dtype=np.float32)
        dist, idx = flat_search(mat, np.asarray(vector, dtype=np.float32)[None], top_k,
                                self.metric, row_group, exclude, self.prefilter)
        cols = select if select is not None else [c for c in self.columns]
        recs = []
        for d, i in zip(dist[0], idx[0]):
            if i < 0:
                continue
            r = {c: _py(self.columns[c][i]) for c in cols}
            r["_distance"] = float(d)
            recs.append(r)
        return self.format_result(recs, output_format)

    def text_search(self, text, top_k=10, table=None, where=None, select=None, nprobes=50,
                    refine_factor=30, output_format="dict"):
        return self.vector_search(text, "text_embedding", top_k, table, where, select, nprobes,
                                  refine_factor, output_format)


    def image_search(self, image_embedding, top_k=10, table=None, where=None, select=None, nprobes=50,
                     refine_factor=30, output_format="dict"):
        return self.vector_search(image_embedding, "image_embedding", top_k, table, where, select, nprobes,
                                  refine_factor, output_format)

    def text_image_search(self, text, image_embedding, top_k=(20, 10), table=None, where=None, select=None,
                          nprobes=50, refine_factor=30, output_format="dict"):
        """src/data/rag.py:101-130: text top-k0 (with the where clause), the hits become a temporary
        table, image top-k1 inside it (no where clause on the second stage)."""
        db = self if table is None else table
        first = db.vector_search(text, "text_embedding", top_k[0], None, where, None, nprobes, refine_factor)
        if not first:
            return self.format_result([], output_format)
        tmp = OracleRAGDatabase({c: np.asarray([r[c] for r in first]) for c in first[0]}, metric=self.metric)
        return tmp.image_search(image_embedding, top_k[1], None, None, select, nprobes, refine_factor, output_format)


def _py(v):
    return v.item() if isinstance(v, np.generic) else v
