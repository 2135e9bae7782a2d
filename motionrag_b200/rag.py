"""`RAGDatabase` — drop-in for the reference's src/data/rag.py:11-130, served from HBM.

Same constructor, same method names, same defaults and the same return schema
(`list[dict]`, ascending `_distance`, at most `top_k` rows, keys = `select` + `_distance`), so
`VideoDataModule.prepare_data / prepare_annotations` (src/data/datamodule.py:102, 231-265)
runs unchanged. What differs is what executes: the LanceDB flat scan becomes the libmrag CUDA
kernels (K1 streaming / K2 tcgen05 scan, K3 merge + fp32 re-score + filter). There is no CPU
path — constructing the object without a B200-class GPU raises.

On-disk forms `RAGDatabase(db_path, table_name)` opens (motionrag_b200/tables.py):
    <db_path>/<table_name>/text_embedding.npy     float32 [N, dim]   (np.load mmap-able; `save_table`)
    <db_path>/<table_name>/image_embedding.npy    optional
    <db_path>/<table_name>/columns.parquet        every scalar column of the reference schema
                                                  (tools/build_rag_database.py:35-45)
    <db_path>/<table_name>.parquet | <table_name>/*.parquet | <table_name>.arrow/.feather/.ipc
                                                  an Arrow dump of the reference's table itself
                                                  (`table.to_arrow()`; FixedSizeList<f32>[768] + scalars)
"""
from __future__ import annotations

from pathlib import Path
from typing import Callable, Literal, Sequence

import numpy as np
import torch

from .store import EmbeddingStore
from .where import parse as parse_where

VECTOR_COLUMNS = ("text_embedding", "image_embedding")


MAX_K = 32       # results per scan the kernels produce (mrag.h); larger top_k runs in passes, RAGDatabase._search_deep

def _as_column(v):
    """Scalar columns are indexed with integer arrays: lists / tuples become numpy arrays."""
    return v if isinstance(v, np.ndarray) else np.asarray(v)


def _as_2d(v):
    """One vector (ndarray / Tensor / list) as a [1, dim] float32 array."""
    a = np.asarray(v.detach().float().cpu().numpy() if isinstance(v, torch.Tensor) else v, dtype=np.float32)
    return a[None] if a.ndim == 1 else a



def save_table(db_path: str | Path, table_name: str, columns: dict) -> Path:
    """Write a table in the layout above. `columns` maps names to equal-length sequences;
    vector columns are float32 [N, dim] arrays."""
    import pandas as pd
    root = Path(db_path) / table_name
    root.mkdir(parents=True, exist_ok=True)
    scalars = {}
    for name, col in columns.items():
        if name in VECTOR_COLUMNS:
            np.save(root / f"{name}.npy", np.ascontiguousarray(col, dtype=np.float32))
        else:
            scalars[name] = list(col) if not isinstance(col, np.ndarray) else col
    pd.DataFrame(scalars).to_parquet(root / "columns.parquet")
    return root


class _DeviceColumn:
    """Stand-in for a vector column that only exists in HBM (RAGDatabase.from_store)."""

    def __init__(self, store: EmbeddingStore):
        self._store = store

    @property
    def shape(self):
        return (len(self._store), self._store.dim)

    def __getitem__(self, i):
        return self._store.rows_f32()[i].cpu().numpy()


class RAGDatabase:
    def __init__(self, db_path: str | None, table_name: str | None,
                 device: Literal['cpu', 'cuda'] | str | torch.device = 'cpu', *,
                 columns: dict | None = None, embed_fn: Callable | None = None,
                 metric: str = "l2", prefilter: bool = False, normalise: bool = False,
                 path: str = "auto", recheck: str | None = "auto"):
        """db_path / table_name / device as in the reference (src/data/rag.py:12-15). In the
        reference `device` only places the query-text embedding model; the store itself always
        lives on the current CUDA device (or on `device` when it names a cuda:N).

        Keyword-only extensions: `columns` (in-memory table instead of a path), `embed_fn`
        (text -> vector, for str queries; the reference delegates that to LanceDB's registered
        gte-base-en-v1.5 function), `metric` / `prefilter` (LanceDB 0.14 defaults: "l2", post-
        filter; metric="reference" picks "dot" for tables of more than 1 M rows, which the reference indexes
        with that metric), `normalise` (L2-normalise rows on upload; the reference's tables already are),
        `path` (force a scan kernel: auto | stream_f32 | stream_bf16 | tensor_bf16), `recheck`
        (what to do with the exactness margin every bf16 scan reports per query, see mrag.h: "auto"
        re-runs a query on the fp32 master rows when its margin is below 6 sigma of the bf16
        rounding noise; "strict" whenever the margin does not rigorously prove exactness; None never
        asks for the margin. Applies to every batch size and to row-sharded tables;
        `fp32_rechecks` counts the re-runs).
        """
        self.db_path, self.table_name = db_path, table_name
        self._ctor = dict(device=str(device), metric=metric, prefilter=prefilter, normalise=normalise,
                          path=path, recheck=recheck)
        self.recheck = recheck
        self.fp32_rechecks = 0      # queries answered from the fp32 master rows after their margin failed
        self.deep_rechecks = 0      # queries re-issued with 32-entry candidate lists first
        self.embed_fn = embed_fn
        self.metric, self.prefilter, self.path = metric, prefilter, path
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        if dev.type == "cuda" and dev.index is not None:
            self.device = dev
        else:
            if not torch.cuda.is_available():
                from ._cabi import MragError
                raise MragError(-3, "RAGDatabase needs a CUDA device: the B200 path has no CPU fallback")
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._from_memory = columns is not None
        if columns is None:
            columns = self._load(db_path, table_name)
        self._columns = {k: _as_column(v) for k, v in columns.items() if k not in VECTOR_COLUMNS}
        self._vectors = {k: columns[k] for k in VECTOR_COLUMNS if k in columns}
        if not self._vectors:
            raise ValueError("table has no vector column (text_embedding / image_embedding)")
        self._stores: dict[str, EmbeddingStore] = {}
        self._normalise = normalise
        self._retriever = None
        self._init_caches()
        self.table = self  # reference attribute name (rag.py:14); `table=` arguments accept it
        self._resolve_metric()

    INDEXED_ABOVE = 1_000_000   # tools/build_rag_database.py:51-52: create_index(metric='dot') iff len(table) > 1 M

    def _resolve_metric(self) -> None:
        """metric="reference": the metric the reference's own table would answer with — squared L2 (LanceDB's
        default, src/data/rag.py:54 sets none) up to 1 M rows, `dot` above, where build_rag_database.py adds an
        IVF-PQ `dot` index. The reference is approximate in that regime (nprobes=50, refine_factor=30); this
        search stays exact, see tests/test_ivf_pq.py for the comparison that is defined there."""
        if self.metric == "reference":
            self.metric = "dot" if len(self) > self.INDEXED_ABOVE else "l2"

    def _init_caches(self) -> None:
        self._group_col: str | None = None
        self._group_ids: dict[str, dict] = {}
        self._thr_cache: dict[tuple, tuple] = {}
        self._where_cache: dict[str, tuple[str, int]] = {}   # SQL string -> (column, group id)
        self._excl_one: dict[str, tuple[str, np.ndarray]] = {}   # SQL string -> (column, int32[1] group id)
        self._rec_plans: dict[tuple, list] = {}              # select tuple -> validated column names

    @classmethod
    def from_store(cls, store: EmbeddingStore, columns: dict, vector_column: str = "text_embedding",
                   retriever=None, **kw) -> "RAGDatabase":
        """Wrap an already HBM-resident store (e.g. generated on device) with its scalar
        columns; the vector column is then not available to `select` on the host side.

        retriever: a `parallel.ShardedRetriever` over a row-sharded table (one process per GPU, `store`
        = this rank's shard). `columns` then cover ALL rows of the table (scalar metadata lives on the
        host of every rank), every rank calls the same method with the same arguments, and every rank
        gets the same records; searches go through the retriever's host-buffer entry."""
        self = object.__new__(cls)
        self.db_path = self.table_name = None
        self._ctor, self.embed_fn = {}, kw.get("embed_fn")
        self.metric, self.prefilter = kw.get("metric", "l2"), kw.get("prefilter", False)
        self.path = kw.get("path", "auto")
        self.recheck = kw.get("recheck", "auto")
        self.fp32_rechecks = self.deep_rechecks = 0
        self.device = store.device
        self._from_memory = True
        self._columns = {k: _as_column(v) for k, v in columns.items() if k not in VECTOR_COLUMNS}
        self._vectors = {vector_column: _DeviceColumn(store)}
        self._stores = {vector_column: store}
        self._normalise = False
        self._retriever = retriever
        self._init_caches()
        self.table = self
        self._resolve_metric()
        return self

    # -- loading ------------------------------------------------------------------------------
    @staticmethod
    def _load(db_path, table_name) -> dict:
        """Own layout, a Parquet file / fragment directory or an Arrow IPC file holding the reference's
        schema (tools/build_rag_database.py:35-45) — see motionrag_b200/tables.py."""
        from .tables import read_table
        return read_table(db_path, table_name)

    def __len__(self) -> int:
        if self._retriever is not None and self._columns:
            return int(len(next(iter(self._columns.values()))))
        return int(next(iter(self._vectors.values())).shape[0])

    def _store(self, column: str) -> EmbeddingStore:
        if column not in self._vectors:
            raise ValueError(f"table has no vector column {column!r}")
        st = self._stores.get(column)
        if st is None:
            vec = self._vectors[column]
            n, dim = vec.shape
            st = EmbeddingStore(dim, n, self.device)
            step = max(1, (256 << 20) // (dim * 4))          # 256 MB host chunks
            for s in range(0, n, step):
                st.append(np.ascontiguousarray(vec[s:s + step], dtype=np.float32), normalise=self._normalise)
            self._stores[column] = st
            self._group_col = None
        return st

    def _groups_for(self, column: str):
        """Dense int ids for a scalar column (the `video != x` predicate); cached per column."""
        g = self._group_ids.get(column)
        if g is None:
            values = np.asarray(self._columns[column])
            uniq, inv = np.unique(values, return_inverse=True)
            g = {"ids": inv.astype(np.int32), "lookup": {v: i for i, v in enumerate(uniq.tolist())}}
            self._group_ids[column] = g
        return g

    def _bind_groups(self, column: str) -> dict:
        g = self._groups_for(column)
        if self._group_col != column:
            for st in self._stores.values():
                ids = g["ids"]
                if self._retriever is not None and self._retriever.world > 1:   # this rank's rows only
                    lo = self._retriever.rank * self._retriever.rows_per_shard
                    ids = ids[lo:lo + len(st)]
                st.set_groups(ids)
            self._group_col = column
        return g

    # -- pickling: the reference ships the bound method into spawn workers (datamodule.py:257) --
    def __getstate__(self):
        if self._from_memory:
            raise TypeError("an in-memory RAGDatabase cannot be pickled; use a db_path, or call "
                            "search_batch() instead of a process pool")
        return {"db_path": self.db_path, "table_name": self.table_name, "ctor": self._ctor,
                "embed_fn": self.embed_fn}

    def __setstate__(self, st):
        c = st["ctor"]
        self.__init__(st["db_path"], st["table_name"], c["device"], embed_fn=st["embed_fn"],
                      metric=c["metric"], prefilter=c["prefilter"], normalise=c["normalise"], path=c["path"],
                      recheck=c.get("recheck", "auto"))

    # -- reference API ------------------------------------------------------------------------
    @staticmethod
    def format_result(result, format: Literal["pandas", "pyarrow", "dict", "list"] = 'dict'):
        """src/data/rag.py:17-34. `result` is the list of record dicts of one query."""
        if format == 'pandas':
            import pandas as pd
            return pd.DataFrame.from_records(result)
        elif format == 'pyarrow':
            import pyarrow as pa
            return pa.Table.from_pylist(result)
        elif format == 'dict':
            return result
        elif format == 'list':
            return result
        else:
            raise ValueError(f'Invalid format: {format}')

    def _as_host_queries(self, vector) -> tuple[np.ndarray, bool]:
        """Any accepted query form -> (contiguous float32 [nq, dim] host array, was it a single vector)."""
        if isinstance(vector, np.ndarray):
            q = vector if vector.dtype == np.float32 else vector.astype(np.float32)
        elif isinstance(vector, torch.Tensor):
            q = vector.detach().to(torch.float32).cpu().numpy()
        else:
            if isinstance(vector, str) or (isinstance(vector, (list, tuple)) and vector and isinstance(vector[0], str)):
                if self.embed_fn is None:
                    raise NotImplementedError(
                        "text queries need an embed_fn (the reference lets LanceDB run gte-base-en-v1.5; "
                        "src/data/datamodule.py:296-304 always passes precomputed vectors)")
                was_str = isinstance(vector, str)
                vector = self.embed_fn(vector)
                if isinstance(vector, torch.Tensor):
                    vector = vector.detach().to(torch.float32).cpu().numpy()
                q = np.asarray(vector, dtype=np.float32)
                if was_str and q.ndim == 2 and q.shape[0] == 1:
                    q = q[0]
            else:
                q = np.asarray(vector, dtype=np.float32)
        single = q.ndim == 1
        if single:
            q = q[None]
        if q.ndim != 2:
            raise ValueError(f"query must be [dim] or [nq, dim], got {tuple(q.shape)}")
        if not q.flags.c_contiguous:
            q = np.ascontiguousarray(q)
        return q, single

    def _as_queries(self, vector) -> tuple[torch.Tensor, bool]:
        if isinstance(vector, str) or (isinstance(vector, (list, tuple)) and vector and isinstance(vector[0], str)):
            if self.embed_fn is None:
                raise NotImplementedError(
                    "text queries need an embed_fn (the reference lets LanceDB run gte-base-en-v1.5; "
                    "src/data/datamodule.py:296-304 always passes precomputed vectors)")
            vector = self.embed_fn(vector)
        if isinstance(vector, torch.Tensor):
            q = vector.detach()
        else:
            q = torch.from_numpy(np.ascontiguousarray(np.asarray(vector), dtype=np.float32))
        single = q.ndim == 1
        if single:
            q = q[None]
        if q.ndim != 2:
            raise ValueError(f"query must be [dim] or [nq, dim], got {tuple(q.shape)}")
        q = q.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        return q, single

    def _where_to_group(self, w: str):
        """One where clause -> (column, dense group id or -1) when it is the reference's own shape
        `<column> != "<value>"` (src/data/datamodule.py:235; runs on the device as an excluded group id), else
        the parsed `where.Predicate` (evaluated on the host, see motionrag_b200/where.py). Parsed once per string."""
        hit = self._where_cache.get(w)
        if hit is None:
            pred = parse_where(w)
            unknown = pred.columns() - set(self._columns)
            if unknown:
                raise ValueError(f"where clause names unknown column {sorted(unknown)[0]!r}")
            if len(self._where_cache) > (1 << 20):
                self._where_cache.clear()
            simple = pred.simple_exclusion()
            hit = self._where_cache[w] = pred if simple is None else (simple[0], self._group_ids_lookup(simple[0], simple[1]), pred)
        return hit

    def _bind_mask(self, pred) -> None:
        """Pre-filter with a general predicate: rows that fail it form group 1 of a two-group labelling the scan
        kernels skip (evaluated once over the whole table, cached while the same clause is in use)."""
        key = ("where", pred.text)
        if self._group_col != key:
            ids = (~pred.evaluate(self._columns)).astype(np.int32)
            for st in self._stores.values():
                part = ids
                if self._retriever is not None and self._retriever.world > 1:
                    lo = self._retriever.rank * self._retriever.rows_per_shard
                    part = ids[lo:lo + len(st)]
                st.set_groups(part)
            self._group_col = key

    def _exclusion_ids(self, where, nq: int):
        """`where` is one SQL string (reference form) or one per query (batched form);
        -> (int32 [nq] group ids to exclude (-1 = none) or None, per-query host predicates or None)."""
        if where is None:
            return None, None
        if nq == 1 and type(where) is str:
            # the reference's call pattern (one query, one `video != "<own>"` clause): one dict lookup
            one = self._excl_one.get(where)
            if one is None:
                hit = self._where_to_group(where)
                if isinstance(hit, tuple):
                    if len(self._excl_one) > (1 << 20):
                        self._excl_one.clear()
                    one = self._excl_one[where] = (hit[0], np.array([hit[1]], dtype=np.int32))
            if one is not None:
                if self._group_col != one[0]:
                    self._bind_groups(one[0])
                return one[1], None
        wheres = [where] * nq if isinstance(where, str) else list(where)
        if len(wheres) != nq:
            raise ValueError("need one where clause per query")
        col, ids, preds = None, np.full(nq, -1, dtype=np.int32), None
        for i, w in enumerate(wheres):
            if w is None:
                continue
            hit = self._where_to_group(w) if i == 0 or w is not wheres[i - 1] else hit
            if isinstance(hit, tuple) and (col is None or col == hit[0]):
                col, ids[i] = hit[0], hit[1]
            else:
                # general predicates — and `!=` clauses on a second column, since the device holds one group
                # labelling at a time — are evaluated on the host
                if isinstance(hit, tuple) and self.prefilter:
                    raise ValueError("all `!=` where clauses of a pre-filtered batch must name the same column")
                if preds is None:
                    preds = [None] * nq
                preds[i] = hit[2] if isinstance(hit, tuple) else hit
        if preds is not None and self.prefilter:
            first = next(p for p in preds if p is not None)
            if col is not None or any(p is not first for p in preds):
                raise ValueError("pre-filter mode takes ONE general where clause per batch (or `!=` clauses only)")
            self._bind_mask(first)
            return np.ones(nq, dtype=np.int32), None
        if col is None:
            return None, preds
        if self._group_col != col:
            self._bind_groups(col)
        return ids, preds

    def _apply_predicates(self, preds, dist: np.ndarray, idx: np.ndarray) -> None:
        """Post-filter of general where clauses (LanceDB's `.where()` without prefilter: the k nearest rows are
        found first, rows failing the predicate are dropped): compacts each query's result in place."""
        for i, p in enumerate(preds):
            if p is None:
                continue
            valid = idx[i] >= 0
            rows = idx[i][valid]
            keep = p.evaluate(self._columns, rows)
            n = int(keep.sum())
            d = dist[i][valid][keep]
            idx[i, :n], idx[i, n:] = rows[keep], -1
            dist[i, :n], dist[i, n:] = d, np.inf

    def _group_ids_lookup(self, col: str, value) -> int:
        g = self._groups_for(col)
        v = g["lookup"].get(value)
        if v is None:  # numeric columns arrive as strings inside the SQL literal
            for cast in (int, float):
                try:
                    v = g["lookup"].get(cast(value))
                except ValueError:
                    v = None
                if v is not None:
                    break
        return -1 if v is None else int(v)

    def _records(self, dist: np.ndarray, idx: np.ndarray, select: Sequence[str] | None) -> list[list[dict]]:
        nq, k = idx.shape
        names = self._rec_plans.get(tuple(select)) if select is not None else None
        if names is None:
            names = list(select) if select is not None else list(self._columns) + list(self._vectors)
            for c in names:
                if c not in self._columns and c not in self._vectors:
                    raise ValueError(f"unknown column {c!r} in select")
            if select is not None:
                self._rec_plans[tuple(select)] = names
        if nq <= 4 and not any(c in self._vectors for c in names):
            # the reference's call pattern (one query per call): one 12-row fancy index + tolist() per column.
            # Columns stay numpy arrays — Python lists of a million cells would be re-traversed by every full
            # garbage collection (tens of ms per stall, measured) and cost the same per lookup at that size.
            out = []
            for qi in range(nq):
                rows = idx[qi]
                if rows[k - 1] < 0:                      # valid entries come first (the kernels compact them)
                    rows = rows[:int((rows >= 0).sum())]
                d = dist[qi, :rows.shape[0]].tolist()
                if len(names) == 3:      # select=['video', 'start_sec', 'end_sec'] (src/data/datamodule.py:236)
                    k0, k1, k2 = names
                    out.append([{k0: a, k1: b, k2: c, "_distance": e}
                                for a, b, c, e in zip(self._columns[k0][rows].tolist(), self._columns[k1][rows].tolist(),
                                                      self._columns[k2][rows].tolist(), d)])
                else:
                    keys = (*names, "_distance")
                    out.append([dict(zip(keys, vals)) for vals in zip(*(self._columns[c][rows].tolist() for c in names), d)])
            return out
        # one fancy-index + tolist() per column for the WHOLE batch, then C-speed dict(zip(...));
        # per-cell numpy scalar handling would dominate a 4096-query batch
        valid = idx >= 0
        flat = idx[valid]
        counts = valid.sum(-1).tolist()
        keys = names + ["_distance"]
        per_col = []
        for c in names:
            if c in self._columns:
                per_col.append(self._columns[c][flat].tolist())
            else:
                col = self._vectors[c]
                per_col.append([np.array(col[int(i)], dtype=np.float32) for i in flat])
        per_col.append(dist[valid].tolist())
        recs = [dict(zip(keys, vals)) for vals in zip(*per_col)]
        out, pos = [], 0
        for n in counts:
            out.append(recs[pos:pos + n])
            pos += n
        return out

    def _search(self, vector, vector_column_name, top_k, where, refine_factor, exclude_group=None, reuse=False):
        """-> (distance f32 [nq,k], index i64 [nq,k]) numpy arrays, single?

        Every bf16 scan is certified: the kernels report the exactness margin of each query (mrag.h)
        and queries that fail the test of `recheck` are re-run on the fp32 master rows — for any
        batch size, for host and device inputs, for single-GPU and row-sharded tables."""
        column = vector_column_name or "text_embedding"
        store = self._store(column)
        q, single = self._as_host_queries(vector)
        if int(top_k) > MAX_K:
            dist, idx = self._search_deep(store, q, int(top_k), where, refine_factor, exclude_group)
            return dist, idx, single
        mode = "pre" if self.prefilter else "post"
        preds = None
        if exclude_group is None:
            excl, preds = self._exclusion_ids(where, q.shape[0])
        else:
            excl = np.ascontiguousarray(exclude_group, dtype=np.int32)
        dist, idx = self._scan(store, q, int(top_k), excl, mode, refine_factor, reuse)
        if preds is not None:
            self._apply_predicates(preds, dist, idx)
        return dist, idx, single

    def _scan(self, store, q, top_k: int, excl, mode: str, refine_factor, reuse: bool = False):
        """One certified scan of at most MAX_K results per query -> (distance, index) numpy arrays."""
        refine = int(min(64, max(top_k, top_k * max(1, int(refine_factor)))))
        certify = self.recheck is not None and self.path != "stream_f32"
        searcher = self._retriever if self._retriever is not None else store
        # reuse: result arrays are recycled between small calls (the caller builds its records right away)
        res = searcher.search_host(q, top_k, metric=self.metric, path=self.path, refine=refine,
                                   exclude_group=excl, filter_mode=mode, certify=certify,
                                   reuse=reuse and q.shape[0] <= 4)
        dist, idx = res[0], res[1]
        if certify:
            self._recheck(searcher, store, q, excl, dist, idx, res[3], top_k, mode)
        return dist, idx

    def _search_deep(self, store, q, top_k: int, where, refine_factor, exclude_group):
        """top_k above the kernels' MAX_K = 32 results per scan (LanceDB takes any `limit`; the reference's own
        callers stay below it, src/data/datamodule.py:234,241): the k nearest rows are collected in passes of 32.
        Every pass scans with the rows already returned labelled as an excluded group (pre-filter in the scan, so
        a pass returns exactly the next-nearest rows; ties keep the (distance, row) order across passes), one
        query at a time because the label array is per query. `where` keeps its meaning: a post-filter over the
        k nearest rows, or — `prefilter=True` — rows failing it are labelled taken before the first pass.
        Costs one label upload (4 B per row) and one scan per 32 results; it exists for API completeness."""
        if exclude_group is not None:
            raise ValueError(f"exclude_group= is limited to top_k <= {MAX_K}; pass the clause as where=")
        nq, n_rows = q.shape[0], len(self)
        wheres = [where] * nq if (where is None or isinstance(where, str)) else list(where)
        if len(wheres) != nq:
            raise ValueError("need one where clause per query")
        preds = []
        for w in wheres:
            hit = None if w is None else self._where_to_group(w)
            preds.append(hit[2] if isinstance(hit, tuple) else hit)
        dist = np.full((nq, top_k), np.inf, dtype=np.float32)
        idx = np.full((nq, top_k), -1, dtype=np.int64)
        one = np.ones(1, dtype=np.int32)
        sharded = self._retriever is not None and self._retriever.world > 1
        try:
            for qi in range(nq):
                taken = np.zeros(n_rows, dtype=np.int32)
                if self.prefilter and preds[qi] is not None:
                    taken[~preds[qi].evaluate(self._columns)] = 1
                got = 0
                while got < top_k:
                    k = min(MAX_K, top_k - got)
                    self._group_col = ("deep", qi, got)
                    for st in self._stores.values():
                        part = taken
                        if sharded:
                            lo = self._retriever.rank * self._retriever.rows_per_shard
                            part = taken[lo:lo + len(st)]
                        st.set_groups(part)
                    d, i = self._scan(store, q[qi:qi + 1], k, one, "pre", refine_factor)
                    n = int((i[0] >= 0).sum())
                    dist[qi, got:got + n], idx[qi, got:got + n] = d[0, :n], i[0, :n]
                    taken[i[0, :n]] = 1
                    got += n
                    if n < k:           # fewer eligible rows than asked for
                        break
        finally:
            self._group_col = None      # the next ordinary call binds its own labelling again
        if not self.prefilter and any(p is not None for p in preds):
            self._apply_predicates(preds, dist, idx)
        return dist, idx

    def _scan_profile(self, store, nq: int, top_k: int) -> tuple[str, int, float]:
        """(scan path AUTO resolves to, length of its candidate lists, margin threshold) for this call shape."""
        from .store import PATH_NAME, margin_threshold
        key = (id(store), nq == 1, int(top_k))
        prof = self._thr_cache.get(key)
        if prof is None:
            if len(store) == 0:      # an empty shard of a sharded table: same resolution rules, nothing to plan
                used = "stream_bf16" if (nq == 1 and self.path in ("auto", "stream_bf16")) else "tensor_bf16"
                rerank = 32 if used == "stream_bf16" or top_k > 12 else 16
            else:
                plan = store.plan(nq, k=int(top_k), path=self.path)
                used, rerank = PATH_NAME[plan.path], int(plan.rerank)
            thr = margin_threshold(used, store.dim, self.recheck == "strict", store.info().max_norm_deviation)
            prof = self._thr_cache[key] = (used, rerank, thr)
        return prof

    def _recheck(self, searcher, store, q, excl, dist, idx, margin, top_k, mode) -> None:
        """Queries whose bf16-scan result is not certified exact (margin <= threshold, see mrag.h) are re-run
        and patched in place, in two stages: (1) tensor-path queries that kept 16 candidates are re-issued
        together with 32-entry lists — the k-th result is then compared with the 32nd instead of the 16th best
        scan score, which certifies almost all of them at a fraction of a full-batch scan; (2) what is still
        in doubt is answered from the fp32 master rows (4 queries per pass over the table)."""
        used, rerank, thr = self._scan_profile(store, q.shape[0], top_k)
        if q.shape[0] == 1:
            if margin[0] > thr:
                return
            doubt = np.zeros(1, dtype=np.int64)
        else:
            doubt = np.nonzero(~(margin > thr))[0]          # NaN counts as doubt
            if doubt.size == 0:
                return
        ex = (lambda rows: None if excl is None else excl[rows])
        if used == "tensor_bf16" and rerank < 32 and top_k < 32:
            self.deep_rechecks += int(doubt.size)
            r2 = searcher.search_host(q[doubt], int(top_k), metric=self.metric, path="tensor_bf16", list_len=32,
                                      exclude_group=ex(doubt), filter_mode=mode, certify=True)
            dist[doubt], idx[doubt] = r2[0], r2[1]
            from .store import margin_threshold
            thr2 = margin_threshold("tensor_bf16", store.dim, self.recheck == "strict", store.info().max_norm_deviation)
            doubt = doubt[~(r2[3] > thr2)]
        self.fp32_rechecks += int(doubt.size)
        for s in range(0, doubt.size, 4):
            rows = doubt[s:s + 4]
            r2 = searcher.search_host(q[rows], int(top_k), metric=self.metric, path="stream_f32",
                                      exclude_group=ex(rows), filter_mode=mode)
            dist[rows], idx[rows] = r2[0], r2[1]

    def search_arrays(self, vectors, top_k: int = 10, where: Sequence[str | None] | str | None = None,
                      refine_factor: int = 30, vector_column_name: str = "text_embedding", batch: int = 4096,
                      exclude_group=None):
        """The certified search without record building: -> (distance f32 [nq, k], index i64 [nq, k]) numpy
        arrays (unused slots +inf / -1). Same scans, filters and re-checks as search_batch. `exclude_group`
        (int32 [nq], -1 = none) may replace `where` when the caller already holds the group ids the store
        was given with set_groups."""
        q_all, _ = self._as_host_queries(vectors)
        wheres = None if where is None else ([where] * q_all.shape[0] if isinstance(where, str) else list(where))
        ds, is_ = [], []
        for s in range(0, q_all.shape[0], batch):
            w = None if wheres is None else wheres[s:s + batch]
            d, i, _ = self._search(q_all[s:s + batch], vector_column_name, top_k, w, refine_factor,
                                   None if exclude_group is None else exclude_group[s:s + batch])
            ds.append(d)
            is_.append(i)
        return np.concatenate(ds), np.concatenate(is_)

    def vector_search(self, vector, vector_column_name: str = None, top_k: int = 10, table=None,
                      where: str = None, select: list[str] = None, nprobes: int = 50, refine_factor: int = 30,
                      output_format: Literal["pandas", "pyarrow", "dict"] = 'dict'):
        """src/data/rag.py:36-61. `nprobes` is accepted and ignored (the scan is exact);
        `refine_factor` sizes the fp32 re-rank of the bf16 scan paths. A [nq, dim] batch returns
        one result per query."""
        if output_format not in ("pandas", "pyarrow", "dict", "list"):
            raise ValueError(f'Invalid format: {output_format}')
        db = self if table is None else table
        dist, idx, single = db._search(vector, vector_column_name, top_k, where, refine_factor, reuse=True)
        recs = db._records(dist, idx, select)
        if single:
            return self.format_result(recs[0], output_format)
        return [self.format_result(r, output_format) for r in recs]

    def text_search(self, text, top_k: int = 10, table=None, where: str = None, select: list[str] = None,
                    nprobes: int = 50, refine_factor: int = 30,
                    output_format: Literal["pandas", "pyarrow", "dict"] = 'dict'):
        """src/data/rag.py:63-80."""
        return self.vector_search(text, vector_column_name="text_embedding", top_k=top_k, table=table,
                                  where=where, select=select, nprobes=nprobes, refine_factor=refine_factor,
                                  output_format=output_format)

    def image_search(self, image_embedding, top_k: int = 10, table=None, where: str = None,
                     select: list[str] = None, nprobes: int = 50, refine_factor: int = 30,
                     output_format: Literal["pandas", "pyarrow", "dict"] = 'dict'):
        """src/data/rag.py:82-99."""
        return self.vector_search(image_embedding, vector_column_name="image_embedding", top_k=top_k,
                                  table=table, where=where, select=select, nprobes=nprobes,
                                  refine_factor=refine_factor, output_format=output_format)

    def text_image_search(self, text, image_embedding, top_k: tuple[int, int] = (20, 10), table=None,
                          where: str = None, select: list[str] = None, nprobes: int = 50,
                          refine_factor: int = 30,
                          output_format: Literal["pandas", "pyarrow", "dict"] = 'dict'):
        """src/data/rag.py:101-130: text top-k0, then image-embedding top-k1 inside that
        candidate set (the reference materialises a temporary table; here the k0 candidate rows are
        scored against the image column by one kernel, see text_image_search_batch)."""
        db = self if table is None else table
        texts = [text] if isinstance(text, str) else _as_2d(text)
        if len(texts) != 1:
            raise ValueError("text_image_search takes one query at a time, like the reference")
        res = db.text_image_search_batch(texts, _as_2d(image_embedding), top_k, where, select, refine_factor)
        return self.format_result(res[0], output_format)

    def text_image_search_batch(self, texts, image_embeddings, top_k: tuple[int, int] = (20, 10),
                                where: Sequence[str | None] | str | None = None, select: list[str] | None = None,
                                refine_factor: int = 30, batch: int = 4096) -> list[list[dict]]:
        """The two-stage search for many queries at once: one text scan per `batch` queries (k0 hits
        each, with the where clause), then ONE kernel that scores every query's k0 candidate rows
        against the image-embedding column in fp32 and keeps the best k1 (mrag_rescore_rows) — the
        candidate rows play the part of the reference's temporary table (src/data/rag.py:118-128)."""
        k0, k1 = int(top_k[0]), int(top_k[1])
        if self._retriever is not None and self._retriever.world > 1:
            raise NotImplementedError("the two-stage text -> image search runs on a single-GPU table")
        q_t, _ = self._as_host_queries(texts)
        q_i, _ = self._as_queries(image_embeddings)
        if q_t.shape[0] != q_i.shape[0]:
            raise ValueError("need one image embedding per text query")
        img = self._store("image_embedding")
        wheres = None if where is None else ([where] * q_t.shape[0] if isinstance(where, str) else list(where))
        out: list[list[dict]] = []
        for s in range(0, q_t.shape[0], batch):
            w = None if wheres is None else wheres[s:s + batch]
            _, idx0, _ = self._search(q_t[s:s + batch], "text_embedding", k0, w, refine_factor)
            cand = torch.from_numpy(np.ascontiguousarray(idx0)).to(self.device)
            d1, i1 = img.rescore(q_i[s:s + batch].contiguous(), cand, min(k1, 64), self.metric)
            out.extend(self._records(d1.cpu().numpy(), i1.cpu().numpy(), select))
        return out

    # -- batched fast path (replaces the spawn pool of datamodule.py:257-262) -----------------------
    def search_batch(self, vectors, top_k: int = 10, where: Sequence[str | None] | str | None = None,
                     select: list[str] | None = None, refine_factor: int = 30,
                     vector_column_name: str = "text_embedding", batch: int = 4096) -> list[list[dict]]:
        """One scan per `batch` queries instead of one LanceDB call per annotation; returns the
        same per-query record lists the reference stores in `anno['ref_videos']`
        (src/data/datamodule.py:264-265)."""
        q_all, _ = self._as_host_queries(vectors)
        wheres = None if where is None else ([where] * q_all.shape[0] if isinstance(where, str) else list(where))
        out: list[list[dict]] = []
        for s in range(0, q_all.shape[0], batch):
            w = None if wheres is None else wheres[s:s + batch]
            dist, idx, _ = self._search(q_all[s:s + batch], vector_column_name, top_k, w, refine_factor)
            out.extend(self._records(dist, idx, select))
        return out

    def retrieve_for_annotations(self, annotations: list[dict], ref_video_num: int, batch: int = 4096,
                                 ref_video_type: str = "rag_text", save_path=None) -> list[dict]:
        """The retrieval branches of VideoDataModule.prepare_annotations
        (src/data/datamodule.py:231-245, 257-268) without the spawn pool.

        `rag_text`: one batched scan per `batch` annotations — k = ref_video_num + 3,
        `where video != "<own video>"`, select video/start_sec/end_sec. `rag_text_image`: the
        two-stage search with top_k = (2*ref_video_num + 3, ref_video_num), batched the same way
        (text scan, then one kernel scoring each annotation's candidates against the image column). The records are attached as `anno['ref_videos']` and,
        like the reference (:268), the list is written with torch.save when `save_path` is given."""
        if ref_video_type == "rag_text":
            vec = np.stack([np.asarray(a['text_embedding'], dtype=np.float32) for a in annotations])
            wheres = [f'video != "{a["video"]}"' for a in annotations]
            results = self.search_batch(vec, top_k=ref_video_num + 3, where=wheres,
                                        select=['video', 'start_sec', 'end_sec'], batch=batch)
        elif ref_video_type == "rag_text_image":
            results = self.text_image_search_batch(
                np.stack([np.asarray(a['text_embedding'], dtype=np.float32) for a in annotations]),
                np.stack([np.asarray(a['image_embedding'], dtype=np.float32) for a in annotations]),
                top_k=(ref_video_num * 2 + 3, ref_video_num), where=[f'video != "{a["video"]}"' for a in annotations],
                select=['video', 'start_sec', 'end_sec'], batch=batch)
        else:
            raise ValueError("Invalid ref_video_type.")   # 'gt' / 'random' never touch the database
        for anno, r in zip(annotations, results):
            anno['ref_videos'] = r
        if save_path is not None:
            torch.save(annotations, save_path)
        return annotations
