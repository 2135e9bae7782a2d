"""Readers for the RAG table `RAGDatabase(db_path, table_name)` opens (reference: src/data/rag.py:13-14
opens a LanceDB table written by tools/build_rag_database.py:28-52).

Schema of that table (build_rag_database.py:35-45): text:str, text_embedding:FixedSizeList<f32>[768],
id:int, uid:str, dataset:str, video:str, start_sec:f64, end_sec:f64 [, image_embedding]. Any Arrow
container holding it can be opened without `lancedb`:

    <db_path>/<table_name>/text_embedding.npy + columns.parquet     this repo's own layout (rag.save_table)
    <db_path>/<table_name>.parquet | <db_path>/<table_name>/*.parquet   Parquet file / directory of fragments
    <db_path>/<table_name>.arrow | .feather | .ipc | .arrows         Arrow IPC file or stream

i.e. what `lancedb.connect(db).open_table(t).to_arrow()` gives, written with `pyarrow.parquet.write_table`
or `pyarrow.feather.write_feather` (tools/export_lancedb.py does exactly that on a machine that has the
wheel). The `.lance` directory itself is LanceDB's private file format and is not parsed here.
Everything in this module is host-side pyarrow/numpy — no GPU needed, so the CPU tests cover it.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

VECTOR_COLUMNS = ("text_embedding", "image_embedding")
IPC_SUFFIXES = (".arrow", ".feather", ".ipc", ".arrows")


def vector_column_to_numpy(col, name: str = "vector") -> np.ndarray:
    """A pyarrow (Chunked)Array of FixedSizeList<float>[dim] / List<float> -> float32 [N, dim].
    NULL rows (LanceDB's `on_bad_vectors='fill'`, build_rag_database.py:47, writes fill values; a NULL
    can still come from a hand-made file) become zero rows, which the store reports in `zero_rows`."""
    import pyarrow as pa
    import pyarrow.compute as pc
    if isinstance(col, pa.ChunkedArray):
        col = col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
    n = len(col)
    t = col.type
    if pa.types.is_fixed_size_list(t):
        dim = t.list_size
        flat = col.values.slice(col.offset * dim, n * dim)     # every slot, NULL rows included
        if flat.null_count:
            flat = pc.fill_null(flat, 0)
        arr = np.asarray(flat.to_numpy(zero_copy_only=False), dtype=np.float32)
        if arr.size != n * dim:
            raise ValueError(f"column {name!r}: {arr.size} values for {n} rows of width {dim}")
        arr = arr.reshape(n, dim)
        if col.null_count:
            arr = np.array(arr)
            arr[~np.asarray(pc.is_valid(col).to_numpy(zero_copy_only=False))] = 0
        return np.ascontiguousarray(arr)
    if pa.types.is_list(t) or pa.types.is_large_list(t):
        lens = pc.list_value_length(col)
        widths = set(pc.unique(pc.drop_null(lens)).to_pylist())
        if len(widths) != 1:
            raise ValueError(f"column {name!r}: rows have different lengths {sorted(widths)[:4]}")
        dim = widths.pop()
        if col.null_count:
            out = np.zeros((n, dim), dtype=np.float32)
            ok = np.asarray(pc.is_valid(col).to_numpy(zero_copy_only=False))
            out[ok] = np.asarray(col.flatten().to_numpy(zero_copy_only=False), dtype=np.float32).reshape(-1, dim)
            return out
        return np.ascontiguousarray(np.asarray(col.flatten().to_numpy(zero_copy_only=False),
                                               dtype=np.float32).reshape(n, dim))
    raise ValueError(f"column {name!r} has Arrow type {t}, expected FixedSizeList<float>[dim]")


def arrow_to_columns(table) -> dict:
    """pyarrow.Table with the reference schema -> {name: ndarray}; vector columns float32 [N, dim],
    strings as object arrays (None = NULL), numbers as numpy numbers (NULL -> NaN for floats)."""
    import pyarrow as pa
    cols = {}
    for name in table.column_names:
        col = table.column(name)
        t = col.type
        if pa.types.is_fixed_size_list(t) or pa.types.is_list(t) or pa.types.is_large_list(t):
            if name not in VECTOR_COLUMNS and name != "vector":
                continue                      # other nested columns are not part of the search schema
            cols["text_embedding" if name == "vector" else name] = vector_column_to_numpy(col, name)
        elif pa.types.is_dictionary(t):
            cols[name] = np.asarray(col.cast(t.value_type).to_numpy(zero_copy_only=False))
        else:
            cols[name] = np.asarray(col.to_numpy(zero_copy_only=False))
    return cols


def _read_ipc(path: Path):
    import pyarrow as pa
    import pyarrow.ipc as ipc
    with pa.memory_map(str(path), "r") as src:
        try:
            return ipc.open_file(src).read_all()
        except pa.ArrowInvalid:
            src.seek(0)
            return ipc.open_stream(src).read_all()


def resolve(db_path, table_name: str) -> tuple[str, Path]:
    """-> (kind, path) of the first container found for the table; kind in own | parquet | parquet_dir | ipc."""
    root = Path(db_path)
    d = root / table_name
    if d.is_dir() and (d / "columns.parquet").exists():
        return "own", d
    if (root / f"{table_name}.parquet").is_file():
        return "parquet", root / f"{table_name}.parquet"
    for suf in IPC_SUFFIXES:
        if (root / f"{table_name}{suf}").is_file():
            return "ipc", root / f"{table_name}{suf}"
    if d.is_dir() and any(d.glob("*.parquet")):
        return "parquet_dir", d
    if d.is_dir():
        for suf in IPC_SUFFIXES:
            hits = sorted(d.glob(f"*{suf}"))
            if hits:
                return "ipc", hits[0]
    if (root / f"{table_name}.lance").exists():
        raise FileNotFoundError(
            f"{root / (table_name + '.lance')} is a LanceDB table directory: its file format is private to the `lance` "
            f"library. Export it once where lancedb is installed (tools/export_lancedb.py {root} {table_name} <out>) "
            f"and open <out> instead.")
    raise FileNotFoundError(f"no table {table_name!r} under {root} (looked for {table_name}/, {table_name}.parquet, "
                            f"{table_name}{{{','.join(IPC_SUFFIXES)}}})")


def read_table(db_path, table_name: str) -> dict:
    """All columns of the table as numpy arrays (vector columns memory-mapped when the layout allows)."""
    kind, path = resolve(db_path, table_name)
    if kind == "own":
        import pandas as pd
        cols = {c: v.to_numpy() for c, v in pd.read_parquet(path / "columns.parquet").items()}
        for name in VECTOR_COLUMNS:
            f = path / f"{name}.npy"
            if f.exists():
                cols[name] = np.load(f, mmap_mode="r")
        return cols
    import pyarrow.parquet as pq
    if kind == "parquet":
        return arrow_to_columns(pq.read_table(path))
    if kind == "parquet_dir":
        import pyarrow as pa
        return arrow_to_columns(pa.concat_tables([pq.read_table(f) for f in sorted(path.glob("*.parquet"))]))
    return arrow_to_columns(_read_ipc(path))


def check_table(cols: dict, norm_tol: float = 1e-2) -> dict:
    """Facts a loader should look at before trusting a real table (build_rag_database.py:40):
    row count agreement, whether `id` is the row number (the feature table is row-aligned with it), the range of
    the embedding row norms (LanceDB's L2 ranking equals the cosine ranking only for unit rows; the store ranks
    non-unit rows exactly anyway through its per-row bias) and the number of all-zero rows."""
    n = {k: len(v) for k, v in cols.items()}
    if len(set(n.values())) != 1:
        raise ValueError(f"columns have different lengths: {n}")
    rows = next(iter(n.values()))
    out = {"rows": rows, "id_is_row_number": None}
    if "id" in cols:
        ids = np.asarray(cols["id"])
        out["id_is_row_number"] = bool(ids.dtype.kind in "iu" and np.array_equal(ids, np.arange(rows)))
    for name in VECTOR_COLUMNS:
        if name in cols:
            v = cols[name]
            lo, hi, zeros = np.inf, 0.0, 0
            for s in range(0, rows, 1 << 16):
                nr = np.sqrt((np.asarray(v[s:s + (1 << 16)], dtype=np.float64) ** 2).sum(-1))
                zeros += int((nr == 0).sum())
                nz = nr[nr > 0]
                if nz.size:
                    lo, hi = min(lo, float(nz.min())), max(hi, float(nz.max()))
            out[name] = {"dim": int(v.shape[1]), "norm_min": lo, "norm_max": hi, "zero_rows": zeros,
                         "unit_norm": bool(rows == 0 or (abs(lo - 1) <= norm_tol and abs(hi - 1) <= norm_tol))}
    return out


def feature_row_alignment(cols: dict, n_feature_rows: int) -> None:
    """The motion-feature table is indexed by search-result ROW numbers. The reference writes `id` =
    position in the annotation list (build_rag_database.py:40), so a feature table built in annotation order
    is aligned iff `id` equals the row number; raise otherwise instead of gathering the wrong clips."""
    info = check_table({k: v for k, v in cols.items() if k not in VECTOR_COLUMNS} or cols)
    if info["rows"] != n_feature_rows:
        raise ValueError(f"feature table has {n_feature_rows} rows, the RAG table {info['rows']}")
    if info["id_is_row_number"] is False:
        raise ValueError("`id` column is not the row number: the feature table (built in annotation order) is not "
                         "row-aligned with this table; re-export it sorted by id or rebuild the features")
