"""ORACLE (test infrastructure) — a minimal stand-in for the `lancedb` package, so that the
reference's REAL `RAGDatabase` class (src/data/rag.py) can be imported and driven in the build
container, where LanceDB itself is not installable.

Only the surface that class touches exists (src/data/rag.py:13-15, 26-32, 54-59, 124-128):
connect / open_table / create_table(data=pyarrow.Table) / drop_table, Table.search(vector, column)
-> a query builder with limit / nprobes / refine_factor / where / select and to_pandas / to_arrow /
to_list, Table.embedding_functions[...].function.device. The ENGINE behind search() is this
repo's own restatement of LanceDB 0.14's flat search (oracle/flat_search.py: squared-L2 default,
`.where` as a post-filter on the k nearest, ties -> lowest row), so golden vectors produced through
this module pin the reference class's argument plumbing, column selection, two-stage logic and
result formatting — NOT LanceDB's arithmetic, which stays "parity unpinned".
"""
from __future__ import annotations

import sys
import types
from types import SimpleNamespace

import numpy as np

from . import flat_search as fs

_DBS: dict[str, "FakeDB"] = {}


class FakeQuery:
    def __init__(self, table: "FakeTable", vector, column: str | None):
        self.table, self.column = table, column or table.default_vector_column
        self.vector = np.asarray(vector.detach().cpu().numpy() if hasattr(vector, "detach") else vector,
                                 dtype=np.float32)
        self.k, self._where, self._select = 10, None, None

    def limit(self, k):
        self.k = int(k)
        return self

    def nprobes(self, n):      # exact flat search: ignored, as LanceDB does without an index
        return self

    def refine_factor(self, r):
        return self

    def where(self, expr, prefilter: bool = False):
        self._where, self._prefilter = expr, prefilter
        return self

    def select(self, cols):
        self._select = list(cols)
        return self

    def _rows(self) -> list[dict]:
        t = self.table
        row_group = exclude = None
        if self._where is not None:
            row_group = (~fs.where_mask(t.columns, self._where)).astype(np.int64)
            exclude = np.array([1])
        dist, idx = fs.flat_search(np.asarray(t.columns[self.column], dtype=np.float32), self.vector[None], self.k,
                                   "l2", row_group, exclude, getattr(self, "_prefilter", False))
        cols = self._select if self._select is not None else list(t.columns)
        out = []
        for d, i in zip(dist[0], idx[0]):
            if i < 0:
                continue
            r = {c: t.columns[c][i] for c in cols}
            r["_distance"] = np.float32(d)
            out.append(r)
        return out

    def to_list(self):
        return [{k: (v.tolist() if isinstance(v, np.ndarray) else (v.item() if isinstance(v, np.generic) else v))
                 for k, v in r.items()} for r in self._rows()]

    def to_arrow(self):
        import pyarrow as pa
        rows = self._rows()
        if not rows:
            return pa.table({})
        cols = {}
        for k in rows[0]:
            vals = [r[k] for r in rows]
            if isinstance(vals[0], np.ndarray):
                cols[k] = pa.FixedSizeListArray.from_arrays(pa.array(np.concatenate(vals), type=pa.float32()),
                                                            len(vals[0]))
            else:
                cols[k] = pa.array(np.asarray(vals))
        return pa.table(cols)

    def to_pandas(self):
        return self.to_arrow().to_pandas()


class FakeTable:
    default_vector_column = "text_embedding"

    def __init__(self, columns: dict):
        self.columns = {k: (np.asarray(v) if not isinstance(v, np.ndarray) else v) for k, v in columns.items()}
        self.embedding_functions = {"text_embedding": SimpleNamespace(function=SimpleNamespace(device=None))}

    @classmethod
    def from_arrow(cls, data) -> "FakeTable":
        import pyarrow as pa
        cols = {}
        for name in data.column_names:
            col = data.column(name).combine_chunks()
            if pa.types.is_fixed_size_list(col.type):
                cols[name] = np.asarray(col.flatten().to_numpy(zero_copy_only=False),
                                        dtype=np.float32).reshape(len(col), col.type.list_size)
            else:
                cols[name] = np.asarray(col.to_numpy(zero_copy_only=False))
        return cls(cols)

    def search(self, vector, vector_column_name=None):
        return FakeQuery(self, vector, vector_column_name)


class FakeDB:
    def __init__(self):
        self.tables: dict[str, FakeTable] = {}

    def open_table(self, name):
        return self.tables[name]

    def create_table(self, name, data=None):
        self.tables[name] = FakeTable.from_arrow(data)
        return self.tables[name]

    def drop_table(self, name):
        del self.tables[name]


def connect(path: str) -> FakeDB:
    return _DBS.setdefault(str(path), FakeDB())


def install() -> None:
    """Register this module as `lancedb` (+ the two submodules rag.py imports names from)."""
    mod = types.ModuleType("lancedb")
    mod.connect = connect
    table = types.ModuleType("lancedb.table")
    table.Table = FakeTable
    query = types.ModuleType("lancedb.query")
    query.LanceQueryBuilder = FakeQuery
    mod.table, mod.query = table, query
    sys.modules.update({"lancedb": mod, "lancedb.table": table, "lancedb.query": query})


def reference_rag_database(columns: dict, reference_root: str = "/root/reference"):
    """An instance of the reference's own RAGDatabase (src/data/rag.py, imported unmodified from
    `reference_root`) over an in-memory table served by this stand-in."""
    install()
    # load the file itself: `import src.data` would pull in the Lightning datamodule (not installed)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_reference_rag", f"{reference_root}/src/data/rag.py")
    rag = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rag)
    path = f"mem://{id(columns)}"
    connect(path).tables["motion_caption"] = FakeTable(columns)
    return rag.RAGDatabase(path, "motion_caption", "cpu")
