"""ORACLE (test infrastructure) — how the "parity unpinned" gap of oracle/flat_search.py gets closed by anyone who
has the reference's engine (`lancedb==0.14.0`, reference requirements.txt:18): a fixed, seeded set of cases is run
through the REAL `lancedb` exactly as the reference drives it (src/data/rag.py:54-59:
`table.search(vector, column).limit(k).nprobes(n).refine_factor(r)[.where(w)][.select(cols)]`) and the answers are
stored as golden vectors (tests/golden/lancedb_golden.json, written by tools/make_lancedb_golden.py). The replay in
tests/test_lancedb_golden.py then holds the oracle (CPU) and the CUDA drop-in (GPU) to them. Without the wheel nothing
here runs LanceDB; the cases and the replay still work, the golden file is simply absent and the tests skip.

Each case probes one assumption SURVEY §8(c) lists as unverifiable in the build container:
  default_metric     (i)   no metric given -> squared L2, `_distance = sum((q - d)^2)`, un-normalised queries
  post_filter        (ii)  `.where(expr)` without prefilter drops rows AFTER the k nearest were chosen (< k rows back)
  general_where      (ii)  a predicate beyond the reference's `video != "x"`
  flat_ignores_knobs (iii) nprobes / refine_factor change nothing on an un-indexed table
  tie_order          (iv)  duplicate rows: which of two identical vectors comes first
  non_unit_rows      (v)   rows that are unit only to bf16 precision, and a zero row (`on_bad_vectors='fill'`)
  select_columns           `select=[...]` -> exactly those keys + `_distance`

Tables are generated with numpy's PCG64 (bit-identical on every platform), never stored.
"""
from __future__ import annotations

import json

import numpy as np

DIM = 768
GOLDEN_NAME = "lancedb_golden.json"


def case_table(seed: int, n: int, kind: str) -> dict:
    """Columns of tools/build_rag_database.py:35-45 for a seeded synthetic table."""
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((64, DIM)).astype(np.float32)
    cent /= np.linalg.norm(cent, axis=-1, keepdims=True)
    emb = cent[rng.integers(0, 64, n)] + (0.3 / np.sqrt(DIM)) * rng.standard_normal((n, DIM)).astype(np.float32)
    emb = (emb / np.linalg.norm(emb, axis=-1, keepdims=True)).astype(np.float32)
    if kind == "ties":            # identical vectors at different rows
        emb[1500], emb[7] = emb[3], emb[3]
        emb[900] = emb[901]
    elif kind == "non_unit":      # unit to bf16 precision only (bfloat16 embedder), one zero-filled row
        import torch
        t = torch.from_numpy(emb * 1.3).bfloat16()
        emb = (t / t.float().norm(dim=-1, keepdim=True).bfloat16()).float().numpy()
        emb[11] = 0.0
    return {"text": np.array([f"caption {j}" for j in range(n)], dtype=object), "text_embedding": np.ascontiguousarray(emb),
            "id": np.arange(n, dtype=np.int64), "uid": np.array([f"u{j}" for j in range(n)], dtype=object),
            "dataset": np.array(["openvid" if j % 5 else "webvid" for j in range(n)], dtype=object),
            "video": np.array([f"video_{j // 3:06d}.mp4" for j in range(n)], dtype=object),
            "start_sec": (np.arange(n) % 3) * 2.0, "end_sec": (np.arange(n) % 3) * 2.0 + 2.0}


def case_queries(seed: int, table: dict, nq: int, rows=None) -> tuple[np.ndarray, np.ndarray]:
    """Un-normalised queries near table rows (src/data/datamodule.py:300-302 never normalises)."""
    rng = np.random.default_rng(seed + 1000)
    src = rng.integers(0, len(table["id"]), nq) if rows is None else np.asarray(rows)
    q = table["text_embedding"][src] + (0.1 / np.sqrt(DIM)) * rng.standard_normal((len(src), DIM)).astype(np.float32)
    return (q * rng.uniform(5, 15, (len(src), 1))).astype(np.float32), src


CASES = [
    {"name": "default_metric", "seed": 1, "n": 20000, "kind": "plain", "nq": 8, "k": 12},
    {"name": "post_filter", "seed": 2, "n": 20000, "kind": "plain", "nq": 8, "k": 12, "where": "own_video"},
    {"name": "general_where", "seed": 3, "n": 20000, "kind": "plain", "nq": 6, "k": 12,
     "where": "start_sec >= 2 AND dataset = 'openvid'"},
    {"name": "flat_ignores_knobs", "seed": 1, "n": 20000, "kind": "plain", "nq": 8, "k": 12, "nprobes": 1, "refine_factor": 1},
    {"name": "tie_order", "seed": 4, "n": 5000, "kind": "ties", "nq": 3, "k": 6, "rows": [3, 900, 7], "exact_rows": True},
    {"name": "non_unit_rows", "seed": 5, "n": 20000, "kind": "non_unit", "nq": 8, "k": 12},
    {"name": "select_columns", "seed": 6, "n": 5000, "kind": "plain", "nq": 2, "k": 5, "select": ["video", "start_sec", "end_sec"]},
]


def case_inputs(case: dict):
    table = case_table(case["seed"], case["n"], case["kind"])
    q, src = case_queries(case["seed"], table, case["nq"], case.get("rows"))
    if case.get("exact_rows"):          # the query IS a (duplicated) row: both copies are at distance ~0
        q = table["text_embedding"][src].copy()
    wheres = [None] * len(q)
    if case.get("where") == "own_video":
        wheres = [f'video != "{table["video"][j]}"' for j in src]
    elif case.get("where"):
        wheres = [case["where"]] * len(q)
    return table, q, wheres


def run_engine(search_one, case: dict) -> dict:
    """`search_one(table, q, k, where, select, nprobes, refine_factor) -> list[dict]` over every query of the case."""
    table, q, wheres = case_inputs(case)
    select = case.get("select", ["id", "video"])
    out = []
    for qi, w in zip(q, wheres):
        recs = search_one(table, qi, case["k"], w, select, case.get("nprobes", 50), case.get("refine_factor", 30))
        out.append({"ids": [int(r["id"]) if "id" in r else None for r in recs],
                    "videos": [str(r["video"]) for r in recs], "keys": sorted(recs[0]) if recs else [],
                    "distances": [float(r["_distance"]) for r in recs]})
    return {"case": case, "results": out}


def lancedb_search_factory(tmpdir: str):
    """search_one over the REAL lancedb: one table per case, driven line for line like src/data/rag.py:54-61."""
    import lancedb
    import pyarrow as pa
    db = lancedb.connect(tmpdir)
    made, keep = {}, []

    def search_one(table, q, k, where, select, nprobes, refine_factor):
        key = id(table)
        if key not in made:
            keep.append(table)                 # keeps id(table) unique for the life of this factory
            cols = {c: (pa.FixedSizeListArray.from_arrays(pa.array(v.reshape(-1), type=pa.float32()), v.shape[1])
                        if c == "text_embedding" else pa.array(v.tolist())) for c, v in table.items()}
            made[key] = db.create_table(f"t{len(made)}", data=pa.table(cols))
        s = made[key].search(q, vector_column_name="text_embedding").limit(k).nprobes(nprobes).refine_factor(refine_factor)
        if where is not None:
            s = s.where(where)
        if select is not None:
            s = s.select(select)
        return s.to_pandas().to_dict("records")
    return search_one


def oracle_search_one(table, q, k, where, select, nprobes, refine_factor):
    from . import flat_search as fs
    cache = oracle_search_one.__dict__.setdefault("dbs", {})
    if cache.get("table") is not table:        # keyed by identity with the table kept alive (an id() can be reused)
        cache["table"], cache["db"] = table, fs.OracleRAGDatabase(table)
    return cache["db"].text_search(q, top_k=k, where=where, select=select, nprobes=nprobes, refine_factor=refine_factor)


def compare_runs(got: dict, gold: dict, rel: float = 1e-3) -> dict:
    """The parity rule of BASELINE.md §5 between two runs of one case: same number of rows, distances within `rel`
    position by position, same keys, ids identical except where the GOLD distances of the two rows are within
    `rel` of each other (near-ties; exact ties are counted separately: tie ORDER is assumption iv)."""
    report = {"queries": len(gold["results"]), "positions": 0, "id_mismatches": 0, "near_ties": 0, "exact_tie_swaps": 0}
    for qi, (g, w) in enumerate(zip(got["results"], gold["results"])):
        assert len(g["ids"]) == len(w["ids"]), f"{gold['case']['name']} q{qi}: {len(g['ids'])} rows, engine returned {len(w['ids'])}"
        assert g["keys"] == w["keys"], (gold["case"]["name"], g["keys"], w["keys"])
        for j, (dg, dw) in enumerate(zip(g["distances"], w["distances"])):
            assert abs(dg - dw) <= rel * max(abs(dw), 1e-3), f"{gold['case']['name']} q{qi} rank {j}: distance {dg} vs {dw}"
        report["positions"] += len(w["ids"])
        ident_g, ident_w = (g["ids"], w["ids"]) if w["ids"] and w["ids"][0] is not None else (g["videos"], w["videos"])
        for j, (a, b) in enumerate(zip(ident_g, ident_w)):
            if a != b:
                report["id_mismatches"] += 1
                if w["distances"][j] == g["distances"][j] and a in ident_w and b in ident_g:
                    report["exact_tie_swaps"] += 1       # same set, same distances, different order of equals
                else:
                    report["near_ties"] += 1             # distance check above already bounded the gap by rel
    return report


def load_golden(path) -> dict | None:
    try:
        return json.loads(open(path).read())
    except FileNotFoundError:
        return None


def live_probe() -> dict:
    """Used by __graft_entry__.smoke(): is the reference's engine importable on this machine, and if so does the
    oracle agree with it on the cases above? Never raises; without the wheel it just says so."""
    try:
        import lancedb
    except Exception as e:   # noqa: BLE001
        return {"lancedb": "not importable", "why": f"{type(e).__name__}: {e}", "parity": "unpinned"}
    import tempfile
    out = {"lancedb": getattr(lancedb, "__version__", "?"), "cases": {}}
    try:
        with tempfile.TemporaryDirectory() as d:
            live = lancedb_search_factory(d)
            for case in CASES:
                small = dict(case, n=min(case["n"], 6000))
                try:
                    out["cases"][case["name"]] = compare_runs(run_engine(oracle_search_one, small), run_engine(live, small))
                except AssertionError as e:
                    out["cases"][case["name"]] = {"differs": str(e)[:300]}
        out["parity"] = "pinned" if all("differs" not in v for v in out["cases"].values()) else "oracle differs from lancedb"
    except Exception as e:   # noqa: BLE001
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    return out
