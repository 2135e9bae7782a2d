"""ORACLE (test infrastructure) — the parity rule of BASELINE.md §5 as code.

Scores within `rtol` relative of the fp32 oracle; indices identical except positions whose
ORACLE scores differ by less than that tolerance (documented near-ties: the count is
returned so every run can state it).
"""
from __future__ import annotations

import numpy as np


def pair_distances(db: np.ndarray, queries: np.ndarray, idx: np.ndarray, metric: str) -> np.ndarray:
    """float64 `_distance` of (query q, row idx[q, j]) pairs; nan where idx < 0."""
    db = np.asarray(db, dtype=np.float32)
    queries = np.asarray(queries, dtype=np.float32)
    out = np.full(idx.shape, np.nan, dtype=np.float64)
    for q in range(idx.shape[0]):
        valid = idx[q] >= 0
        if not valid.any():
            continue
        rows = db[idx[q][valid]].astype(np.float64)
        qv = queries[q].astype(np.float64)
        if metric == "l2":
            d = ((rows - qv[None]) ** 2).sum(-1)
        elif metric == "dot":
            d = 1.0 - rows @ qv
        elif metric == "cosine":
            d = 1.0 - (rows @ qv) / np.maximum(np.linalg.norm(rows, axis=-1) * np.linalg.norm(qv), 1e-30)
        else:
            raise ValueError(metric)
        out[q, valid] = d
    return out


def check_retrieval(got_dist: np.ndarray, got_idx: np.ndarray, ref_dist: np.ndarray,
                    ref_idx: np.ndarray, db: np.ndarray, queries: np.ndarray, metric: str = "l2",
                    rtol: float = 1e-3, atol: float = 1e-6) -> dict:
    """Raises AssertionError on a parity violation; returns a report otherwise.

    atol guards relative comparisons of distances that are ~0 (a query equal to a row).
    """
    got_dist = np.asarray(got_dist, dtype=np.float64)
    ref_dist = np.asarray(ref_dist, dtype=np.float64)
    got_idx = np.asarray(got_idx, dtype=np.int64)
    ref_idx = np.asarray(ref_idx, dtype=np.int64)
    assert got_idx.shape == ref_idx.shape, (got_idx.shape, ref_idx.shape)
    # 1. same number of rows per query
    n_got = (got_idx >= 0).sum(-1)
    n_ref = (ref_idx >= 0).sum(-1)
    bad = np.nonzero(n_got != n_ref)[0]
    assert bad.size == 0, f"row-count mismatch for queries {bad[:8]}: got {n_got[bad[:8]]} ref {n_ref[bad[:8]]}"
    # 2. returned scores match the exact distance of the rows actually returned
    exact = pair_distances(db, queries, got_idx, metric)
    valid = got_idx >= 0
    err = np.abs(got_dist[valid] - exact[valid]) / np.maximum(np.abs(exact[valid]), atol / rtol)
    max_rel = float(err.max()) if err.size else 0.0
    assert max_rel <= rtol, f"score error {max_rel:.3e} exceeds rtol {rtol:g}"
    # 3. ascending order, ties by index
    for q in range(got_idx.shape[0]):
        d = got_dist[q][valid[q]]
        assert np.all(np.diff(d) >= 0), f"query {q}: distances not ascending"
    # 4. indices identical up to near-ties of the ORACLE scores
    diff = (got_idx != ref_idx) & valid
    near = 0
    for q, j in zip(*np.nonzero(diff)):
        gap = abs(exact[q, j] - ref_dist[q, j])
        lim = rtol * max(abs(ref_dist[q, j]), atol / rtol)
        assert gap <= lim, (f"query {q} rank {j}: row {got_idx[q, j]} (d={exact[q, j]:.9g}) returned, "
                            f"oracle has row {ref_idx[q, j]} (d={ref_dist[q, j]:.9g}); gap {gap:.3e} > {lim:.3e}")
        near += 1
    return {"queries": int(got_idx.shape[0]), "positions": int(valid.sum()), "index_mismatches": int(diff.sum()),
            "near_tie_positions": int(near), "max_rel_score_err": max_rel}


def replay_reference_class(factory, golden_json, rel: float = 1e-5) -> int:
    """Replays tests/golden/rag_reference_class.json — recorded from the reference's REAL RAGDatabase
    class (src/data/rag.py) run over the LanceDB stand-in, see oracle/make_golden.py — against the class
    `factory(table_columns)` returns: same rows in the same order, same keys, same container type per
    output_format, `_distance` within `rel`, same ValueError. Returns the number of calls replayed."""
    import json
    import math
    from . import make_golden as mg
    gold = json.loads(open(golden_json).read())
    table = mg.rag_class_table(**gold["table"])
    for entry in gold["calls"]:
        call, want = entry["call"], entry["result"]
        kw = mg.rag_class_kwargs(table, call)
        db = factory(table)
        if "raises" in want:
            try:
                getattr(db, call["method"])(**kw)
            except ValueError as e:
                assert want["message"] in str(e), (call, str(e))
            else:
                raise AssertionError(f"{call}: expected ValueError({want['message']!r})")
            continue
        got = mg.rag_class_summary(getattr(db, call["method"])(**kw))
        assert got["kind"] == want["kind"] and got["keys"] == want["keys"], (call, got["kind"], got["keys"])
        assert len(got["records"]) == len(want["records"]), (call, len(got["records"]), len(want["records"]))
        for g, w in zip(got["records"], want["records"]):
            for key, val in w.items():
                if key == "_distance":
                    assert math.isclose(g[key], val, rel_tol=rel, abs_tol=1e-6), (call, g[key], val)
                else:
                    assert g[key] == val, (call, key, g[key], val)
    return len(gold["calls"])


def check_recall(exact_dist: np.ndarray, exact_idx: np.ndarray, approx_dist: np.ndarray, approx_idx: np.ndarray,
                 rtol: float = 1e-3, atol: float = 1e-6) -> dict:
    """Parity rule for the regime where the REFERENCE is approximate (tables > 1 M rows carry an IVF-PQ index,
    tools/build_rag_database.py:51-52; see oracle/ivf_pq.py): index-identical answers are not defined there, so an
    exact search is held to what must be true of it against ANY approximate answer over the same rows:
      1. domination: rank by rank the exact `_distance` is <= the approximate one (within rtol);
      2. agreement: a row both return carries the same `_distance` (the reference re-scores its candidates exactly,
         refine_factor=30, so distances of shared rows are comparable);
      3. nothing better was missed: a row only the approximate answer returns is no nearer than the exact answer's
         last row (within rtol).
    Inputs are one query's or a batch's (distance, index) lists, unused slots (inf / -1). Returns recall@k of the
    approximate answer against the exact one (reported, not asserted: it measures the reference, not the product)."""
    exact_dist, approx_dist = np.atleast_2d(np.asarray(exact_dist, np.float64)), np.atleast_2d(np.asarray(approx_dist, np.float64))
    exact_idx, approx_idx = np.atleast_2d(np.asarray(exact_idx, np.int64)), np.atleast_2d(np.asarray(approx_idx, np.int64))
    hits = total = 0
    for q in range(exact_idx.shape[0]):
        ei, ai = exact_idx[q][exact_idx[q] >= 0], approx_idx[q][approx_idx[q] >= 0]
        ed, ad = exact_dist[q][:len(ei)], approx_dist[q][:len(ai)]
        assert len(ei) >= len(ai), f"query {q}: exact search returned {len(ei)} rows, approximate {len(ai)}"
        tol = lambda v: rtol * max(abs(v), atol / rtol)
        for j in range(len(ai)):
            assert ed[j] <= ad[j] + tol(ad[j]), f"query {q} rank {j}: exact {ed[j]:.9g} > approximate {ad[j]:.9g}"
        where = {int(r): j for j, r in enumerate(ei)}
        for j, r in enumerate(ai):
            if int(r) in where:
                assert abs(ed[where[int(r)]] - ad[j]) <= tol(ad[j]), f"query {q}: row {r} scored {ed[where[int(r)]]} vs {ad[j]}"
                hits += 1
            elif len(ei):
                assert ad[j] >= ed[-1] - tol(ed[-1]), f"query {q}: row {r} (d={ad[j]:.9g}) beats the exact answer's last row"
        total += len(ei)
    return {"queries": int(exact_idx.shape[0]), "recall_at_k": hits / max(total, 1), "positions": int(total)}
