"""CPU tests of the oracle itself: golden fixtures, independent implementations, semantics."""
import numpy as np
import pytest
import torch

from oracle import cama_context as cc
from oracle import compare, flat_search as fs


def _bf16(a):
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)


# ---- retrieval ---------------------------------------------------------------------------------
def test_retrieval_golden_regression(golden_dir):
    z = np.load(golden_dir / "retrieval_small.npz")
    k = int(z["k"])
    for metric in fs.METRICS:
        d, i = fs.flat_search(z["db"], z["queries"], k, metric)
        np.testing.assert_array_equal(i, z[f"{metric}_idx"])
        np.testing.assert_array_equal(d, z[f"{metric}_dist"])
    d, i = fs.flat_search(z["db"], z["queries"], k, "l2", z["row_group"], z["exclude_group"], prefilter=False)
    np.testing.assert_array_equal(i, z["l2_post_idx"])
    d, i = fs.flat_search(z["db"], z["queries"], k, "l2", z["row_group"], z["exclude_group"], prefilter=True)
    np.testing.assert_array_equal(i, z["l2_pre_idx"])


def test_oracle_matches_sklearn_bruteforce(golden_dir):
    """Independent third-party flat kNN (scikit-learn) agrees with the restatement."""
    from sklearn.neighbors import NearestNeighbors
    z = np.load(golden_dir / "retrieval_small.npz")
    db, q, k = z["db"].astype(np.float64), z["queries"].astype(np.float64), int(z["k"])
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="sqeuclidean").fit(db)
    dist, idx = nn.kneighbors(q)
    d, i = fs.flat_search(z["db"], z["queries"], k, "l2")
    rep = compare.check_retrieval(dist, idx, d, i, z["db"], z["queries"], "l2", rtol=1e-5)
    # duplicates (rows 3, 7, 1500) may come back in another order from sklearn: near-ties only
    assert rep["index_mismatches"] == rep["near_tie_positions"]
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(db)
    dist, idx = nn.kneighbors(q)
    d, i = fs.flat_search(z["db"], z["queries"], k, "cosine")
    compare.check_retrieval(dist, idx, d, i, z["db"], z["queries"], "cosine", rtol=1e-4)


def test_oracle_matches_python_loops():
    rng = np.random.default_rng(0)
    db = fs.normalise_rows(rng.standard_normal((60, 16)).astype(np.float32))
    q = rng.standard_normal((5, 16)).astype(np.float32) * 3
    d, i = fs.flat_search(db, q, 7, "l2")
    for qi in range(5):
        exact = [sum((float(q[qi, c]) - float(db[r, c])) ** 2 for c in range(16)) for r in range(60)]
        order = sorted(range(60), key=lambda r: (np.float32(exact[r]), r))[:7]
        assert list(i[qi]) == order
        np.testing.assert_allclose(d[qi], [exact[r] for r in order], rtol=1e-6)
    np.testing.assert_allclose(fs.distances(db, q, "l2"), fs.distances_direct_l2(db, q), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(fs.distances(db, q, "l2", accumulate="f32"), fs.distances(db, q, "l2"),
                               rtol=1e-4, atol=1e-5)


def test_ties_resolve_to_lowest_index_and_zero_distance():
    db = fs.normalise_rows(np.random.default_rng(1).standard_normal((50, 8)).astype(np.float32))
    db[10] = db[4]
    db[30] = db[4]
    d, i = fs.flat_search(db, db[4][None], 3, "l2")
    assert list(i[0]) == [4, 10, 30] and float(d[0, 0]) == 0.0


def test_l2_ranking_equals_cosine_and_dot_ranking_on_unit_rows():
    """unit-norm rows make squared-L2, cosine and dot rank identically."""
    rng = np.random.default_rng(2)
    db = fs.normalise_rows(rng.standard_normal((3000, 64)).astype(np.float32))
    q = rng.standard_normal((8, 64)).astype(np.float32) * 9
    _, i_l2 = fs.flat_search(db, q, 12, "l2")
    _, i_cos = fs.flat_search(db, q, 12, "cosine")
    _, i_dot = fs.flat_search(db, q, 12, "dot")
    np.testing.assert_array_equal(i_l2, i_cos)
    np.testing.assert_array_equal(i_l2, i_dot)


def test_post_filter_can_return_fewer_rows_prefilter_cannot():
    rng = np.random.default_rng(3)
    db = fs.normalise_rows(rng.standard_normal((200, 32)).astype(np.float32))
    groups = np.arange(200) // 3
    q = db[30][None] * 7
    d, i = fs.flat_search(db, q, 5, "l2", groups, np.array([10]), prefilter=False)
    kept = i[0][i[0] >= 0]
    assert 0 < kept.size < 5 and 30 not in kept and np.all(groups[kept] != 10)
    assert np.all(np.isinf(d[0][kept.size:]))
    d2, i2 = fs.flat_search(db, q, 5, "l2", groups, np.array([10]), prefilter=True)
    assert np.all(i2[0] >= 0) and np.all(groups[i2[0]] != 10)
    assert list(i2[0][:kept.size]) == list(kept)


def test_k_larger_than_table_and_empty_exclusion():
    db = fs.normalise_rows(np.random.default_rng(4).standard_normal((5, 8)).astype(np.float32))
    d, i = fs.flat_search(db, db[:2] * 2, 8, "l2", np.arange(5), np.array([-1, 0]))
    assert (i[0] >= 0).sum() == 5 and (i[1] >= 0).sum() == 4 and np.all(i[:, 5:] == -1)


def test_oracle_rag_database_record_schema():
    """Return schema of RAGDatabase.text_search (rag.py:54-61) as prepare_annotations uses it
    (datamodule.py:231-236): select keys + _distance, ascending, own video dropped."""
    rng = np.random.default_rng(5)
    n = 90
    cols = {"text_embedding": fs.normalise_rows(rng.standard_normal((n, 16)).astype(np.float32)),
            "video": np.array([f"v{j // 3}.mp4" for j in range(n)]),
            "start_sec": np.arange(n, dtype=np.float64), "end_sec": np.arange(n, dtype=np.float64) + 2}
    db = fs.OracleRAGDatabase(cols)
    recs = db.text_search(cols["text_embedding"][12] * 3, top_k=12, where='video != "v4.mp4"',
                          select=["video", "start_sec", "end_sec"])
    assert 0 < len(recs) <= 12
    assert all(set(r) == {"video", "start_sec", "end_sec", "_distance"} for r in recs)
    assert all(r["video"] != "v4.mp4" for r in recs)
    assert [r["_distance"] for r in recs] == sorted(r["_distance"] for r in recs)
    with pytest.raises(ValueError):
        db.text_search(cols["text_embedding"][0], where="start_sec >> ) 3")
    later = db.text_search(cols["text_embedding"][12] * 3, top_k=12, where="start_sec > 40 AND video != 'v20.mp4'")
    assert all(r["start_sec"] > 40 and r["video"] != "v20.mp4" for r in later)
    with pytest.raises(ValueError):
        fs.OracleRAGDatabase.format_result([], "csv")


def test_comparator_accepts_near_ties_and_rejects_real_errors():
    rng = np.random.default_rng(6)
    db = fs.normalise_rows(rng.standard_normal((500, 32)).astype(np.float32))
    q = rng.standard_normal((4, 32)).astype(np.float32)
    d, i = fs.flat_search(db, q, 6, "l2")
    rep = compare.check_retrieval(d, i, d, i, db, q)
    assert rep["index_mismatches"] == 0
    bad_i = i.copy()
    bad_i[0, 0] = int(np.setdiff1d(np.arange(500), i[0])[0])
    bad_d = d.copy()
    bad_d[0, 0] = compare.pair_distances(db, q, bad_i, "l2")[0, 0]
    with pytest.raises(AssertionError):
        compare.check_retrieval(bad_d, bad_i, d, i, db, q)
    with pytest.raises(AssertionError):
        compare.check_retrieval(d * 1.01, i, d, i, db, q)


# ---- CAMA context: pinned to the reference's own ActionTransformer ------------------------------
@pytest.mark.parametrize("name,dt", [("bf16", torch.bfloat16), ("f32", torch.float32)])
def test_context_restatement_equals_reference_capture(golden_dir, name, dt):
    z = np.load(golden_dir / f"cama_context_{name}.npz")
    t = (lambda k: _bf16(z[k])) if dt == torch.bfloat16 else (lambda k: torch.from_numpy(z[k]))
    x = cc.context_restatement(t("ref_feats"), t("sos"), torch.from_numpy(z["pos_table"]), t("cond"))
    assert x.dtype == dt and torch.equal(x, t("x"))
    b, K, L, C = t("ref_feats").shape
    assert torch.equal(torch.from_numpy(z["mask"]), cc.block_causal_mask(K + 1, L))
    assert torch.equal(torch.from_numpy(z["pos_table"])[0], cc.sinusoid_table(256, C)[0])
    # layout contract: group 0 = sos, group g = reference of similarity rank K-g
    bare = cc.context_restatement(t("ref_feats"), t("sos"), None, None)
    assert torch.equal(bare[:, :L], t("sos").expand(b, -1, -1))
    for g in range(1, K + 1):
        assert torch.equal(bare[:, g * L:(g + 1) * L], t("ref_feats")[:, K - g])


def test_mask_matches_reference_probe_values():
    m = cc.block_causal_mask(10, 25)                # SURVEY appendix A observations
    assert m.shape == (250, 250) and not m[25, :50].any() and bool(m[25, 50])


def test_gather_restatement_uncond_for_missing():
    table = torch.arange(6 * 2 * 3, dtype=torch.float32).view(6, 2, 3)
    un = -torch.ones(2, 3)
    out = cc.gather_restatement(table, torch.tensor([[5, -1, 0]]), un)
    assert torch.equal(out[0, 0], table[5]) and torch.equal(out[0, 1], un) and torch.equal(out[0, 2], table[0])


def test_oracle_class_matches_the_reference_class_recording(golden_dir):
    """OracleRAGDatabase (the restated class) == the reference's own class on the same engine."""
    from oracle import flat_search as fs
    cache = {}
    from oracle import compare
    n = compare.replay_reference_class(lambda t: cache.setdefault(id(t), fs.OracleRAGDatabase(t)),
                                       golden_dir / "rag_reference_class.json")
    assert n >= 15
