"""`RAGDatabase` drop-in against the oracle's restatement of the reference class."""
import numpy as np
import pytest
import torch

from oracle import flat_search as fs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def table():
    rng = np.random.default_rng(0)
    n, dim = 6000, 768
    emb = fs.normalise_rows(rng.standard_normal((n, dim)).astype(np.float32))
    return {"text": np.array([f"caption {j}" for j in range(n)]), "text_embedding": emb,
            "image_embedding": fs.normalise_rows(rng.standard_normal((n, 768)).astype(np.float32)),
            "id": np.arange(n), "uid": np.array([f"u{j}" for j in range(n)]),
            "dataset": np.array(["openvid"] * n), "video": np.array([f"clip_{j // 3:05d}.mp4" for j in range(n)]),
            "start_sec": (np.arange(n) % 3) * 2.0, "end_sec": (np.arange(n) % 3) * 2.0 + 2.0}


def _same(a, b):
    assert len(a) == len(b)
    for ra, rb in zip(a, b):
        assert set(ra) == set(rb)
        for key in ra:
            if key == "_distance":
                assert ra[key] == pytest.approx(rb[key], rel=1e-3, abs=1e-6)
            elif isinstance(ra[key], np.ndarray):
                np.testing.assert_array_equal(ra[key], rb[key])
            else:
                assert ra[key] == rb[key]


def test_text_search_matches_reference_call_pattern(libmrag, table):
    """The exact kwargs prepare_annotations builds (src/data/datamodule.py:231-236)."""
    from motionrag_b200 import RAGDatabase
    db = RAGDatabase(None, None, 'cuda', columns=table)
    ora = fs.OracleRAGDatabase(table)
    rng = np.random.default_rng(1)
    for j in rng.integers(0, 6000, 6):
        q = (table["text_embedding"][j] + 0.01 * rng.standard_normal(768)).astype(np.float32) * 8
        kw = dict(text=q, top_k=9 + 3, where=f'video != "{table["video"][j]}"', select=['video', 'start_sec', 'end_sec'])
        got, want = db.text_search(**kw), ora.text_search(**kw)
        _same(got, want)
        assert 9 <= len(got) <= 12 and all(r["video"] != table["video"][j] for r in got)
    # defaults: top_k=10, all columns + vector + _distance, torch / CUDA tensors accepted
    full = db.text_search(torch.from_numpy(table["text_embedding"][5]))
    assert len(full) == 10 and full[0]["id"] == 5 and set(full[0]) == set(table) | {"_distance"}
    assert db.text_search(torch.from_numpy(table["text_embedding"][5]).cuda().half(), top_k=3)[0]["id"] == 5
    assert db.vector_search(table["text_embedding"][5], "text_embedding", output_format="pandas").shape[0] == 10
    assert db.image_search(table["image_embedding"][77], top_k=4)[0]["id"] == 77
    with pytest.raises(ValueError, match="Invalid format"):
        db.text_search(q, output_format="csv")
    with pytest.raises(ValueError, match="where clause"):
        db.text_search(q, where="start_sec >> 1")
    with pytest.raises(ValueError, match="unknown column"):
        db.text_search(q, where="nope > 1")
    with pytest.raises(NotImplementedError):
        db.text_search("a person pours water")
    db2 = RAGDatabase(None, None, columns=table, embed_fn=lambda s: table["text_embedding"][42])
    assert db2.text_search("anything", top_k=1)[0]["id"] == 42


def test_batched_fast_path_equals_per_query_calls(libmrag, table):
    from motionrag_b200 import RAGDatabase
    db = RAGDatabase(None, None, 'cuda', columns=table)
    rng = np.random.default_rng(2)
    src = rng.integers(0, 6000, 300)
    annos = [{"video": table["video"][j], "text_embedding": (table["text_embedding"][j] * 6).astype(np.float32)}
             for j in src]
    out = db.retrieve_for_annotations([dict(a) for a in annos], ref_video_num=9, batch=128)
    ora = fs.OracleRAGDatabase(table)
    for a, o in list(zip(annos, out))[::25]:
        want = ora.text_search(a["text_embedding"], top_k=12, where=f'video != "{a["video"]}"',
                               select=['video', 'start_sec', 'end_sec'])
        _same(o["ref_videos"], want)
        single = db.text_search(a["text_embedding"], top_k=12, where=f'video != "{a["video"]}"',
                                select=['video', 'start_sec', 'end_sec'])
        _same(o["ref_videos"], single)


def test_annotation_cache_file_and_text_image_branch(libmrag, table, tmp_path):
    """prepare_annotations writes the annotated list with torch.save (datamodule.py:268); the
    `rag_text_image` branch uses top_k = (2K+3, K) per annotation (:239-245)."""
    from motionrag_b200 import RAGDatabase
    db = RAGDatabase(None, None, 'cuda', columns=table)
    annos = [{"video": table["video"][j], "text_embedding": table["text_embedding"][j] * 4,
              "image_embedding": table["image_embedding"][j] * 2} for j in (5, 50, 500)]
    out = db.retrieve_for_annotations([dict(a) for a in annos], 9, save_path=tmp_path / "annos.pt")
    cached = torch.load(tmp_path / "annos.pt", weights_only=False)
    assert [a["ref_videos"] for a in cached] == [a["ref_videos"] for a in out]
    out2 = db.retrieve_for_annotations([dict(a) for a in annos], 4, ref_video_type="rag_text_image")
    for a, o in zip(annos, out2):
        want = db.text_image_search(a["text_embedding"], a["image_embedding"], top_k=(11, 4),
                                    where=f'video != "{a["video"]}"', select=['video', 'start_sec', 'end_sec'])
        assert o["ref_videos"] == want and len(want) == 4 and all(r["video"] != a["video"] for r in want)
    with pytest.raises(ValueError, match="Invalid ref_video_type"):
        db.retrieve_for_annotations(annos, 9, ref_video_type="gt")


def test_text_image_two_stage(libmrag, table):
    """src/data/rag.py:101-130 / datamodule.py:239-245: text top-(2K+3), then image top-K."""
    from motionrag_b200 import RAGDatabase
    db = RAGDatabase(None, None, 'cuda', columns=table)
    q_t = table["text_embedding"][100] * 3
    q_i = table["image_embedding"][100] * 2
    got = db.text_image_search(q_t, q_i, top_k=(21, 9), select=['video', 'id'])
    _, i0 = fs.flat_search(table["text_embedding"], q_t[None], 21)
    cand = i0[0]
    d1, i1 = fs.flat_search(table["image_embedding"][cand], q_i[None], 9)
    assert [r["id"] for r in got] == cand[i1[0]].tolist()
    np.testing.assert_allclose([r["_distance"] for r in got], d1[0], rtol=1e-3)


def test_text_image_batch_equals_single_calls_and_oracle(libmrag, table):
    """Batched two-stage search (one text scan + mrag_rescore_rows) == per-query calls == the oracle class."""
    from motionrag_b200 import RAGDatabase
    db = RAGDatabase(None, None, 'cuda', columns=table)
    ora = fs.OracleRAGDatabase(table)
    rows = [3, 250, 1999, 4000, 5999]
    texts = np.stack([table["text_embedding"][j] * 5 for j in rows]).astype(np.float32)
    imgs = np.stack([table["image_embedding"][(j * 7) % 6000] * 2 for j in rows]).astype(np.float32)
    wheres = [f'video != "{table["video"][j]}"' for j in rows]
    got = db.text_image_search_batch(texts, imgs, top_k=(21, 9), where=wheres, select=['id', 'video'])
    for j, t, im, w, g in zip(rows, texts, imgs, wheres, got):
        _same(g, db.text_image_search(t, im, top_k=(21, 9), where=w, select=['id', 'video']))
        _same(g, ora.text_image_search(t, im, top_k=(21, 9), where=w, select=['id', 'video']))
        assert len(g) == 9 and all(r["video"] != table["video"][j] for r in g)
    # raw kernel: candidates with holes, k_out larger than the candidate count
    st = db._store("image_embedding")
    cand = torch.tensor([[5, -1, 17, 5999, -1, 42]], dtype=torch.int64, device="cuda")
    q = torch.from_numpy(imgs[:1]).cuda()
    d, i = st.rescore(q, cand, 8)
    want = sorted(((float(((imgs[0] - table["image_embedding"][r]) ** 2).sum()), p, r)
                   for p, r in enumerate([5, -1, 17, 5999, -1, 42]) if r >= 0))
    assert i[0].tolist() == [r for _, _, r in want] + [-1] * 4
    np.testing.assert_allclose(d[0, :4].cpu().numpy(), [x for x, _, _ in want], rtol=1e-5)
    assert torch.isinf(d[0, 4:]).all()


def test_on_disk_table_and_pickle_roundtrip(libmrag, table, tmp_path):
    import pickle
    from motionrag_b200 import RAGDatabase, save_table
    small = {k: v[:900] for k, v in table.items()}
    save_table(tmp_path / "rag.db", "motion_caption", small)
    db = RAGDatabase(str(tmp_path / "rag.db"), "motion_caption", 'cuda')
    r = db.text_search(small["text_embedding"][10], top_k=3, select=["id"])
    db2 = pickle.loads(pickle.dumps(db.text_search)).__self__      # what the spawn pool does
    assert [x["id"] for x in db2.text_search(small["text_embedding"][10], top_k=3, select=["id"])] == [x["id"] for x in r]
    with pytest.raises(TypeError):
        pickle.dumps(RAGDatabase(None, None, columns=small))


def test_drop_in_class_matches_the_reference_class_recording(libmrag, golden_dir):
    """motionrag_b200.RAGDatabase replays every call recorded from the reference's REAL RAGDatabase
    class (src/data/rag.py run over oracle/fake_lancedb.py): same rows in the same order, same keys,
    same container type per output_format, distances within 1e-3, same ValueError."""
    from motionrag_b200 import RAGDatabase
    from oracle import compare
    cache = {}
    n = compare.replay_reference_class(lambda t: cache.setdefault(id(t), RAGDatabase(None, None, 'cuda', columns=t)),
                                       golden_dir / "rag_reference_class.json", rel=1e-3)
    assert n >= 15


GENERAL_WHERE = ["start_sec > 1", "start_sec >= 2 AND video != 'clip_00004.mp4'", "id < 3000 or dataset = 'webvid'",
                 "video LIKE 'clip_000%' and not (start_sec = 0)", "uid in ('u12', 'u13', 'u14', 'u4000') or id between 20 and 900",
                 "video is not null and end_sec <= 4", "dataset != 'openvid'"]


@pytest.mark.parametrize("prefilter", [False, True])
def test_general_where_clauses_match_the_oracle(libmrag, table, prefilter):
    """Any predicate of the SQL subset (src/data/rag.py:56-57 forwards the string to LanceDB): post-filter on the
    k nearest (default) and pre-filter, against the oracle whose predicate engine is SQLite."""
    from motionrag_b200 import RAGDatabase
    db = RAGDatabase(None, None, 'cuda', columns=table, prefilter=prefilter)
    ora = fs.OracleRAGDatabase(table, prefilter=prefilter)
    rng = np.random.default_rng(11)
    for w in GENERAL_WHERE:
        for j in rng.integers(0, 6000, 3):
            q = (table["text_embedding"][j] + 0.02 * rng.standard_normal(768)).astype(np.float32) * 5
            kw = dict(text=q, top_k=12, where=w, select=['id', 'video', 'start_sec'])
            got, want = db.text_search(**kw), ora.text_search(**kw)
            _same(got, want)
            if prefilter and w != "dataset != 'openvid'":
                assert len(got) == 12
    # a batch with one clause per query: `!=` clauses run on the device, the rest on the host, None = no filter
    if not prefilter:
        src = rng.integers(0, 6000, 40)
        qs = (table["text_embedding"][src] * 4).astype(np.float32)
        wheres = [None if i % 5 == 0 else (f'video != "{table["video"][j]}"' if i % 2 else GENERAL_WHERE[i % len(GENERAL_WHERE)])
                  for i, j in enumerate(src)]
        out = db.search_batch(qs, top_k=12, where=wheres, select=['id', 'video', 'start_sec'])
        for q, w, o in zip(qs, wheres, out):
            _same(o, ora.text_search(q, top_k=12, where=w, select=['id', 'video', 'start_sec']))


@pytest.mark.parametrize("prefilter", [False, True])
def test_top_k_above_the_kernels_32_runs_in_passes_and_matches_the_oracle(libmrag, table, prefilter):
    """LanceDB takes any `limit` (src/data/rag.py:54 forwards top_k); above 32 results the drop-in collects the k
    nearest rows in passes of 32: same rows, order and distances as the oracle — with where clauses (post- and
    pre-filter), duplicate rows across a pass border, a batch, and more results asked for than rows exist."""
    from motionrag_b200 import RAGDatabase
    t = dict(table)
    emb = table["text_embedding"].copy()
    emb[[40, 41, 42, 43]] = emb[39]                   # five identical rows: ties straddle the 32-result border below
    t["text_embedding"] = emb
    db = RAGDatabase(None, None, 'cuda', columns=t, prefilter=prefilter)
    ora = fs.OracleRAGDatabase(t, prefilter=prefilter)
    rng = np.random.default_rng(21)
    sel = ['id', 'video', 'start_sec']
    for k, w in [(33, None), (50, 'video != "clip_00013.mp4"'), (100, "start_sec >= 2"), (64, "id >= 3000 or start_sec = 0")]:
        for j in (39, int(rng.integers(0, 6000))):
            q = (emb[j] + 0.02 * rng.standard_normal(768)).astype(np.float32) * 3
            got, want = db.text_search(q, top_k=k, where=w, select=sel), ora.text_search(q, top_k=k, where=w, select=sel)
            _same(got, want)
            assert len(got) == k if (w is None or prefilter) else len(got) <= k
    # the duplicates sit inside the first 40 results in ascending row order
    ids = [r["id"] for r in db.text_search(emb[39] * 2, top_k=40, select=['id'])]
    assert ids[:5] == [39, 40, 41, 42, 43]
    # an ordinary call afterwards binds its own labelling again
    kw = dict(text=emb[7] * 2, top_k=12, where=f'video != "{t["video"][7]}"', select=sel)
    _same(db.text_search(**kw), ora.text_search(**kw))
    # batch form
    qs = (emb[[5, 600, 4242]] * 2).astype(np.float32)
    for q, o in zip(qs, db.search_batch(qs, top_k=45, where="end_sec <= 4", select=sel)):
        _same(o, ora.text_search(q, top_k=45, where="end_sec <= 4", select=sel))
    # more than the table holds
    small = {c: v[:50] for c, v in t.items()}
    got = RAGDatabase(None, None, 'cuda', columns=small).text_search(emb[3], top_k=80, select=['id'])
    _same(got, fs.OracleRAGDatabase(small).text_search(emb[3], top_k=80, select=['id']))
    assert len(got) == 50
    # two-stage search with ref_video_num = 20: k0 = 2 * 20 + 3 = 43 text hits feed the image stage (datamodule.py:241)
    if not prefilter:
        kw = dict(text=emb[100] * 2, image_embedding=table["image_embedding"][100] * 2, top_k=(43, 20),
                  where=f'video != "{t["video"][100]}"', select=sel)
        _same(db.text_image_search(**kw), ora.text_image_search(**kw))


@pytest.mark.parametrize("container", ["parquet", "arrow", "fragments"])
def test_arrow_dump_of_the_reference_table_opens_and_matches_the_oracle(libmrag, table, tmp_path, container):
    """A pyarrow-written table with the schema tools/build_rag_database.py:35-45 produces (FixedSizeList<f32>[768]
    embedding + scalar columns; bf16-normalised rows like the reference's bfloat16 embedder writes) opens through
    RAGDatabase(db_path, table_name) — no lancedb — pickles by path, and answers like the oracle on the same rows."""
    import pickle

    import pyarrow as pa
    import pyarrow.feather as pf
    import pyarrow.parquet as pq
    from motionrag_b200 import RAGDatabase
    n = 1500
    emb = torch.from_numpy(table["text_embedding"][:n] * 1.7).bfloat16()
    emb = (emb / emb.float().norm(dim=-1, keepdim=True).bfloat16()).float().numpy()   # unit only to bf16 precision
    t = pa.table({"text": table["text"][:n].tolist(),
                  "text_embedding": pa.FixedSizeListArray.from_arrays(pa.array(emb.reshape(-1), type=pa.float32()), 768),
                  "id": np.arange(n), "uid": table["uid"][:n].tolist(), "dataset": table["dataset"][:n].tolist(),
                  "video": table["video"][:n].tolist(), "start_sec": table["start_sec"][:n], "end_sec": table["end_sec"][:n]})
    root = tmp_path / "openvid.db"
    root.mkdir()
    if container == "parquet":
        pq.write_table(t, root / "motion_caption.parquet")
    elif container == "arrow":
        pf.write_feather(t, root / "motion_caption.arrow", compression="uncompressed")
    else:
        (root / "motion_caption").mkdir()
        for i, s in enumerate(range(0, n, 400)):
            pq.write_table(t.slice(s, 400), root / "motion_caption" / f"part-{i:03d}.parquet")
    db = RAGDatabase(str(root), "motion_caption", 'cuda')
    assert len(db) == n
    cols = {name: np.asarray(t[name].to_pylist()) for name in t.column_names if name != "text_embedding"}
    cols["text_embedding"] = emb
    ora = fs.OracleRAGDatabase(cols)
    rng = np.random.default_rng(21)
    for j in rng.integers(0, n, 5):
        q = (emb[j] + 0.02 * rng.standard_normal(768)).astype(np.float32) * 7
        kw = dict(text=q, top_k=12, where=f'video != "{cols["video"][j]}"', select=['video', 'start_sec', 'end_sec'])
        _same(db.text_search(**kw), ora.text_search(**kw))
    again = pickle.loads(pickle.dumps(db.text_search)).__self__
    _same(again.text_search(q, top_k=5, select=["id"]), ora.text_search(q, top_k=5, select=["id"]))
