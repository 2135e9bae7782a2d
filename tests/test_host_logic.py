"""Host-side logic that needs no GPU: where-clause parsing, record formatting, layouts, plans."""
import numpy as np
import pytest
import torch

from motionrag_b200 import parallel, rag
from motionrag_b200.context import block_causal_mask, sinusoid_table
from oracle import cama_context as cc


def _db_no_gpu(n=30, dim=8):
    """A RAGDatabase shell with host columns only (bypasses __init__, which requires CUDA)."""
    db = object.__new__(rag.RAGDatabase)
    rng = np.random.default_rng(0)
    db._columns = {"video": np.array([f"v{j // 3}" for j in range(n)]), "id": np.arange(n),
                   "start_sec": np.arange(n, dtype=np.float64)}
    db._vectors = {"text_embedding": rng.standard_normal((n, dim)).astype(np.float32)}
    db._stores, db._retriever = {}, None
    db._init_caches()
    db.recheck, db.fp32_rechecks = "auto", 0
    return db


def test_reference_where_form_is_recognised_as_a_device_exclusion():
    from motionrag_b200.where import parse
    assert parse('video != "a b/c.mp4"').simple_exclusion() == ("video", "a b/c.mp4")
    assert parse("video != 'x'").simple_exclusion() == ("video", "x")
    for general in ('video = "a"', 'start_sec > 3', 'video != "a" AND id != "3"'):
        assert parse(general).simple_exclusion() is None


def test_group_ids_and_lookup():
    db = _db_no_gpu()
    g = db._groups_for("video")
    assert g["ids"].dtype == np.int32 and len(set(g["ids"].tolist())) == 10
    assert db._group_ids_lookup("video", "v3") == g["ids"][9]
    assert db._group_ids_lookup("video", "missing") == -1
    assert db._group_ids_lookup("id", "7") == 7          # numeric literal inside the SQL string


def test_records_schema_and_padding_rows_are_dropped():
    db = _db_no_gpu()
    dist = np.array([[0.1, 0.2, np.inf]], dtype=np.float32)
    idx = np.array([[4, 9, -1]])
    recs = db._records(dist, idx, ["video", "start_sec"])
    assert recs == [[{"video": "v1", "start_sec": 4.0, "_distance": pytest.approx(0.1)},
                     {"video": "v3", "start_sec": 9.0, "_distance": pytest.approx(0.2)}]]
    full = db._records(dist, idx, None)[0][0]
    assert set(full) == {"video", "id", "start_sec", "text_embedding", "_distance"}
    with pytest.raises(ValueError):
        db._records(dist, idx, ["nope"])
    # the batched builder (more than 4 queries) produces the same records as the per-query one
    idx8, dist8 = np.repeat(idx, 8, 0), np.repeat(dist, 8, 0)
    assert db._records(dist8, idx8, ["video", "start_sec"]) == recs * 8
    # columns handed over as Python lists work too
    db2 = _db_no_gpu()
    db2._columns = {k: rag._as_column(v.tolist()) for k, v in db2._columns.items()}
    assert db2._records(dist8, idx8, ["video", "start_sec"]) == recs * 8


def test_where_clauses_are_parsed_once_and_validated():
    db = _db_no_gpu()
    db.prefilter = False
    db._bind_groups = lambda col: setattr(db, "_group_col", col)      # no store in this shell
    ids, preds = db._exclusion_ids('video != "v3"', 2)
    assert ids.tolist() == [db._group_ids_lookup("video", "v3")] * 2 and 'video != "v3"' in db._where_cache and preds is None
    ids, preds = db._exclusion_ids(['video != "v1"', None, 'video != "nope"'], 3)
    assert ids.tolist() == [db._group_ids_lookup("video", "v1"), -1, -1] and preds is None
    assert db._exclusion_ids(None, 3) == (None, None) and db._exclusion_ids([None, None], 2) == (None, None)
    # anything but `col != literal` is a host predicate, applied to the rows a query returned (post-filter)
    ids, preds = db._exclusion_ids(['start_sec > 3', 'video != "v1"', None], 3)
    assert ids.tolist() == [-1, db._group_ids_lookup("video", "v1"), -1] and preds[1] is None and preds[2] is None
    dist = np.array([[0.1, 0.2, 0.3, np.inf], [0.1, 0.2, 0.3, 0.4], [0.5, 0.6, 0.7, 0.8]], dtype=np.float32)
    idx = np.array([[2, 9, 4, -1], [1, 2, 3, 4], [5, 6, 7, 8]])
    db._apply_predicates(preds, dist, idx)
    assert idx.tolist() == [[9, 4, -1, -1], [1, 2, 3, 4], [5, 6, 7, 8]]
    assert dist[0, :2].tolist() == pytest.approx([0.2, 0.3]) and np.isinf(dist[0, 2:]).all()
    with pytest.raises(ValueError, match="where clause"):
        db._exclusion_ids('start_sec >> 3', 1)
    with pytest.raises(ValueError, match="unknown column"):
        db._exclusion_ids('nope != "x"', 1)
    ids, preds = db._exclusion_ids(['video != "v1"', 'id != "3"'], 2)       # second column -> host predicate
    assert ids.tolist() == [db._group_ids_lookup("video", "v1"), -1] and preds[0] is None and preds[1].text == 'id != "3"'
    with pytest.raises(ValueError, match="one where clause per query"):
        db._exclusion_ids(['video != "v1"'], 2)
    # pre-filter: one general clause per batch becomes a two-group row mask (group 1 = rows that fail)
    db.prefilter = True
    bound = {}
    db._stores = {"text_embedding": type("S", (), {"set_groups": lambda self, g: bound.update(g=g), "__len__": lambda self: 30})()}
    ids, preds = db._exclusion_ids('start_sec > 3', 2)
    assert ids.tolist() == [1, 1] and preds is None and bound["g"].tolist() == [1] * 4 + [0] * 26
    with pytest.raises(ValueError, match="ONE general where"):
        db._exclusion_ids(['start_sec > 3', 'start_sec > 4'], 2)
    with pytest.raises(ValueError, match="same column"):
        db._exclusion_ids(['video != "v1"', 'id != "3"'], 2)


def test_format_result_formats_and_error():
    recs = [{"video": "a", "_distance": 0.5}]
    assert rag.RAGDatabase.format_result(recs, "dict") == recs
    assert rag.RAGDatabase.format_result(recs, "list") == recs
    assert list(rag.RAGDatabase.format_result(recs, "pandas").columns) == ["video", "_distance"]
    assert rag.RAGDatabase.format_result(recs, "pyarrow").num_rows == 1
    with pytest.raises(ValueError, match="Invalid format"):
        rag.RAGDatabase.format_result(recs, "csv")


def test_save_table_roundtrip(tmp_path):
    n = 12
    cols = {"text_embedding": np.random.default_rng(1).standard_normal((n, 8)).astype(np.float32),
            "video": [f"v{j}" for j in range(n)], "start_sec": np.arange(n, dtype=np.float64),
            "end_sec": np.arange(n, dtype=np.float64) + 1, "id": np.arange(n)}
    rag.save_table(tmp_path, "motion_caption", cols)
    back = rag.RAGDatabase._load(tmp_path, "motion_caption")
    np.testing.assert_array_equal(back["text_embedding"], cols["text_embedding"])
    assert list(back["video"]) == cols["video"] and back["start_sec"].dtype == np.float64


def test_shard_range_covers_table_exactly():
    for n, g in [(10, 3), (1_000_000, 8), (7, 8), (256, 2)]:
        spans = [parallel.shard_range(n, g, r) for r in range(g)]
        assert spans[0][1] == 0 and spans[-1][2] == n
        assert all(a[2] == b[1] for a, b in zip(spans, spans[1:]))
        assert all(lo == min(n, r * rps) for r, (rps, lo, hi) in enumerate(spans))


def test_packed_layout_views_alias_one_buffer():
    lay = parallel.PackedLayout(nq=3, k=4)
    buf = torch.zeros(2 * lay.nbytes, dtype=torch.uint8)
    d, g, i = lay.views(buf, 2)
    assert d.shape == g.shape == i.shape == (2, 3, 4)
    d[1, 2, 3] = 1.5
    g[0, 0, 0] = 7
    i[1, 0, 0] = 1 << 40
    d2, g2, i2 = lay.views(buf, 2)
    assert d2[1, 2, 3] == 1.5 and g2[0, 0, 0] == 7 and i2[1, 0, 0] == 1 << 40
    assert d[1].data_ptr() - d[0].data_ptr() == lay.nbytes and i[0].data_ptr() % 8 == 0


def test_mask_and_sinusoid_match_oracle():
    assert torch.equal(block_causal_mask(10, 25), cc.block_causal_mask(10, 25))
    assert torch.equal(sinusoid_table(256, 1024), cc.sinusoid_table(256, 1024)[0])


def test_select_refs_follows_get_ref_videos_semantics():
    """src/data/dataset.py:285-312: first K hits, dropped/missing -> (-1, distance 1.0)."""
    from motionrag_b200 import select_refs
    idx = torch.tensor([[5, 9, 2, 7, -1, -1], [1, -1, -1, -1, -1, -1]])
    dist = torch.tensor([[.1, .2, .3, .4, float("inf"), float("inf")], [.5] + [float("inf")] * 5])
    i, d = select_refs(idx, dist, 4)
    assert i.tolist() == [[5, 9, 2, 7], [1, -1, -1, -1]]
    assert d.tolist() == [[pytest.approx(.1), pytest.approx(.2), pytest.approx(.3), pytest.approx(.4)], [.5, 1, 1, 1]]
    i, d = select_refs(idx[:, :2], dist[:, :2], 4)          # fewer hits than slots
    assert i.tolist() == [[5, 9, -1, -1], [1, -1, -1, -1]] and d[0, 2:].tolist() == [1.0, 1.0]
    g = torch.Generator().manual_seed(0)
    i, d = select_refs(idx, dist, 4, uncond_video_ratio=1.0, generator=g)   # everything dropped
    assert bool((i == -1).all()) and bool((d == 1.0).all())
    g = torch.Generator().manual_seed(0)
    big = torch.arange(4000).view(1000, 4)
    i, d = select_refs(big, torch.zeros(1000, 4), 4, uncond_video_ratio=0.25, generator=g)
    frac = float((i == -1).float().mean())
    assert 0.2 < frac < 0.3 and bool((d[i == -1] == 1.0).all()) and bool((i[i >= 0] == big[i >= 0]).all())


def test_feature_table_builder_round_trip(tmp_path):
    """Row f-5: encode_vision driven over (row ids, clips) batches -> row-aligned bf16 table; rows never
    written (failed decodes) and the CFG branch share the encode_vision(zeros) row (module.py:327-329)."""
    from motionrag_b200 import build_feature_table, load_feature_rows
    L, C, n = 5, 64, 37

    class Model:
        calls = 0

        def encode_vision(self, videos):                      # [b, k, T, C, H, W] -> [b, k, L, C]
            Model.calls += 1
            b, k = videos.shape[:2]
            m = videos.float().mean(dim=(2, 3, 4, 5))           # a deterministic function of the clip
            return (m[..., None, None] + torch.arange(L)[:, None] * 0.5 + torch.arange(C)[None] * 0.01 + 7.0).expand(b, k, L, C)

    g = torch.Generator().manual_seed(0)
    clips = torch.randn(n, 4, 3, 8, 8, generator=g)
    order = torch.randperm(n, generator=g)
    skipped = {3, 20}                                           # unreadable clips

    def batches():
        ids = [int(i) for i in order if int(i) not in skipped]
        for lo in range(0, len(ids), 6):
            sel = torch.tensor(ids[lo:lo + 6])
            yield sel, clips[sel]

    meta = build_feature_table(Model(), batches(), n, tmp_path / "feat", tokens=L, width=C, device="cpu")
    assert meta["missing_rows"] == 2 and meta["n_rows"] == n
    feats, uncond, meta2 = load_feature_rows(tmp_path / "feat")
    assert feats.dtype == torch.bfloat16 and tuple(feats.shape) == (n, L, C) and meta2 == meta
    want = Model().encode_vision(clips[:, None])[:, 0].bfloat16()
    want_un = Model().encode_vision(torch.zeros(1, 1, 4, 3, 8, 8))[0, 0].bfloat16()
    assert torch.equal(uncond, want_un)
    for r in range(n):
        assert torch.equal(feats[r], want_un if r in skipped else want[r]), r
    part, _, _ = load_feature_rows(tmp_path / "feat", rows=(10, 25))       # a shard
    assert torch.equal(part, feats[10:25])
    with pytest.raises(IndexError):
        from motionrag_b200 import FeatureTableWriter
        FeatureTableWriter(tmp_path / "bad", 4, L, C).write([9], torch.zeros(1, L, C))


def _reference_schema_table(n=40, dim=768, seed=3):
    """A pyarrow table with the schema tools/build_rag_database.py:35-45 writes."""
    import pyarrow as pa
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=-1, keepdims=True)
    t = pa.table({"text": [f"caption {j}" for j in range(n)],
                  "text_embedding": pa.FixedSizeListArray.from_arrays(pa.array(emb.reshape(-1), type=pa.float32()), dim),
                  "id": pa.array(np.arange(n), type=pa.int64()), "uid": [f"uid{j}" for j in range(n)],
                  "dataset": ["openvid"] * n, "video": [f"v{j // 3}.mp4" for j in range(n)],
                  "start_sec": np.arange(n) * 2.0, "end_sec": np.arange(n) * 2.0 + 2.0})
    return t, emb


def test_arrow_dumps_of_the_reference_schema_open_without_lancedb(tmp_path):
    """Parquet file, directory of Parquet fragments, Arrow IPC file and IPC stream holding the reference's schema
    (FixedSizeList<f32>[768] + scalar columns) all read into the same columns."""
    import pyarrow as pa
    import pyarrow.feather as pf
    import pyarrow.parquet as pq
    from motionrag_b200 import tables
    t, emb = _reference_schema_table()
    pq.write_table(t, tmp_path / "a.parquet")
    pf.write_feather(t, tmp_path / "b.arrow")
    (tmp_path / "c").mkdir()
    pq.write_table(t.slice(0, 17), tmp_path / "c" / "part-0.parquet")
    pq.write_table(t.slice(17), tmp_path / "c" / "part-1.parquet")
    with pa.OSFile(str(tmp_path / "d.arrows"), "wb") as f, pa.ipc.new_stream(f, t.schema) as w:
        w.write_table(t, max_chunksize=16)
    for name in "abcd":
        cols = rag.RAGDatabase._load(tmp_path, name)
        assert cols["text_embedding"].dtype == np.float32 and np.array_equal(cols["text_embedding"], emb), name
        assert cols["video"].tolist() == t["video"].to_pylist() and cols["start_sec"].dtype == np.float64
        assert set(cols) == set(t.column_names)
        facts = tables.check_table(cols)
        assert facts["rows"] == 40 and facts["id_is_row_number"] and facts["text_embedding"]["unit_norm"]
        tables.feature_row_alignment(cols, 40)
    with pytest.raises(FileNotFoundError, match="no table"):
        tables.read_table(tmp_path, "missing")
    (tmp_path / "real.lance").mkdir()
    with pytest.raises(FileNotFoundError, match="export_lancedb"):
        tables.read_table(tmp_path, "real")


def test_table_checks_flag_misaligned_ids_nulls_and_non_unit_rows():
    import pyarrow as pa
    from motionrag_b200 import tables
    t, emb = _reference_schema_table(12, 16)
    cols = tables.arrow_to_columns(t)
    cols["id"] = cols["id"][::-1].copy()
    assert tables.check_table(cols)["id_is_row_number"] is False
    with pytest.raises(ValueError, match="not row-aligned"):
        tables.feature_row_alignment(cols, 12)
    with pytest.raises(ValueError, match="feature table has 11 rows"):
        tables.feature_row_alignment(tables.arrow_to_columns(t), 11)
    # NULL vectors become zero rows; a sliced array keeps its offset; List<float> columns are accepted
    mask = np.zeros(12, dtype=bool)
    mask[[3, 7]] = True
    vec = pa.FixedSizeListArray.from_arrays(pa.array(emb.reshape(-1)), 16, mask=pa.array(mask))
    want = emb.copy()
    want[mask] = 0
    assert np.array_equal(tables.vector_column_to_numpy(vec), want)
    assert np.array_equal(tables.vector_column_to_numpy(vec.slice(2, 8)), want[2:10])
    lst = pa.array([r.tolist() for r in emb], type=pa.list_(pa.float32()))
    assert np.array_equal(tables.vector_column_to_numpy(lst), emb)
    facts = tables.check_table({"text_embedding": want * 3})
    assert facts["text_embedding"]["zero_rows"] == 2 and not facts["text_embedding"]["unit_norm"]
    with pytest.raises(ValueError, match="different lengths"):
        tables.vector_column_to_numpy(pa.array([[1.0, 2.0], [1.0]], type=pa.list_(pa.float32())))


def test_loss_path_host_logic_equals_the_reference_recording(monkeypatch, golden_dir):
    """attach().batch_forward(return_loss=True) with the K4 gather replaced by its CPU restatement: everything
    around the kernel (target-only encoding, differentiable SOS / position / condition adds, the model's own
    transformer and get_loss) reproduces the losses and the sos_token gradient recorded from the reference's REAL
    ActionTransformer (tests/golden/cama_loss_f32.npz). The GPU twin runs the real kernel (tests/test_gpu_context.py)."""
    import types

    import torch.nn as nn
    import torch.nn.functional as F
    from motionrag_b200 import context as ctxmod
    gold = np.load(golden_dir / "cama_loss_f32.npz")
    C, heads, ff, layers = (int(v) for v in gold["shape"])
    ref, tgt = torch.from_numpy(gold["ref_feats"]), torch.from_numpy(gold["target"])
    cond, sos = torch.from_numpy(gold["cond"]), torch.from_numpy(gold["sos"])
    b, K, L, _ = ref.shape

    class HostTable:
        def __init__(self, t):
            self.local, self.L, self.Cdim, self.device, self.n_rows = t, t.shape[1], t.shape[2], t.device, t.shape[0]

    def host_gather(table, ref_index, sos_, un, pos, cond_, out=None, validate=False):
        feats = cc.gather_restatement(table.local, ref_index, un)
        return cc.context_restatement(feats, sos_.reshape(1, table.L, table.Cdim), None if pos is None else pos[None], cond_)
    monkeypatch.setattr(ctxmod, "gather_context", host_gather)
    monkeypatch.setattr(ctxmod, "FeatureTable", HostTable)

    class Cama(nn.Module):
        def __init__(self):
            super().__init__()
            layer = nn.TransformerEncoderLayer(C, heads, ff, 0.0, "gelu", batch_first=True, norm_first=False)
            self.transformer = nn.TransformerEncoder(layer, layers, enable_nested_tensor=False)
            self.transformer.load_state_dict({k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w:")})
            self.sos_token = nn.Parameter(sos.clone())
            self.register_buffer("pos_table", cc.sinusoid_table(256, C))

        def vision_pe(self, x):
            return x + self.pos_table[:, :x.size(-2)].type_as(x)

        def encode_condition(self, images):
            return images

        def encode_vision(self, videos):
            assert videos.shape[1] == 1          # only the target clip is ever encoded
            return videos[:, :, 0]

        def get_loss(self, pred, emb):
            pred, emb = pred.flatten(0, 1), emb.flatten(0, 1)
            mse = F.mse_loss(pred, emb)
            return types.SimpleNamespace(main=mse, mse=mse, smooth=F.smooth_l1_loss(pred, emb))

        def batch_forward(self, *a, **k):
            raise AssertionError("the video path must not run")

    rows = torch.randperm(64, generator=torch.Generator().manual_seed(1))[:b * K].view(b, K)
    table = torch.zeros(64, L, C)
    table[rows.flatten()] = ref.reshape(b * K, L, C)
    un = torch.randn(L, C, generator=torch.Generator().manual_seed(2))
    for tag, ignore in (("train", False), ("val", True)):
        model = Cama()
        ctxmod.attach(model, ctxmod.MotionContext(HostTable(table), sos, un, pe_max_length=256))
        for batch in ({"ref_index": rows, "ref_images": cond, "target_features": tgt},
                      {"ref_index": rows, "ref_images": cond, "video": tgt[:, None]},
                      {"ref_features": ref, "ref_images": cond, "target_features": tgt}):
            model.zero_grad()
            loss = model.batch_forward(batch, return_loss=True, ignore_ref_loss=ignore)
            loss.main.backward()
            assert float(loss.mse.detach()) == pytest.approx(float(gold[f"{tag}_mse"]), rel=1e-6)
            assert float(loss.smooth.detach()) == pytest.approx(float(gold[f"{tag}_smooth"]), rel=1e-6)
            torch.testing.assert_close(model.sos_token.grad, torch.from_numpy(gold[f"{tag}_sos_grad"]), rtol=1e-5, atol=1e-8)


def test_erfc_polynomial_gelu_of_the_gemm_epilogue_is_within_4e7_of_the_exact_gelu():
    """k5_cama.cu::gelu_erf evaluates the reference's exact (erf) GELU (nn.TransformerEncoderLayer(activation='gelu'),
    configs/cogvideox/MotionRAG_open.yml:253-267) through Abramowitz & Stegun 7.1.26; restated here in float32 step by
    step and held to the float64 GELU: the bound the kernel comment quotes."""
    from scipy.special import erf
    f = np.float32
    x = np.linspace(-8, 8, 400001).astype(f)
    z = np.abs(x) * f(0.70710678118654752)
    t = (f(1) / (f(0.3275911) * z + f(1))).astype(f)
    p = f(1.061405429) * t + f(-1.453152027)
    for c in (1.421413741, -0.284496736, 0.254829592):
        p = (p * t + f(c)).astype(f)
    e = np.exp2((z * z * f(-1.4426950408889634)).astype(f)).astype(f)
    r = (np.abs(f(0.5) * x) * (p * t * e).astype(f)).astype(f)
    got = np.where(x > 0, x - r, -r).astype(f)
    want = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    assert np.abs(got - want).max() < 4e-7
    assert np.abs(got - want).max() < 2.0 ** -9 * 1e-3      # three orders below a bf16 half-ulp at unit scale


@pytest.mark.parametrize("prefilter", [False, True])
def test_top_k_above_32_is_collected_in_passes_of_32(prefilter):
    """RAGDatabase._search_deep (host logic; the scans are played by the oracle's flat search standing in for the
    store): same rows, order and distances as ONE oracle search with the large k, for post- and pre-filter clauses."""
    from oracle import flat_search as fs
    db = _db_no_gpu(n=300, dim=16)
    emb = db._vectors["text_embedding"]
    emb[[40, 41, 42]] = emb[39]                                      # ties across the first pass border
    calls = []

    class Store:
        dim = 16
        groups = None

        def __len__(self):
            return 300

        def set_groups(self, g):
            self.groups = np.array(g, dtype=np.int64)

        def search_host(self, q, k, *, metric, path, refine, exclude_group, filter_mode, certify, reuse=False, list_len=0):
            assert 1 <= k <= 32 and filter_mode == "pre" and exclude_group.tolist() == [1] and q.shape[0] == 1
            calls.append(k)
            d, i = fs.flat_search(emb, q, k, metric, self.groups, np.array([1]), True)
            return d.astype(np.float32), i.astype(np.int64), None

    db._stores = {"text_embedding": Store()}
    db.metric, db.path, db.prefilter, db.recheck = "l2", "auto", prefilter, None
    ora = fs.OracleRAGDatabase({**db._columns, "text_embedding": emb}, prefilter=prefilter)
    rng = np.random.default_rng(3)
    for k, w in [(33, None), (70, "start_sec >= 100"), (100, 'video != "v13"'), (64, "id < 20 or id > 250")]:
        q = (emb[39] + 0.05 * rng.standard_normal(16)).astype(np.float32)
        calls.clear()
        dist, idx, single = db._search(q, None, k, w, 30)
        want = ora.text_search(q, top_k=k, where=w, select=["id"])
        n = int((idx[0] >= 0).sum())
        assert single and idx[0, :n].tolist() == [r["id"] for r in want] and (idx[0, n:] == -1).all()
        np.testing.assert_allclose(dist[0, :n], [r["_distance"] for r in want], rtol=1e-5, atol=1e-6)
        assert calls == [32] * (k // 32) + ([k % 32] if k % 32 else [])
        assert db._group_col is None                                 # the next ordinary call re-binds its labelling
    # more results than rows: passes stop when a scan comes back short
    dist, idx, _ = db._search(emb[:2], None, 320, None, 30)
    assert (idx >= 0).sum(-1).tolist() == [300, 300] and sorted(idx[1, :300].tolist()) == list(range(300))
    with pytest.raises(ValueError, match="exclude_group"):
        db._search(emb[0], None, 40, None, 30, exclude_group=np.zeros(1, dtype=np.int32))
