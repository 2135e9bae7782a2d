"""Parity of the CUDA search path (through the C ABI) against the oracle. Needs a B200."""
import numpy as np
import pytest
import torch

from oracle import compare, flat_search as fs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case(libmrag):
    """20 011 clustered rows (ragged against every tile size), 200 un-normalised queries."""
    from motionrag_b200 import EmbeddingStore, synthetic
    n, dim = 20_011, 768
    db = synthetic.database(n, dim, "clustered", seed=3, device="cpu").numpy()
    rng = np.random.default_rng(0)
    src = rng.integers(0, n, 200)
    q = synthetic.queries_from_rows(torch.from_numpy(db[src]), seed=9).numpy()
    groups = (np.arange(n) // 3).astype(np.int32)
    store = EmbeddingStore(dim, n, 0)
    store.append(db, normalise=False)
    store.set_groups(groups)
    torch.cuda.synchronize()
    # the kernels and the oracle must see the very same fp32 rows
    np.testing.assert_array_equal(store.rows_f32().cpu().numpy(), db)
    return dict(store=store, db=db, q=q, groups=groups, excl=groups[src].astype(np.int32), n=n)


def _run(case, nq, k, path, metric="l2", filt=None, refine=0, qoff=0):
    q = case["q"][qoff:qoff + nq]
    qd = torch.from_numpy(q).cuda()
    ex = ex_d = None
    mode = "post"
    if filt:
        ex = case["excl"][qoff:qoff + nq].copy()
        ex[::5] = -1
        ex_d = torch.from_numpy(ex).cuda()
        mode = filt
    res = case["store"].search(qd, k, metric=metric, path=path, refine=refine, exclude_group=ex_d, filter_mode=mode)
    torch.cuda.synchronize()
    rd, ri = fs.flat_search(case["db"], q, k, metric, case["groups"] if filt else None, ex, prefilter=(filt == "pre"))
    rep = compare.check_retrieval(res.distance.cpu().numpy(), res.index.cpu().numpy(), rd, ri, case["db"], q, metric)
    grp = res.group.cpu().numpy()
    idx = res.index.cpu().numpy()
    assert np.all(grp[idx >= 0] == case["groups"][idx[idx >= 0]]) and np.all(grp[idx < 0] == -1)
    return rep


@pytest.mark.parametrize("nq", [1, 2, 3, 4])
@pytest.mark.parametrize("k", [1, 12, 32])
def test_stream_f32_matches_oracle(case, nq, k):
    rep = _run(case, nq, k, "stream_f32", qoff=7 * nq)
    assert rep["index_mismatches"] == rep["near_tie_positions"]


@pytest.mark.parametrize("nq,path", [(5, "stream_f32"), (11, "stream_f32"), (9, "stream_bf16")])
def test_streaming_paths_loop_over_groups_of_four(case, nq, path):
    rep = _run(case, nq, 12, path, filt="post", qoff=20)
    assert rep["queries"] == nq


@pytest.mark.parametrize("nq,k", [(1, 12), (4, 12), (3, 32), (2, 1)])
def test_stream_bf16_rerank_matches_oracle(case, nq, k):
    _run(case, nq, k, "stream_bf16", qoff=11)


@pytest.mark.parametrize("nq,k", [(5, 12), (64, 12), (128, 12), (200, 12), (130, 32), (17, 1)])
def test_tensor_bf16_matches_oracle(case, nq, k):
    _run(case, nq, k, "tensor_bf16")


@pytest.mark.parametrize("nq,k", [(129, 12), (200, 12), (256, 5), (300, 32)])
def test_tensor_single_cta_kernel_on_multi_tile_batches(case, nq, k, monkeypatch):
    """nq > 128 defaults to the CTA-pair kernel; MRAG_K2_SINGLE=1 keeps the single-CTA one covered."""
    nq = min(nq, case["q"].shape[0])
    monkeypatch.setenv("MRAG_K2_SINGLE", "1")
    assert case["store"].plan(nq, k=k).grid % 2 == 1 or case["store"].plan(nq, k=k).grid <= 148
    _run(case, nq, k, "tensor_bf16")
    monkeypatch.delenv("MRAG_K2_SINGLE")
    _run(case, nq, k, "tensor_bf16")


@pytest.mark.parametrize("path,nq", [("stream_f32", 4), ("stream_bf16", 3), ("tensor_bf16", 64)])
@pytest.mark.parametrize("metric", ["cosine", "dot"])
def test_metrics(case, path, nq, metric):
    _run(case, nq, 12, path, metric=metric)


@pytest.mark.parametrize("path,nq", [("stream_f32", 4), ("stream_bf16", 4), ("tensor_bf16", 150)])
@pytest.mark.parametrize("filt", ["post", "pre"])
def test_video_exclusion_filter(case, path, nq, filt):
    rep = _run(case, nq, 12, path, filt=filt)
    assert rep["positions"] > 0


def test_auto_path_dispatch(case):
    st = case["store"]
    assert st.plan(1, k=12).path == 2 and st.plan(2, k=12).path == 3 and st.plan(5, k=12).path == 3
    p = st.plan(1, k=12, path="stream_f32")
    assert p.scan_bytes == case["n"] * 768 * 4 and p.cands_per_query == p.grid * 16
    assert st.plan(1, k=12).scan_bytes == case["n"] * 768 * 2 and st.plan(1, k=12).cands_per_query == st.plan(1, k=12).grid * 32
    p2 = st.plan(4096, k=12)
    assert p2.m_tiles == 32 and p2.n_tiles == (case["n"] + 255) // 256 and p2.scan_flops == 2 * 4096 * case["n"] * 768


def test_host_buffer_entry_point_equals_device_path(case):
    q = case["q"][:3]
    ex = case["excl"][:3]
    d, i, g = case["store"].search_host(q, 12, exclude_group=ex)
    res = case["store"].search(torch.from_numpy(q).cuda(), 12, exclude_group=torch.from_numpy(ex).cuda())
    np.testing.assert_array_equal(i, res.index.cpu().numpy())
    np.testing.assert_array_equal(d, res.distance.cpu().numpy())
    np.testing.assert_array_equal(g, res.group.cpu().numpy())


@pytest.mark.parametrize("nq,path", [(1, "auto"), (4, "stream_bf16"), (64, "auto"), (200, "auto")])
def test_host_buffer_path_replays_and_matches_oracle(case, nq, path):
    """Small host calls run as a captured CUDA graph; repeated calls with different data and a
    store mutation in between must stay correct."""
    for rep in range(3):
        q = case["q"][rep * 5: rep * 5 + nq]
        ex = case["excl"][rep * 5: rep * 5 + nq]
        d, i, g = case["store"].search_host(q, 12, path=path, exclude_group=ex)
        rd, ri = fs.flat_search(case["db"], q, 12, "l2", case["groups"], ex)
        compare.check_retrieval(d, i, rd, ri, case["db"], q)
    case["store"].set_groups(case["groups"])          # drops cached graphs
    d2, i2, _ = case["store"].search_host(q, 12, path=path, exclude_group=ex)
    np.testing.assert_array_equal(i2, i)


def _bf16_round(a):
    return torch.from_numpy(a).bfloat16().float().numpy()


@pytest.mark.parametrize("path,nq,rr", [("stream_bf16", 3, 32), ("tensor_bf16", 40, 16), ("tensor_bf16", 200, 16)])
def test_exactness_margin_matches_its_definition(case, path, nq, rr):
    """margin = (true q.d of the k-th hit - scan score of the weakest re-ranked row) / |q|."""
    from motionrag_b200.store import EPS, eps_typical
    q = case["q"][:nq]
    res = case["store"].search(torch.from_numpy(q).cuda(), 12, path=path, certify=True)
    m = res.margin.cpu().numpy()
    db16 = _bf16_round(case["db"])
    q16 = _bf16_round(q) if path == "tensor_bf16" else q
    scan = (q16.astype(np.float64) @ db16.T.astype(np.float64))
    true = (q.astype(np.float64) @ case["db"].T.astype(np.float64))
    idx = res.index.cpu().numpy()
    for r in range(nq):
        weakest = np.sort(scan[r])[::-1][rr - 1]
        want = (true[r, idx[r, 11]] - weakest) / np.linalg.norm(q[r])
        assert m[r] == pytest.approx(want, abs=2e-4), (r, m[r], want)
    assert np.all(m > eps_typical(path, 768))        # far above the rounding noise (statistical pass) ...
    assert np.mean(m > EPS[path]) > 0.3              # ... and often above the worst-case bound (proof)
    # the fp32 stream has nothing to certify against rounding: margins are >= 0 by construction
    r32 = case["store"].search(torch.from_numpy(q[:4]).cuda(), 12, path="stream_f32", certify=True)
    assert bool((r32.margin >= 0).all())


def test_uncertified_queries_fall_back_to_the_fp32_scan():
    """Rows that differ only below bf16 resolution: the bf16 scan cannot rank them, the margin
    says so, and RAGDatabase re-runs those queries on the fp32 master rows -> exact answer."""
    from motionrag_b200 import EmbeddingStore, RAGDatabase
    from motionrag_b200.store import EPS
    rng = np.random.default_rng(0)
    n, dim = 4000, 768
    base = fs.normalise_rows(rng.standard_normal((1, dim)).astype(np.float32))[0]
    db = fs.normalise_rows(rng.standard_normal((n, dim)).astype(np.float32))
    twins = rng.choice(n, 200, replace=False)
    # 200 rows within ~1e-4 (cosine) of each other: resolvable in fp32 (noise ~1e-6), not by a bf16 scan (~2e-5)
    db[twins] = fs.normalise_rows(base[None] + 1e-2 / np.sqrt(dim) * rng.standard_normal((200, dim)).astype(np.float32))
    q = (base * 9).astype(np.float32)[None]
    st = EmbeddingStore(dim, n, 0)
    st.append(db, normalise=False)
    res = st.search(torch.from_numpy(q).cuda(), 12, path="stream_bf16", certify=True)
    from motionrag_b200.store import eps_typical
    assert float(res.margin[0]) <= eps_typical("stream_bf16", dim) < EPS["stream_bf16"]   # flagged
    cols = {"text_embedding": db, "video": np.array([f"v{j}" for j in range(n)])}
    rdb = RAGDatabase(None, None, columns=cols)
    got = rdb.text_search(q[0], top_k=12, select=["video"])
    rd, ri = fs.flat_search(db, q, 12)
    assert rdb.fp32_rechecks == 1
    rep = compare.check_retrieval(np.array([[r["_distance"] for r in got]]),
                                  np.array([[int(r["video"][1:]) for r in got]]), rd, ri, db, q)
    assert rep["positions"] == 12
    exact_set = set(ri[0].tolist())
    assert len(exact_set & {int(r["video"][1:]) for r in got}) >= 10        # the fp32 scan resolves the twins
    # an ordinary query on the same table is certified and does not pay the second scan
    got2 = rdb.text_search(db[7] * 3, top_k=5, select=["video"])
    assert got2[0]["video"] == "v7" and rdb.fp32_rechecks == 1
    st.close()


def test_index_base_offsets_global_ids(case):
    qd = torch.from_numpy(case["q"][:2]).cuda()
    a = case["store"].search(qd, 5).index.cpu()
    b = case["store"].search(qd, 5, index_base=1_000_000_007).index.cpu()
    assert torch.equal(a + 1_000_000_007, b)


def test_argument_errors_are_loud(case):
    from motionrag_b200 import MragError
    qd = torch.from_numpy(case["q"][:5]).cuda()
    with pytest.raises(MragError, match="k must be"):
        case["store"].search(qd, 33)
    with pytest.raises(MragError, match="nq must be <= 65536"):
        case["store"].plan(70000, k=12)
    with pytest.raises(ValueError):
        case["store"].search(qd.double(), 12)
    with pytest.raises(MragError, match="capacity"):
        case["store"].append(case["db"][:4096], normalise=False)


@pytest.mark.parametrize("path", ["stream_f32", "stream_bf16", "tensor_bf16"])
def test_golden_fixture_with_duplicates_and_zero_distance(golden_dir, path):
    from motionrag_b200 import EmbeddingStore
    z = np.load(golden_dir / "retrieval_small.npz")
    db, q, k = z["db"], z["queries"], int(z["k"])
    st = EmbeddingStore(db.shape[1], db.shape[0], 0)
    st.append(db, normalise=False)
    st.set_groups(z["row_group"])
    nq = 4 if path != "tensor_bf16" else q.shape[0]
    res = st.search(torch.from_numpy(q[:nq]).cuda(), k, path=path)
    rep = compare.check_retrieval(res.distance.cpu().numpy(), res.index.cpu().numpy(), z["l2_dist"][:nq],
                                  z["l2_idx"][:nq], db, q[:nq])
    # query 0 equals rows 3, 7 and 1500 exactly: ties resolve to the lowest index, distance 0
    assert res.index[0, :3].tolist() == [3, 7, 1500] and float(res.distance[0, 0]) == 0.0
    ex = torch.from_numpy(z["exclude_group"][:nq]).cuda()
    for mode in ("post", "pre"):
        r = st.search(torch.from_numpy(q[:nq]).cuda(), k, path=path, exclude_group=ex, filter_mode=mode)
        compare.check_retrieval(r.distance.cpu().numpy(), r.index.cpu().numpy(), z[f"l2_{mode}_dist"][:nq],
                                z[f"l2_{mode}_idx"][:nq], db, q[:nq])
    st.close()
    assert rep["queries"] == nq


@pytest.mark.parametrize("n", [1, 5, 31, 257])
@pytest.mark.parametrize("path,nq", [("stream_f32", 2), ("stream_bf16", 1), ("tensor_bf16", 9)])
def test_tiny_tables_pad_with_minus_one(n, path, nq):
    from motionrag_b200 import EmbeddingStore
    rng = np.random.default_rng(n)
    db = fs.normalise_rows(rng.standard_normal((n, 256)).astype(np.float32))
    q = rng.standard_normal((nq, 256)).astype(np.float32) * 4
    st = EmbeddingStore(256, n, 0)
    st.append(db, normalise=False)
    res = st.search(torch.from_numpy(q).cuda(), 12, path=path)
    rd, ri = fs.flat_search(db, q, 12)
    compare.check_retrieval(res.distance.cpu().numpy(), res.index.cpu().numpy(), rd, ri, db, q)
    assert (res.index >= 0).sum(-1).tolist() == [min(n, 12)] * nq
    st.close()


@pytest.mark.parametrize("dim", [256, 512, 1024])
@pytest.mark.parametrize("path,nq", [("stream_f32", 3), ("stream_bf16", 2), ("tensor_bf16", 70)])
def test_other_embedding_widths(dim, path, nq):
    """dim is a store property (gte-base is 768; the image column of rag.py:82-99 may differ)."""
    from motionrag_b200 import EmbeddingStore
    rng = np.random.default_rng(dim)
    n = 3001
    db = fs.normalise_rows(rng.standard_normal((n, dim)).astype(np.float32))
    q = (db[rng.integers(0, n, nq)] + 0.02 * rng.standard_normal((nq, dim)).astype(np.float32)) * 7
    st = EmbeddingStore(dim, n, 0)
    st.append(db, normalise=False)
    res = st.search(torch.from_numpy(q.astype(np.float32)).cuda(), 12, path=path)
    rd, ri = fs.flat_search(db, q, 12)
    compare.check_retrieval(res.distance.cpu().numpy(), res.index.cpu().numpy(), rd, ri, db, q)
    st.close()


def test_unsupported_width_is_loud():
    from motionrag_b200 import EmbeddingStore, MragError
    with pytest.raises(MragError, match="dim must be"):
        EmbeddingStore(100, 10, 0)
    st = EmbeddingStore(320, 10, 0)                      # tensor path handles any multiple of 64 ...
    st.append(np.random.default_rng(0).standard_normal((10, 320)).astype(np.float32))
    q = torch.randn(9, 320).cuda()
    assert st.search(q, 3).index.shape == (9, 3)
    assert st.plan(1, k=3).path == 3                      # ... and AUTO routes around the streaming kernels,
    with pytest.raises(MragError, match="streaming path"):  # which are instantiated per width
        st.search(q[:1].contiguous(), 3, path="stream_f32")
    st.close()


def test_store_normalises_on_upload():
    from motionrag_b200 import EmbeddingStore
    raw = np.random.default_rng(1).standard_normal((1000, 512)).astype(np.float32) * 5
    st = EmbeddingStore(512, 1000, 0)
    st.append(torch.from_numpy(raw).cuda(), normalise=True)
    got = st.rows_f32().cpu().numpy()
    np.testing.assert_allclose(got, fs.normalise_rows(raw), rtol=2e-6, atol=1e-7)
    bf = st.rows_bf16().float().cpu().numpy()
    np.testing.assert_allclose(bf, got, rtol=2 ** -8, atol=1e-6)
    st.close()


def test_non_unit_rows_are_refused_for_l2_and_cosine():
    """The scan ranks by q.d; that equals the L2 / cosine order only for unit rows, so a table
    that is not normalised must fail loudly instead of returning a wrong ranking."""
    from motionrag_b200 import EmbeddingStore, MragError
    raw = np.random.default_rng(3).standard_normal((500, 256)).astype(np.float32) * 2
    raw[17] = 0                                            # a 'filled' bad vector
    st = EmbeddingStore(256, 500, 0)
    st.append(raw, normalise=False)
    info = st.info()
    assert info.max_norm_deviation > 1 and info.zero_rows == 1
    q = torch.randn(2, 256).cuda()
    for metric in ("l2", "cosine"):
        with pytest.raises(MragError, match="not unit-norm"):
            st.search(q, 5, metric=metric)
    res = st.search(q, 5, metric="dot", path="stream_f32")          # dot ranks by q.d: fine
    want = torch.topk(q @ torch.from_numpy(raw).cuda().T, 5).indices
    assert torch.equal(res.index, want)
    st2 = EmbeddingStore(256, 500, 0)
    st2.append(raw, normalise=True)
    assert st2.info().max_norm_deviation < 1e-5 and st2.info().zero_rows == 1
    assert st2.search(q, 5).index.shape == (2, 5)
    st.close()
    st2.close()


def test_shard_save_load_roundtrip(case, tmp_path):
    from motionrag_b200 import EmbeddingStore
    case["store"].save(tmp_path / "shard0")
    st = EmbeddingStore.load(tmp_path / "shard0", 0)
    assert len(st) == case["n"] and st.info().has_groups == 1
    assert torch.equal(st.rows_f32(), case["store"].rows_f32()) and torch.equal(st.rows_bf16(), case["store"].rows_bf16())
    qd = torch.from_numpy(case["q"][:4]).cuda()
    exd = torch.from_numpy(case["excl"][:4]).cuda()
    a = case["store"].search(qd, 12, exclude_group=exd)
    b = st.search(qd, 12, exclude_group=exd)
    assert torch.equal(a.index, b.index) and torch.equal(a.distance, b.distance)
    st.close()


def test_merge_topk_matches_numpy():
    from motionrag_b200 import merge_topk
    rng = np.random.default_rng(2)
    G, nq, k = 4, 37, 12
    dist = np.sort(rng.random((G, nq, k)).astype(np.float32), -1)
    dist[1, :, 0] = dist[0, :, 0]                       # cross-shard ties -> lower shard first
    idx = np.arange(G * nq * k, dtype=np.int64).reshape(nq, G, k).transpose(1, 0, 2).copy()
    idx = np.sort(idx, -1) + (np.arange(G) * 10_000_000)[:, None, None]
    idx[3, 5, 8:] = -1
    dist[3, 5, 8:] = np.inf
    grp = (idx % 7).astype(np.int32)
    ex = rng.integers(-1, 7, nq).astype(np.int32)
    for mode in ("none", "post", "pre"):
        out = merge_topk(torch.from_numpy(dist).cuda(), torch.from_numpy(idx).cuda(), torch.from_numpy(grp).cuda(),
                         k, torch.from_numpy(ex).cuda() if mode != "none" else None, mode)
        for q in range(nq):
            flat = [(dist[g, q, j], g * k + j, idx[g, q, j], grp[g, q, j]) for g in range(G) for j in range(k)
                    if idx[g, q, j] >= 0]
            flat.sort(key=lambda t: (t[0], t[1]))
            if mode == "post":
                flat = [t for t in flat[:k] if ex[q] < 0 or t[3] != ex[q]]
            elif mode == "pre":
                flat = [t for t in flat if ex[q] < 0 or t[3] != ex[q]]
            flat = flat[:k]
            assert out.index[q, :len(flat)].tolist() == [t[2] for t in flat], (mode, q)
            assert out.index[q, len(flat):].tolist() == [-1] * (k - len(flat))
            np.testing.assert_array_equal(out.distance[q, :len(flat)].cpu().numpy(), [t[0] for t in flat])


# ---- BASELINE.json sizes: size-independent properties at 1 M rows ---------------------------------
@pytest.fixture(scope="module")
def big(libmrag):
    from motionrag_b200 import EmbeddingStore, synthetic
    n = 1_000_000
    st = EmbeddingStore(768, n, 0)
    synthetic.fill_store(st, n, "iid", seed=0)     # iid: worst case for near-ties
    st.set_groups(synthetic.groups(n, device="cuda"))
    torch.cuda.synchronize()
    return st


def _exact_on_gpu(st, q, idx):
    rows = st.rows_f32()[idx.clamp_min(0)].double()
    return ((rows - q[:, None, :].double()) ** 2).sum(-1)


def test_1m_paths_agree_and_scores_are_exact(big):
    from motionrag_b200 import synthetic
    src = torch.randint(0, len(big), (256,), generator=torch.Generator().manual_seed(5))
    q = synthetic.queries_from_rows(big.rows_f32()[src.cuda()], seed=4)
    ref = big.search(q, 12, path="tensor_bf16")
    exact = _exact_on_gpu(big, q, ref.index)
    assert torch.allclose(ref.distance.double(), exact, rtol=1e-3, atol=1e-6)
    assert bool((ref.distance[:, 1:] >= ref.distance[:, :-1]).all())
    assert torch.equal(ref.index[:, 0].cpu(), src)              # each query's own source row wins
    # independent check of the whole top-12 with a torch fp32 GEMM + topk on the GPU
    sims = q @ big.rows_f32().T
    top = sims.topk(12, dim=-1).indices
    same = (top.sort(-1).values == ref.index.sort(-1).values).all(-1)
    assert same.float().mean() > 0.98                            # the rest are documented near-ties
    for s in range(0, 16, 4):
        for path in ("stream_f32", "stream_bf16"):
            r = big.search(q[s:s + 4].contiguous(), 12, path=path)
            agree = (r.index == ref.index[s:s + 4])
            if not bool(agree.all()):                            # only near-ties may differ
                e1 = _exact_on_gpu(big, q[s:s + 4], r.index)
                assert torch.allclose(e1, exact[s:s + 4], rtol=1e-3)
            assert torch.allclose(r.distance, ref.distance[s:s + 4], rtol=1e-3, atol=1e-6)


def test_1m_batch_4096_filter_properties(big):
    """BASELINE config 2 shape: 4096 queries; post-filter never returns the excluded video."""
    from motionrag_b200 import synthetic
    g = torch.Generator().manual_seed(6)
    src = torch.randint(0, len(big), (4096,), generator=g).cuda()
    q = synthetic.queries_from_rows(big.rows_f32()[src], seed=8)
    ex = (src // 3).to(torch.int32)
    res = big.search(q, 12, exclude_group=ex, filter_mode="post")
    assert bool(((res.group != ex[:, None]) | (res.index < 0)).all())
    n_valid = (res.index >= 0).sum(-1)
    assert int(n_valid.min()) >= 9 and int(n_valid.max()) <= 12   # own video has 3 clips
    plain = big.search(q, 12)
    assert torch.equal(plain.index[:, 0], src)
    # the filtered list is the unfiltered list minus the excluded rows, order preserved
    for qi in range(0, 4096, 512):
        keep = [i for i, gq in zip(plain.index[qi].tolist(), plain.group[qi].tolist()) if gq != int(ex[qi])]
        assert res.index[qi, :len(keep)].tolist() == keep


def test_10m_table_properties_single_gpu(libmrag):
    """BASELINE config 3 size on one GPU (30.7 GB fp32 + 15.4 GB bf16): self-retrieval, ordering,
    agreement of the three scan paths, exact distances — properties that need no 10 M-row oracle."""
    from motionrag_b200 import EmbeddingStore, synthetic
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~50 GB of free HBM")
    n = 10_000_000
    st = EmbeddingStore(768, n, 0)
    synthetic.fill_store(st, n, "clustered", seed=2)
    assert len(st) == n and st.info().max_norm_deviation < 1e-5
    src = torch.randint(0, n, (260,), generator=torch.Generator().manual_seed(1)).cuda()
    q = synthetic.queries_from_rows(st.rows_f32()[src], seed=5)
    big = st.search(q, 12)                                           # tensor path, CTA pairs
    assert torch.equal(big.index[:, 0], src)
    assert bool((big.distance[:, 1:] >= big.distance[:, :-1]).all()) and bool((big.index >= 0).all())
    rows = st.rows_f32()[big.index.flatten()].view(260, 12, 768).double()
    exact = ((rows - q[:, None].double()) ** 2).sum(-1)
    assert torch.allclose(big.distance.double(), exact, rtol=1e-3, atol=1e-6)
    for path in ("stream_bf16", "stream_f32"):
        one = st.search(q[:3].contiguous(), 12, path=path)
        same = one.index == big.index[:3]
        if not bool(same.all()):                                     # only near-ties may differ
            assert torch.allclose(one.distance, big.distance[:3], rtol=1e-3, atol=1e-6)
        assert torch.equal(one.index[:, 0], src[:3])
    st.close()


def test_searches_on_different_streams_do_not_share_scratch(case):
    """SURVEY §8b ownership: the store is immutable after upload, so searches issued on different
    streams may overlap; each stream gets its own candidate/threshold scratch."""
    store = case["store"]
    qa = torch.from_numpy(case["q"][:64]).cuda()
    qb = torch.from_numpy(case["q"][64:128]).cuda()
    want_a, want_b = store.search(qa, 12, path="tensor_bf16"), store.search(qb, 12, path="stream_bf16")
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    got = []
    for _ in range(10):                      # interleave so the kernels of both streams overlap
        with torch.cuda.stream(s1):
            ra = store.search(qa, 12, path="tensor_bf16")
        with torch.cuda.stream(s2):
            rb = store.search(qb, 12, path="stream_bf16")
        got.append((ra, rb))
    torch.cuda.synchronize()
    for ra, rb in got:
        assert torch.equal(ra.index, want_a.index) and torch.equal(ra.distance, want_a.distance)
        assert torch.equal(rb.index, want_b.index) and torch.equal(rb.distance, want_b.distance)


@pytest.mark.parametrize("seed", range(14))
def test_random_configurations_match_oracle(seed):
    """Seeded random draws over table size, width, batch size, k, scan path, metric and filter mode
    (ragged against every tile size; small tables, groups that swallow whole result lists)."""
    from motionrag_b200 import EmbeddingStore
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([1, 7, 40, 255, 256, 257, 1000, 2999]))
    dim = int(rng.choice([256, 768]))
    path = str(rng.choice(["stream_f32", "stream_bf16", "tensor_bf16"]))
    nq = int(rng.integers(1, 5)) if path.startswith("stream") and rng.random() < 0.7 else int(rng.choice([1, 5, 9, 127, 128, 129, 140]))
    k = int(rng.choice([1, 3, 12, 20, 32]))
    metric = str(rng.choice(["l2", "l2", "cosine", "dot"]))
    filt = str(rng.choice(["none", "post", "pre"]))
    cent = fs.normalise_rows(rng.standard_normal((8, dim)).astype(np.float32))
    db = fs.normalise_rows(cent[rng.integers(0, 8, n)] + 0.5 / np.sqrt(dim) * rng.standard_normal((n, dim)).astype(np.float32))
    src = rng.integers(0, n, nq)
    q = ((db[src] + 0.1 / np.sqrt(dim) * rng.standard_normal((nq, dim))) * rng.uniform(0.5, 20, (nq, 1))).astype(np.float32)
    # 50: one group can swallow a whole post-filtered top-k. A PRE-filter is exact while the excluded rows
    # among the re-ranked candidates leave k of them (documented limit), i.e. for video-sized groups.
    group_size = int(rng.choice([1, 3, 50])) if filt != "pre" else int(rng.choice([1, 3]))
    groups = (np.arange(n) // group_size).astype(np.int32)
    excl = groups[src].astype(np.int32)
    excl[::3] = -1
    st = EmbeddingStore(dim, n, 0)
    st.append(db, normalise=False)
    st.set_groups(groups)
    ex_d = torch.from_numpy(excl).cuda() if filt != "none" else None
    res = st.search(torch.from_numpy(q).cuda(), k, metric=metric, path=path, exclude_group=ex_d,
                    filter_mode=filt if filt != "none" else "post")
    rd, ri = fs.flat_search(db, q, k, metric, groups if filt != "none" else None, excl if filt != "none" else None,
                            prefilter=(filt == "pre"))
    got_i, got_d = res.index.cpu().numpy(), res.distance.cpu().numpy()
    cfg = dict(n=n, dim=dim, path=path, nq=nq, k=k, metric=metric, filt=filt, group_size=group_size)
    assert ((got_i >= 0).sum(-1) == (ri >= 0).sum(-1)).all(), cfg       # same number of results per query
    rep = compare.check_retrieval(got_d, got_i, rd, ri, db, q, metric)
    assert rep["index_mismatches"] == rep["near_tie_positions"], (cfg, rep)
    st.close()
