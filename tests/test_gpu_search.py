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


_KNOB_WORKER = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from motionrag_b200 import EmbeddingStore, synthetic
from oracle import compare, flat_search as fs
n, dim = 20_011, 768
db = synthetic.database(n, dim, "clustered", seed=3, device="cpu").numpy()
src = np.random.default_rng(0).integers(0, n, 300)
q = synthetic.queries_from_rows(torch.from_numpy(db[src]), seed=9).numpy()
groups = (np.arange(n) // 3).astype(np.int32)
st = EmbeddingStore(dim, n, 0); st.append(db, normalise=False); st.set_groups(groups)
for nq, k, path in %s:
    ex = groups[src[:nq]].copy(); ex[::5] = -1
    for filt in (None, "post", "pre"):
        r = st.search(torch.from_numpy(q[:nq]).cuda(), k, path=path, exclude_group=None if filt is None else torch.from_numpy(ex).cuda(),
                      filter_mode=filt or "post", certify=True)
        rd, ri = fs.flat_search(db, q[:nq], k, "l2", groups if filt else None, ex if filt else None, prefilter=(filt == "pre"))
        rep = compare.check_retrieval(r.distance.cpu().numpy(), r.index.cpu().numpy(), rd, ri, db, q[:nq])
        assert rep["index_mismatches"] == rep["near_tie_positions"], rep
        assert not bool(torch.isnan(r.margin).any()) and (k > 12 or bool((r.margin > 0).all()))   # k == re-rank count: margin ~ 0
    print("PLAN", nq, st.plan(nq, k=k, path=path).grid, st.plan(nq, k=k, path=path).fused_tail)
print("KNOB_OK")
"""


def _run_with_env(env, cases):
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    r = subprocess.run([sys.executable, "-c", _KNOB_WORKER % (root, repr(cases))], capture_output=True, text=True,
                       timeout=600, env={**os.environ, **env})
    assert r.returncode == 0 and "KNOB_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    return r.stdout


def test_tensor_single_cta_kernel_on_multi_tile_batches(libmrag):
    """nq > 128 defaults to the CTA-pair kernel; MRAG_K2_SINGLE=1 (read once at library load, hence the
    subprocess) keeps the single-CTA kernel covered on multi-tile batches."""
    out = _run_with_env({"MRAG_K2_SINGLE": "1"}, [(129, 12, "tensor_bf16"), (200, 12, "tensor_bf16"), (256, 5, "tensor_bf16"),
                                                  (300, 32, "tensor_bf16")])
    assert "PLAN 300 148" in out


def test_single_query_scan_with_a_separate_k3_launch(libmrag):
    """The default single-query search is ONE launch (fused tail); MRAG_K1_FUSE=0 keeps the two-launch form
    (K1, then K3 under programmatic dependent launch) covered."""
    out = _run_with_env({"MRAG_K1_FUSE": "0"}, [(1, 12, "stream_bf16"), (1, 32, "stream_f32"), (1, 1, "auto")])
    assert "PLAN 1" in out and out.strip().splitlines()[0].endswith(" 0")


def test_single_query_scan_without_the_pipelined_launch(libmrag):
    """MRAG_K1_OVERLAP=0: fully stream-ordered launches on all 148 SMs (the A/B form of the default, which leaves
    one SM to the previous search's tail and starts its scan under it)."""
    out = _run_with_env({"MRAG_K1_OVERLAP": "0"}, [(1, 12, "stream_bf16"), (1, 32, "stream_f32")])
    assert "PLAN 1 148 1" in out
    assert "PLAN 1 147 1" in _run_with_env({}, [(1, 12, "auto")])


@pytest.mark.parametrize("path", ["auto", "stream_f32"])
def test_back_to_back_single_query_searches_pipeline_correctly(case, path):
    """Single-query searches are one launch each, started with programmatic stream serialization: the read-only
    scan of call i+1 runs under the select / re-rank tail of call i. 200 calls enqueued without any host sync in
    between (alternating filters, shared workspace and rotating tickets), every result checked afterwards, and
    the same sequence with a dependent torch op squeezed between the calls (which must see finished results)."""
    store, q, excl = case["store"], case["q"], case["excl"]
    qd, exd = torch.from_numpy(q).cuda(), torch.from_numpy(excl).cuda()
    pending = []
    for j in range(200):
        if j % 3 == 0:
            pending.append(store.search(qd[j:j + 1], 12, path=path, certify=(path == "auto")))
        else:
            pending.append(store.search(qd[j:j + 1], 12, path=path, exclude_group=exd[j:j + 1],
                                        filter_mode=("post", "pre")[j % 2], certify=(path == "auto")))
    torch.cuda.synchronize()
    for j, r in enumerate(pending):
        filt = None if j % 3 == 0 else ("post", "pre")[j % 2]
        rd, ri = fs.flat_search(case["db"], q[j:j + 1], 12, "l2", case["groups"] if filt else None,
                                excl[j:j + 1] if filt else None, prefilter=(filt == "pre"))
        rep = compare.check_retrieval(r.distance.cpu().numpy(), r.index.cpu().numpy(), rd, ri, case["db"], q[j:j + 1])
        assert rep["index_mismatches"] == rep["near_tie_positions"], (j, rep)
    # a consumer kernel right behind each search reads complete results (ordinary stream order)
    sums = []
    for j in range(50):
        r = store.search(qd[j:j + 1], 12, path=path)
        sums.append(r.index.sum() + 0)                 # torch kernel enqueued directly after the search
    torch.cuda.synchronize()
    for j, t in enumerate(sums):
        assert int(t) == int(fs.flat_search(case["db"], q[j:j + 1], 12)[1].sum()), j


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
    # the same guard covers BATCHES (the reference's real use, one query per annotation): of 40 queries the
    # 8 that sit on the twins are re-run in fp32, the rest are certified by the tensor-path margin
    others = rng.choice(np.setdiff1d(np.arange(n), twins), 32, replace=False)
    qb = np.concatenate([np.stack([base * s for s in (3, 5, 7, 9, 11, 13, 15, 17)]), db[others] * 6]).astype(np.float32)
    res = rdb.search_batch(qb, top_k=12, select=["video"])
    assert rdb.fp32_rechecks == 1 + 8 and rdb.deep_rechecks == 8      # deeper lists cannot separate 200 twins either
    rd, ri = fs.flat_search(db, qb, 12)
    got_i = np.array([[int(r["video"][1:]) for r in rr] for rr in res])
    got_d = np.array([[r["_distance"] for r in rr] for rr in res])
    rep = compare.check_retrieval(got_d, got_i, rd, ri, db, qb)
    assert rep["index_mismatches"] == rep["near_tie_positions"]
    assert [rr[0]["video"] for rr in res[8:]] == [f"v{j}" for j in others]
    for j in range(8):
        assert len(set(ri[j].tolist()) & set(got_i[j].tolist())) >= 10
    st.close()


def test_deeper_lists_certify_what_sixteen_candidates_cannot():
    """20 near-duplicates around the query: with 16-entry lists the 12th and the 16th best are both twins (margin ~ 0),
    with 32-entry lists the 32nd best is an ordinary row far away -> the re-issue with list_len = 32 certifies the
    query and the fp32 scan is never needed; the twins themselves are resolved by the exact fp32 re-rank."""
    from motionrag_b200 import EmbeddingStore, RAGDatabase
    from motionrag_b200.store import eps_typical
    rng = np.random.default_rng(5)
    n, dim = 5000, 768
    base = fs.normalise_rows(rng.standard_normal((1, dim)).astype(np.float32))[0]
    db = fs.normalise_rows(rng.standard_normal((n, dim)).astype(np.float32))
    twins = rng.choice(n, 20, replace=False)
    db[twins] = fs.normalise_rows(base[None] + 3e-2 / np.sqrt(dim) * rng.standard_normal((20, dim)).astype(np.float32))
    others = np.setdiff1d(np.arange(n), twins)[:7]
    qb = np.concatenate([(base * 6)[None], db[others] * 5]).astype(np.float32)
    st = EmbeddingStore(dim, n, 0)
    st.append(db, normalise=False)
    thr = eps_typical("tensor_bf16", dim)
    m16 = st.search(torch.from_numpy(qb).cuda(), 12, certify=True).margin.cpu().numpy()
    m32 = st.search(torch.from_numpy(qb).cuda(), 12, certify=True, list_len=32).margin.cpu().numpy()
    assert st.plan(8, k=12).rerank == 16 and st.plan(8, k=12, list_len=32).rerank == 32
    assert m16[0] <= thr < m32[0] and (m16[1:] > thr).all()
    rdb = RAGDatabase(None, None, columns={"text_embedding": db, "video": np.array([f"v{j}" for j in range(n)])})
    d, i = rdb.search_arrays(qb, top_k=12)
    assert rdb.deep_rechecks == 1 and rdb.fp32_rechecks == 0
    rd, ri = fs.flat_search(db, qb, 12)
    rep = compare.check_retrieval(d, i, rd, ri, db, qb)
    assert rep["index_mismatches"] == rep["near_tie_positions"]
    assert set(i[0].tolist()) <= set(twins.tolist())
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


def _bf16_normalised_table(rng, n, dim):
    """Rows normalised in bf16 like the reference's embedding model runs (tools/build_rag_database.py:17,
    torch_dtype bfloat16): |d|^2 is off by up to ~8e-3; plus zero-filled 'bad' vectors (on_bad_vectors='fill')."""
    raw = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)).bfloat16()
    rows = (raw / raw.float().norm(dim=-1, keepdim=True).bfloat16()).float().numpy()
    rows[rng.choice(n, 7, replace=False)] = 0
    return rows


@pytest.mark.parametrize("path,nq", [("stream_f32", 3), ("stream_bf16", 1), ("stream_bf16", 4), ("tensor_bf16", 40),
                                     ("tensor_bf16", 200)])
def test_l2_ranks_exactly_on_rows_that_are_not_unit_norm(path, nq):
    """Squared L2 orders like q.d - |d|^2/2: the per-row term is added to the scan score, so a table whose
    rows were normalised in bf16 (or contain zero-filled vectors) ranks exactly as LanceDB's L2 does —
    with the drop-in defaults (no re-normalisation, `_distance` of the rows as stored)."""
    from motionrag_b200 import EmbeddingStore
    rng = np.random.default_rng(11)
    n, dim = 6000, 768
    db = _bf16_normalised_table(rng, n, dim)
    # short queries make the norm term matter: |q| ~ 0.2, so 0.5 * 8e-3 is comparable to score gaps
    q = (db[rng.integers(0, n, nq)] * 0.2 + 0.02 * rng.standard_normal((nq, dim))).astype(np.float32)
    st = EmbeddingStore(dim, n, 0)
    st.append(db, normalise=False)
    info = st.info()
    assert 1e-4 < info.max_norm_deviation < 2e-2 and info.zero_rows == 7
    assert st.plan(nq, k=12, path=path).row_bias == 1
    res = st.search(torch.from_numpy(q).cuda(), 12, path=path, certify=True)
    rd, ri = fs.flat_search(db, q, 12, "l2")
    rep = compare.check_retrieval(res.distance.cpu().numpy(), res.index.cpu().numpy(), rd, ri, db, q, rtol=1e-4)
    assert rep["index_mismatches"] == rep["near_tie_positions"]
    # ranking by q.d alone would NOT have produced this list
    by_dot = np.argsort(-(q.astype(np.float64) @ db.T.astype(np.float64)), axis=-1, kind="stable")[:, :12]
    assert (by_dot != ri).any()
    assert bool((res.margin >= 0).all())
    st.close()


def test_cosine_needs_unit_rows_and_dot_takes_any():
    from motionrag_b200 import EmbeddingStore, MragError
    raw = np.random.default_rng(3).standard_normal((500, 256)).astype(np.float32) * 2
    raw[17] = 0                                            # a 'filled' bad vector
    st = EmbeddingStore(256, 500, 0)
    st.append(raw, normalise=False)
    info = st.info()
    assert info.max_norm_deviation > 1 and info.zero_rows == 1
    q = torch.randn(2, 256).cuda()
    with pytest.raises(MragError, match="cosine search needs unit-norm rows"):
        st.search(q, 5, metric="cosine")
    for path in ("stream_f32", "tensor_bf16"):
        res = st.search(q, 5, metric="l2", path=path)         # l2 on arbitrary rows: exact via the norm term
        rd, ri = fs.flat_search(raw, q.cpu().numpy(), 5, "l2")
        compare.check_retrieval(res.distance.cpu().numpy(), res.index.cpu().numpy(), rd, ri, raw, q.cpu().numpy())
    res = st.search(q, 5, metric="dot", path="stream_f32")          # dot ranks by q.d: fine
    want = torch.topk(q @ torch.from_numpy(raw).cuda().T, 5).indices
    assert torch.equal(res.index, want)
    st2 = EmbeddingStore(256, 500, 0)
    st2.append(raw, normalise=True)
    assert st2.info().max_norm_deviation < 1e-5 and st2.info().zero_rows == 1
    assert st2.search(q, 5).index.shape == (2, 5)
    st.close()
    st2.close()


@pytest.mark.parametrize("path,nq", [("stream_f32", 2), ("stream_bf16", 1), ("tensor_bf16", 130)])
def test_prefilter_is_exact_for_groups_larger_than_any_candidate_list(case, path, nq):
    """`video != own` as a PRE-filter is applied inside the scan: even when the excluded group swallows far
    more than the 32 nearest rows, the k nearest ELIGIBLE rows come back (and the margin is defined)."""
    big_groups = (np.arange(case["n"]) // 2000).astype(np.int32)          # ~10 groups of 2000 rows
    case["store"].set_groups(big_groups)
    try:
        q = case["q"][:nq]
        ex = big_groups[np.random.default_rng(0).integers(0, case["n"], 200)[:nq]]   # the query's own (source-row) group
        res = case["store"].search(torch.from_numpy(q).cuda(), 12, path=path, exclude_group=torch.from_numpy(ex).cuda(),
                                   filter_mode="pre", certify=True)
        rd, ri = fs.flat_search(case["db"], q, 12, "l2", big_groups, ex, prefilter=True)
        rep = compare.check_retrieval(res.distance.cpu().numpy(), res.index.cpu().numpy(), rd, ri, case["db"], q)
        assert rep["index_mismatches"] == rep["near_tie_positions"] and rep["positions"] == 12 * nq
        assert not np.isnan(res.margin.cpu().numpy()).any()
        assert np.all(big_groups[res.index.cpu().numpy()] != ex[:, None])
    finally:
        case["store"].set_groups(case["groups"])


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


# ---- BASELINE.json sizes: exact comparison against a float64 brute force ---------------------------
def _brute_force_fp64(rows_f32: torch.Tensor, q: torch.Tensor, k: int, chunk: int = 1 << 18):
    """Exact top-k of squared L2 on the GPU in float64, chunked over rows: -> (dist f64 [nq,k], idx i64 [nq,k]),
    ascending distance, ties by lowest row (test infrastructure; independent of libmrag)."""
    qd = q.double()
    qq = (qd * qd).sum(-1, keepdim=True)
    best_d = torch.full((q.shape[0], 0), 0.0, dtype=torch.float64, device=q.device)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64, device=q.device)
    for s0 in range(0, rows_f32.shape[0], chunk):
        r = rows_f32[s0:s0 + chunk].double()
        d = qq + (r * r).sum(-1)[None] - 2.0 * (qd @ r.T)
        kk = min(k, d.shape[1])
        cd, ci = torch.topk(d, kk, dim=-1, largest=False, sorted=True)
        best_d = torch.cat([best_d, cd], 1)
        best_i = torch.cat([best_i, ci + s0], 1)
        # keep the k best so far; ties -> lowest row: sort by (distance, row) via a stable two-pass sort
        o = torch.argsort(best_i, dim=-1, stable=True)
        best_d, best_i = best_d.gather(1, o), best_i.gather(1, o)
        o = torch.argsort(best_d, dim=-1, stable=True)[:, :k]
        best_d, best_i = best_d.gather(1, o), best_i.gather(1, o)
    return best_d, best_i


def _check_against_brute_force(store, q, res, k, rtol=1e-3):
    """oracle.compare.check_retrieval (the parity rule of BASELINE.md §5) against the float64 brute force.
    The table is too large for the host, so only the rows either side returned are brought over, under an
    order-preserving renumbering (index equality and tie order are unaffected)."""
    rd, ri = _brute_force_fp64(store.rows_f32(), q, k)
    gi = res.index
    assert bool((gi >= 0).all())
    rows = torch.unique(torch.cat([gi.flatten(), ri.flatten()]))            # sorted
    small = store.rows_f32()[rows].cpu().numpy()
    remap = lambda t: torch.searchsorted(rows, t.contiguous()).cpu().numpy()
    return compare.check_retrieval(res.distance.cpu().numpy(), remap(gi), rd.cpu().numpy(), remap(ri), small,
                                   q.cpu().numpy(), "l2", rtol=rtol)


@pytest.fixture(scope="module")
def big(libmrag):
    from motionrag_b200 import EmbeddingStore, synthetic
    n = 1_000_000
    st = EmbeddingStore(768, n, 0)
    synthetic.fill_store(st, n, "iid", seed=0)     # iid: worst case for near-ties
    st.set_groups(synthetic.groups(n, device="cuda"))
    torch.cuda.synchronize()
    return st


def test_c0_100k_table_64_queries_match_the_oracle(libmrag):
    """BASELINE config 0 (the reference's CPU-runnable case): 100 k entries, 64 queries, top-12 — every scan path
    against oracle.flat_search on the host, post-filter on, index mismatches only at documented near-ties."""
    from motionrag_b200 import EmbeddingStore, synthetic
    n, nq, k = 100_000, 64, 12
    db = synthetic.database(n, 768, "clustered", seed=4, device="cpu").numpy()
    src = np.random.default_rng(2).integers(0, n, nq)
    q = synthetic.queries_from_rows(torch.from_numpy(db[src]), seed=3).numpy()
    groups = (np.arange(n) // 3).astype(np.int32)
    ex = groups[src].astype(np.int32)
    st = EmbeddingStore(768, n, 0)
    st.append(db, normalise=False)
    st.set_groups(groups)
    rd, ri = fs.flat_search(db, q, k, "l2", groups, ex)
    for path in ("auto", "tensor_bf16", "stream_bf16", "stream_f32"):
        res = st.search(torch.from_numpy(q).cuda(), k, path=path, exclude_group=torch.from_numpy(ex).cuda(),
                        filter_mode="post", certify=True)
        rep = compare.check_retrieval(res.distance.cpu().numpy(), res.index.cpu().numpy(), rd, ri, db, q)
        assert rep["index_mismatches"] == rep["near_tie_positions"], (path, rep)
        assert rep["positions"] >= 9 * nq
    # the reference's call pattern: one query per call through the host entry point (single-launch scan)
    for j in range(8):
        d, i, _ = st.search_host(q[j:j + 1], k, exclude_group=ex[j:j + 1])
        compare.check_retrieval(d, i, rd[j:j + 1], ri[j:j + 1], db, q[j:j + 1])
    st.close()


def test_1m_paths_match_fp64_brute_force(big):
    """BASELINE configs 1/2 size: 256 queries against 1 M iid rows (worst case for near-ties) — tensor path for
    the batch, both streaming paths for the first queries, all against an exact float64 scan."""
    from motionrag_b200 import synthetic
    src = torch.randint(0, len(big), (256,), generator=torch.Generator().manual_seed(5))
    q = synthetic.queries_from_rows(big.rows_f32()[src.cuda()], seed=4)
    ref = big.search(q, 12, path="tensor_bf16", certify=True)
    rep = _check_against_brute_force(big, q, ref, 12)
    assert rep["index_mismatches"] == rep["near_tie_positions"]
    assert torch.equal(ref.index[:, 0].cpu(), src)              # each query's own source row wins
    assert bool((ref.margin > 0).all())
    for s in range(0, 16, 4):
        for path in ("stream_f32", "stream_bf16"):
            r = big.search(q[s:s + 4].contiguous(), 12, path=path)
            rep = _check_against_brute_force(big, q[s:s + 4].contiguous(), r, 12)
            assert rep["index_mismatches"] == rep["near_tie_positions"]
    for j in range(4):                                           # single-query, single-launch form
        r = big.search(q[j:j + 1].contiguous(), 12, certify=True)
        rep = _check_against_brute_force(big, q[j:j + 1].contiguous(), r, 12)
        assert rep["index_mismatches"] == rep["near_tie_positions"]


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
    plain = big.search(q, 12, certify=True)
    assert torch.equal(plain.index[:, 0], src)
    # the whole 4096-query batch against the exact scan (512 queries at a time to bound memory)
    mism = near = 0
    for s0 in range(0, 4096, 512):
        sub = type(plain)(plain.distance[s0:s0 + 512], plain.index[s0:s0 + 512], plain.group[s0:s0 + 512])
        rep = _check_against_brute_force(big, q[s0:s0 + 512].contiguous(), sub, 12)
        mism, near = mism + rep["index_mismatches"], near + rep["near_tie_positions"]
    assert mism == near
    # the filtered list is the unfiltered list minus the excluded rows, order preserved
    for qi in range(0, 4096, 512):
        keep = [i for i, gq in zip(plain.index[qi].tolist(), plain.group[qi].tolist()) if gq != int(ex[qi])]
        assert res.index[qi, :len(keep)].tolist() == keep
    # every query of the batch clears the statistical exactness test (nothing would be re-run in fp32)
    from motionrag_b200.store import eps_typical
    assert bool((plain.margin > eps_typical("tensor_bf16", 768)).all())


def test_10m_table_matches_fp64_brute_force_single_gpu(libmrag):
    """BASELINE config 3 size on one GPU (30.7 GB fp32 + 15.4 GB bf16): 64 queries against the exact
    float64 scan of all 10 M rows, tensor path and both streaming paths."""
    from motionrag_b200 import EmbeddingStore, synthetic
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~50 GB of free HBM")
    n = 10_000_000
    st = EmbeddingStore(768, n, 0)
    synthetic.fill_store(st, n, "clustered", seed=2)
    assert len(st) == n and st.info().max_norm_deviation < 1e-5
    src = torch.randint(0, n, (64,), generator=torch.Generator().manual_seed(1)).cuda()
    q = synthetic.queries_from_rows(st.rows_f32()[src], seed=5)
    big = st.search(q, 12, certify=True)                              # tensor path (one query tile)
    rep = _check_against_brute_force(st, q, big, 12)
    assert rep["index_mismatches"] == rep["near_tie_positions"]
    assert torch.equal(big.index[:, 0], src)
    q300 = synthetic.queries_from_rows(st.rows_f32()[torch.randint(0, n, (300,), generator=torch.Generator().manual_seed(2)).cuda()], seed=6)
    pair = st.search(q300, 12)                                        # CTA-pair kernel
    rep = _check_against_brute_force(st, q300, pair, 12)
    assert rep["index_mismatches"] == rep["near_tie_positions"]
    for path in ("stream_bf16", "stream_f32"):
        one = st.search(q[:3].contiguous(), 12, path=path)
        rep = _check_against_brute_force(st, q[:3].contiguous(), one, 12)
        assert rep["index_mismatches"] == rep["near_tie_positions"]
    one = st.search(q[:1].contiguous(), 12, certify=True)            # single launch
    rep = _check_against_brute_force(st, q[:1].contiguous(), one, 12)
    assert rep["index_mismatches"] == rep["near_tie_positions"] and float(one.margin[0]) > 0
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


@pytest.mark.parametrize("seed", range(20))
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
    # 50: one group can swallow a whole top-k (post-filter: fewer rows come back; pre-filter: the next
    # nearest eligible rows do)
    group_size = int(rng.choice([1, 3, 50]))
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
