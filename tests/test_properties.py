"""Property tests (hypothesis) of the oracle and the host-side layouts — CPU only."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from motionrag_b200 import parallel
from motionrag_b200.context import select_refs
from oracle import flat_search as fs


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 300), dim=st.sampled_from([4, 16, 33]), nq=st.integers(1, 5), k=st.integers(1, 20),
       seed=st.integers(0, 10_000), metric=st.sampled_from(fs.METRICS), dup=st.booleans())
def test_flat_search_invariants(n, dim, nq, k, seed, metric, dup):
    rng = np.random.default_rng(seed)
    db = fs.normalise_rows(rng.standard_normal((n, dim)).astype(np.float32))
    if dup and n > 3:
        db[n // 2] = db[0]                                        # exact tie somewhere
    q = rng.standard_normal((nq, dim)).astype(np.float32) * rng.uniform(0.5, 12)
    d, i = fs.flat_search(db, q, k, metric)
    full = fs.distances(db, q, metric)
    for r in range(nq):
        valid = i[r] >= 0
        assert valid.sum() == min(k, n) and np.all(valid[:valid.sum()])          # prefix-valid, padded with -1
        assert len(set(i[r][valid].tolist())) == valid.sum()                     # no duplicates
        dv = d[r][valid]
        assert np.all(np.diff(dv) >= 0)                                          # ascending
        ties = np.diff(dv) == 0
        assert np.all(np.diff(i[r][valid])[ties] > 0)                            # ties -> lower index first
        kth = dv[-1]
        outside = np.setdiff1d(np.arange(n), i[r][valid])
        assert np.all(full[r][outside] >= kth)                                   # nothing better was left out
        np.testing.assert_array_equal(dv, full[r][i[r][valid]])
    d2, i2 = fs.flat_search(db, q, k + 3, metric)                                # top-k is a prefix of top-(k+3)
    np.testing.assert_array_equal(i2[:, :k][i[:, :k] >= 0], i[i >= 0])


@settings(max_examples=40, deadline=None)
@given(n=st.integers(3, 200), k=st.integers(1, 15), seed=st.integers(0, 10_000), gsize=st.integers(1, 5))
def test_filter_semantics(n, k, seed, gsize):
    rng = np.random.default_rng(seed)
    db = fs.normalise_rows(rng.standard_normal((n, 8)).astype(np.float32))
    q = db[rng.integers(0, n, 3)] * 4
    groups = (np.arange(n) // gsize).astype(np.int32)
    ex = rng.integers(-1, groups.max() + 1, 3).astype(np.int32)
    _, plain = fs.flat_search(db, q, k)
    _, post = fs.flat_search(db, q, k, "l2", groups, ex, prefilter=False)
    _, pre = fs.flat_search(db, q, k, "l2", groups, ex, prefilter=True)
    for r in range(3):
        keep = [j for j in plain[r] if j >= 0 and (ex[r] < 0 or groups[j] != ex[r])]
        assert post[r][post[r] >= 0].tolist() == keep                            # post = plain minus excluded
        pv = pre[r][pre[r] >= 0]
        assert np.all((ex[r] < 0) | (groups[pv] != ex[r]))
        assert pv.tolist()[:len(keep)] == keep                                   # pre extends post
        n_allowed = n if ex[r] < 0 else int((groups != ex[r]).sum())
        assert len(pv) == min(k, n_allowed)


@settings(max_examples=50, deadline=None)
@given(nq=st.integers(1, 40), k=st.integers(1, 32), world=st.integers(1, 8))
def test_packed_layout_is_disjoint_and_aligned(nq, k, world):
    lay = parallel.PackedLayout(nq, k)
    buf = torch.zeros(world * lay.nbytes, dtype=torch.uint8)
    d, g, i = lay.views(buf, world)
    d.fill_(1.0)
    assert int(g.abs().sum()) == 0 and int(i.abs().sum()) == 0                   # fields do not overlap
    i.fill_(-1)
    assert float(d.min()) == 1.0 and i[0].data_ptr() % 8 == 0 and lay.nbytes % 16 == 0


@settings(max_examples=50, deadline=None)
@given(n=st.integers(1, 10_000_000), world=st.integers(1, 8))
def test_shard_ranges_partition_the_table(n, world):
    spans = [parallel.shard_range(n, world, r) for r in range(world)]
    assert sum(hi - lo for _, lo, hi in spans) == n
    assert all(0 <= lo <= hi <= n for _, lo, hi in spans)
    assert all(a[2] == b[1] for a, b in zip(spans, spans[1:]))


@settings(max_examples=50, deadline=None)
@given(b=st.integers(1, 6), k=st.integers(1, 14), K=st.integers(1, 12), ratio=st.floats(0, 1), seed=st.integers(0, 999))
def test_select_refs_properties(b, k, K, ratio, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(-1, 50, (b, k), generator=g)
    dist = torch.rand(b, k, generator=g)
    i, d = select_refs(idx, dist, K, uncond_video_ratio=ratio, generator=g)
    assert i.shape == (b, K) and d.shape == (b, K)
    kept = i >= 0
    kk = min(k, K)
    assert bool((i[:, :kk][kept[:, :kk]] == idx[:, :kk][kept[:, :kk]]).all())   # kept slots are untouched
    assert bool((d[~kept] == 1.0).all())                                         # dropped / missing -> distance 1.0
    assert bool((i[:, kk:] == -1).all())
