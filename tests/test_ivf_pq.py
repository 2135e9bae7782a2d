"""Tables of more than 1 M rows: the reference searches an IVF-PQ `dot` index (tools/build_rag_database.py:51-52,
nprobes=50 / refine_factor=30 from src/data/rag.py:37) and is approximate; the product stays exact. These tests pin
the comparison that is defined in that regime (oracle/compare.py::check_recall) at a scaled-down size."""
import numpy as np
import pytest

from oracle import compare, flat_search as fs, ivf_pq
from oracle import lancedb_golden as lg


@pytest.fixture(scope="module")
def indexed():
    table = lg.case_table(seed=7, n=12000, kind="plain")
    rows = table["text_embedding"]
    # same shape as LanceDB's defaults (96 sub-vectors of 8 dims, 8-bit codes), partitions scaled with the table
    index = ivf_pq.IvfPqIndex(rows, num_partitions=48, num_sub_vectors=96, num_bits=8, sample=4096, iters=4, seed=1)
    q, src = lg.case_queries(7, table, 24)
    return table, index, q, src


def test_ivf_pq_restatement_behaves_like_an_ann_index(indexed):
    table, index, q, _ = indexed
    rows = table["text_embedding"]
    ed, ei = fs.flat_search(rows, q, 12, "dot")
    # probing every partition and re-scoring everything IS the exact search
    full = [index.search(v, 12, nprobes=48, refine_factor=10 ** 6) for v in q]
    assert all(np.array_equal(i, ei[j]) for j, (_, i) in enumerate(full))
    # the reference's knobs, scaled (nprobes 50/256 of the partitions): approximate but close
    ad, ai = map(np.stack, zip(*[index.search(v, 12, nprobes=10, refine_factor=30) for v in q]))
    rep = compare.check_recall(ed, ei, ad, ai)
    assert 0.7 <= rep["recall_at_k"] <= 1.0, rep
    # without the refine step distances are PQ estimates and recall drops: refine_factor matters
    nd, ni = map(np.stack, zip(*[index.search(v, 12, nprobes=10, refine_factor=None) for v in q]))
    recall_raw = np.mean([len(set(a) & set(b)) / 12 for a, b in zip(ni, ei)])
    assert recall_raw <= rep["recall_at_k"] + 1e-9


def test_check_recall_rejects_an_inexact_search(indexed):
    table, index, q, _ = indexed
    rows = table["text_embedding"]
    ed, ei = fs.flat_search(rows, q, 12, "dot")
    ad, ai = map(np.stack, zip(*[index.search(v, 12, nprobes=10, refine_factor=30) for v in q]))
    with pytest.raises(AssertionError):        # a "search" that lost its best row no longer dominates
        compare.check_recall(np.roll(ed, -1, 1), np.roll(ei, -1, 1), ad, ai)
    with pytest.raises(AssertionError):        # wrong distances for shared rows
        compare.check_recall(ed * 0.9, ei, ad, ai)


@pytest.mark.gpu
def test_exact_cuda_search_dominates_the_indexed_reference(libmrag, indexed):
    """RAGDatabase(metric='dot') — what `metric="reference"` resolves to above 1 M rows — against the IVF-PQ
    restatement of the reference: rank-by-rank domination, equal distances on shared rows, recall reported."""
    from motionrag_b200 import RAGDatabase
    table, index, q, _ = indexed
    db = RAGDatabase(None, None, 'cuda', columns=table, metric="dot")
    gd, gi = db.search_arrays(q, top_k=12)
    ad, ai = map(np.stack, zip(*[index.search(v, 12, nprobes=10, refine_factor=30) for v in q]))
    rep = compare.check_recall(gd, gi, ad, ai)
    assert rep["recall_at_k"] >= 0.7, rep
    ed, ei = fs.flat_search(table["text_embedding"], q, 12, "dot")
    compare.check_retrieval(gd, gi, ed, ei, table["text_embedding"], q, "dot")
    # metric="reference": squared L2 up to 1 M rows (LanceDB's default on an un-indexed table), dot above
    assert RAGDatabase(None, None, 'cuda', columns=table, metric="reference").metric == "l2"
