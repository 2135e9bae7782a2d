"""ORACLE (test infrastructure) — writes the committed fixtures under tests/golden/.

Run in the build container (the only place /root/reference exists):
    python -m oracle.make_golden

cama_context_{bf16,f32}.npz   inputs + the (x, mask) captured from the reference's REAL
                              ActionTransformer.batch_forward (oracle/cama_context.py) —
                              these pin the gather/context kernel to the reference itself.
rag_reference_class.json      what the reference's REAL RAGDatabase class (src/data/rag.py) returns for
                              the calls prepare_annotations makes (and the other public methods), run
                              over the LanceDB stand-in oracle/fake_lancedb.py.
cama_loss_f32.npz             mse / smooth-L1 losses and the sos_token gradient of the reference's REAL
                              ActionTransformer.batch_forward(return_loss=True) (training_step and
                              validation_step, module.py:305-311, 333-351) with a seeded toy transformer.
retrieval_small.npz           seeded database / queries / group ids with the oracle's own
                              answers for l2 / cosine / dot, post- and pre-filter. LanceDB is
                              not installable here, so these pin the oracle against
                              regressions only (parity unpinned, see oracle/flat_search.py).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import cama_context, flat_search

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _bits(t: torch.Tensor) -> np.ndarray:
    """bf16 tensors are stored as their uint16 bit patterns (numpy has no bfloat16)."""
    if t.dtype == torch.bfloat16:
        return t.contiguous().view(torch.int16).numpy().view(np.uint16)
    return t.contiguous().numpy()


def make_cama(dtype: torch.dtype, name: str, b=3, K=4, L=5, C=64, seed=11):
    g = torch.Generator().manual_seed(seed)
    ref = torch.randn(b, K, L, C, generator=g).to(dtype)
    tgt = torch.randn(b, L, C, generator=g).to(dtype)
    cond = torch.randn(b, (K + 1) * L, C, generator=g).to(dtype)
    sos = (torch.randn(1, L, C, generator=g) / C ** 0.5).to(dtype)   # module.py:258-259
    x, mask, pos = cama_context.reference_context(ref, tgt, cond.clone(), sos)
    assert x.dtype == dtype and x.shape == (b, (K + 1) * L, C)
    np.savez_compressed(OUT / name, ref_feats=_bits(ref), target=_bits(tgt), cond=_bits(cond),
                        sos=_bits(sos), x=_bits(x), mask=mask.numpy(), pos_table=pos.numpy(),
                        dtype=str(dtype))


def make_cama_loss(name="cama_loss_f32.npz", b=2, K=3, L=4, C=32, heads=4, ff=64, layers=2, seed=31):
    """Loss values and the sos_token gradient of the reference's real training / validation forward
    (module.py:305-311, 333-351) on known features and a seeded toy transformer whose weights are stored."""
    g = torch.Generator().manual_seed(seed)
    ref = torch.randn(b, K, L, C, generator=g)
    tgt = torch.randn(b, L, C, generator=g)
    cond = torch.randn(b, (K + 1) * L, C, generator=g)
    sos = torch.randn(1, L, C, generator=g) / C ** 0.5
    enc = cama_context.tiny_encoder(C, heads, ff, layers, seed)
    out = {"ref_feats": ref.numpy(), "target": tgt.numpy(), "cond": cond.numpy(), "sos": sos.numpy(),
           "shape": np.array([C, heads, ff, layers])}
    for k, v in enc.state_dict().items():
        out["w:" + k] = v.numpy()
    for ignore in (False, True):
        mse, smooth, grad = cama_context.reference_loss(ref, tgt, cond, sos, enc, ignore)
        tag = "val" if ignore else "train"
        out[f"{tag}_mse"], out[f"{tag}_smooth"], out[f"{tag}_sos_grad"] = np.float64(mse), np.float64(smooth), grad.numpy()
    np.savez_compressed(OUT / name, **out)


def make_retrieval(name="retrieval_small.npz", n=2000, dim=256, nq=16, k=12, seed=5):
    rng = np.random.default_rng(seed)
    cent = flat_search.normalise_rows(rng.standard_normal((32, dim)).astype(np.float32))
    db = cent[rng.integers(0, 32, n)] + 0.3 / np.sqrt(dim) * rng.standard_normal((n, dim)).astype(np.float32)
    db = flat_search.normalise_rows(db)
    db[7] = db[3]                      # exact duplicate rows: tie must resolve to the lower index
    db[1500] = db[3]
    src = rng.integers(0, n, nq)
    q = (db[src] + 0.05 / np.sqrt(dim) * rng.standard_normal((nq, dim)).astype(np.float32))
    q = (q * rng.uniform(5, 15, (nq, 1))).astype(np.float32)   # un-normalised queries
    q[0] = db[3]                       # a query identical to a (duplicated) row: distance 0
    groups = (np.arange(n) // 3).astype(np.int32)
    excl = groups[src].astype(np.int32)
    excl[1] = -1
    out = {"db": db, "queries": q, "row_group": groups, "exclude_group": excl, "k": k}
    for metric in flat_search.METRICS:
        d, i = flat_search.flat_search(db, q, k, metric)
        out[f"{metric}_dist"], out[f"{metric}_idx"] = d, i
    d, i = flat_search.flat_search(db, q, k, "l2", groups, excl, prefilter=False)
    out["l2_post_dist"], out["l2_post_idx"] = d, i
    d, i = flat_search.flat_search(db, q, k, "l2", groups, excl, prefilter=True)
    out["l2_pre_dist"], out["l2_pre_idx"] = d, i
    np.savez_compressed(OUT / name, **out)


def rag_class_table(n=1500, dim=256, seed=21) -> dict:
    """The in-memory RAG table (schema of tools/build_rag_database.py:35-45) behind
    rag_reference_class.json; tests rebuild it from the same seed."""
    rng = np.random.default_rng(seed)
    cent = flat_search.normalise_rows(rng.standard_normal((24, dim)).astype(np.float32))
    emb = flat_search.normalise_rows(cent[rng.integers(0, 24, n)] +
                                     0.4 / np.sqrt(dim) * rng.standard_normal((n, dim)).astype(np.float32))
    img = flat_search.normalise_rows(rng.standard_normal((n, dim)).astype(np.float32))
    return {"text": np.array([f"caption {j}" for j in range(n)]), "text_embedding": emb, "image_embedding": img,
            "id": np.arange(n), "uid": np.array([f"u{j}" for j in range(n)]), "dataset": np.array(["openvid"] * n),
            "video": np.array([f"clip_{j // 3:05d}.mp4" for j in range(n)]),
            "start_sec": (np.arange(n) % 3) * 2.0, "end_sec": (np.arange(n) % 3) * 2.0 + 2.0}


def rag_class_calls(table: dict) -> list[dict]:
    """The calls recorded in rag_reference_class.json: (method, kwargs) with array arguments given
    as (column, row, scale) so that the file stays small and tests can rebuild them."""
    calls = []
    for j in (5, 77, 640, 1201, 1499):       # the prepare_annotations pattern, datamodule.py:231-236
        calls.append({"method": "text_search", "text": ("text_embedding", j, 7.5), "top_k": 12,
                      "where": f'video != "{table["video"][j]}"', "select": ["video", "start_sec", "end_sec"]})
    calls.append({"method": "text_search", "text": ("text_embedding", 33, 1.0)})                     # all defaults
    calls.append({"method": "text_search", "text": ("text_embedding", 34, 2.0), "top_k": 5, "select": ["id", "uid"],
                  "output_format": "list"})
    calls.append({"method": "text_search", "text": ("text_embedding", 35, 2.0), "top_k": 4, "select": ["id", "video"],
                  "output_format": "pandas"})
    calls.append({"method": "text_search", "text": ("text_embedding", 36, 2.0), "top_k": 4, "select": ["id", "dataset"],
                  "output_format": "pyarrow"})
    calls.append({"method": "vector_search", "vector": ("image_embedding", 90, 3.0), "vector_column_name": "image_embedding",
                  "top_k": 6, "select": ["id"]})
    calls.append({"method": "image_search", "image_embedding": ("image_embedding", 91, 0.5), "top_k": 3,
                  "where": f'video != "{table["video"][91]}"', "select": ["id", "video"]})
    for j in (100, 900):                      # datamodule.py:239-245 with K = 9
        calls.append({"method": "text_image_search", "text": ("text_embedding", j, 4.0),
                      "image_embedding": ("image_embedding", j, 2.0), "top_k": (21, 9),
                      "where": f'video != "{table["video"][j]}"', "select": ["video", "start_sec", "end_sec"]})
    calls.append({"method": "text_image_search", "text": ("text_embedding", 7, 1.0),
                  "image_embedding": ("image_embedding", 8, 1.0), "top_k": (6, 3), "select": ["id"]})
    calls.append({"method": "text_search", "text": ("text_embedding", 1, 1.0), "output_format": "csv"})  # ValueError
    return calls


def rag_class_kwargs(table: dict, call: dict) -> dict:
    kw = {k: v for k, v in call.items() if k != "method"}
    for key in ("text", "vector", "image_embedding"):
        if key in kw:
            col, row, scale = kw[key]
            kw[key] = (table[col][row] * np.float32(scale)).astype(np.float32)
    if "top_k" in kw and isinstance(kw["top_k"], list):
        kw["top_k"] = tuple(kw["top_k"])
    return kw


def rag_class_summary(result) -> dict:
    """A JSON-able digest of what a RAGDatabase call returned, whatever the output format."""
    import pandas as pd
    import pyarrow as pa
    if isinstance(result, pd.DataFrame):
        kind, recs = "pandas", result.to_dict("records")
    elif isinstance(result, pa.Table):
        kind, recs = "pyarrow", result.to_pylist()
    else:
        kind, recs = "list", list(result)
    out = {"kind": kind, "keys": sorted(recs[0]) if recs else [], "records": []}
    for r in recs:
        out["records"].append({k: (float(v) if isinstance(v, (float, np.floating)) else
                                   int(v) if isinstance(v, (int, np.integer)) else
                                   v if isinstance(v, str) else None)      # vectors are not stored
                               for k, v in r.items()})
    return out


def make_rag_class(name="rag_reference_class.json"):
    """Runs the reference's REAL RAGDatabase class (src/data/rag.py, unmodified) over the LanceDB
    stand-in of oracle/fake_lancedb.py and records what it returns: pins argument plumbing, column
    selection, the two-stage text->image search and result formatting of the drop-in class."""
    import json
    from . import fake_lancedb
    table = rag_class_table()
    ref = fake_lancedb.reference_rag_database(table)
    recorded = []
    for call in rag_class_calls(table):
        kw = rag_class_kwargs(table, call)
        try:
            res = rag_class_summary(getattr(ref, call["method"])(**kw))
        except ValueError as e:
            res = {"raises": "ValueError", "message": str(e)}
        recorded.append({"call": call, "result": res})
    (OUT / name).write_text(json.dumps({"table": {"n": 1500, "dim": 256, "seed": 21}, "calls": recorded}, indent=1))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    make_rag_class()
    make_cama(torch.bfloat16, "cama_context_bf16.npz")
    make_cama(torch.float32, "cama_context_f32.npz")
    make_retrieval()
    make_cama_loss()
    for f in sorted(OUT.glob("*.npz")) + sorted(OUT.glob("*.json")):
        print(f, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
