"""ORACLE (test infrastructure) — writes the committed fixtures under tests/golden/.

Run in the build container (the only place /root/reference exists):
    python -m oracle.make_golden

cama_context_{bf16,f32}.npz   inputs + the (x, mask) captured from the reference's REAL
                              ActionTransformer.batch_forward (oracle/cama_context.py) —
                              these pin the gather/context kernel to the reference itself.
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


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    make_cama(torch.bfloat16, "cama_context_bf16.npz")
    make_cama(torch.float32, "cama_context_f32.npz")
    make_retrieval()
    for f in sorted(OUT.glob("*.npz")):
        print(f, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
