"""Offline motion-feature table: what `ActionTransformer.encode_vision`
(src/projects/condition/module.py:264-268: VideoMAE -> Resampler, 25 x 1024 per clip) yields for
every clip of the RAG table, computed once and stored row-aligned with the embedding store, so that
retrieval results index it directly (K4 gather) instead of decoding and re-encoding K clips per sample.

On disk (a directory): `features.npy` uint16 [n_rows, L, C] holding bfloat16 bit patterns
(memory-mapped, so a 1 M-clip table (51 GB) is written and later sharded without being resident),
`uncond_row.npy` uint16 [L, C] = encode_vision(all-zero clip) — the row a dropped / unreadable /
never-written reference maps to (src/data/dataset.py:292,305-310; module.py:327-329), `written.npy`
bool [n_rows], `meta.json`.

The encoders themselves (VideoMAE, Resampler and their checkpoints) stay the reference's: this
module only drives `model.encode_vision` and owns the table format. Pure PyTorch / numpy.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Iterable

import numpy as np
import torch

FORMAT_VERSION = 1


def _bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.bfloat16).contiguous().cpu().view(torch.int16).numpy().view(np.uint16)


def _from_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(torch.bfloat16)


class FeatureTableWriter:
    def __init__(self, path, n_rows: int, tokens: int = 25, width: int = 1024):
        self.path = Path(path)
        self.path.mkdir(parents=True, exist_ok=True)
        self.n_rows, self.L, self.C = int(n_rows), int(tokens), int(width)
        self._feat = np.lib.format.open_memmap(self.path / "features.npy", mode="w+", dtype=np.uint16,
                                               shape=(self.n_rows, self.L, self.C))
        self._written = np.zeros(self.n_rows, dtype=bool)
        self._uncond: np.ndarray | None = None

    def write(self, row_ids, features: torch.Tensor) -> None:
        """features [n, L, C] (any float dtype / device) for table rows `row_ids` [n]."""
        ids = np.asarray(torch.as_tensor(row_ids).cpu(), dtype=np.int64).reshape(-1)
        if features.ndim != 3 or tuple(features.shape) != (len(ids), self.L, self.C):
            raise ValueError(f"features must be [{len(ids)}, {self.L}, {self.C}], got {tuple(features.shape)}")
        if len(ids) and (ids.min() < 0 or ids.max() >= self.n_rows):
            raise IndexError("row id outside the table")
        self._feat[ids] = _bf16_bits(features)
        self._written[ids] = True

    def set_uncond_row(self, row: torch.Tensor) -> None:
        if tuple(row.shape) != (self.L, self.C):
            raise ValueError(f"uncond row must be [{self.L}, {self.C}]")
        self._uncond = _bf16_bits(row)

    def close(self) -> dict:
        """Rows never written (failed decodes) become the uncond row, as an all-zero clip would."""
        if self._uncond is None:
            raise RuntimeError("set_uncond_row() was never called")
        missing = np.flatnonzero(~self._written)
        for lo in range(0, len(missing), 4096):
            self._feat[missing[lo:lo + 4096]] = self._uncond
        self._feat.flush()
        np.save(self.path / "uncond_row.npy", self._uncond)
        np.save(self.path / "written.npy", self._written)
        meta = {"format": FORMAT_VERSION, "n_rows": self.n_rows, "tokens": self.L, "width": self.C,
                "dtype": "bfloat16", "missing_rows": int(len(missing))}
        (self.path / "meta.json").write_text(json.dumps(meta))
        del self._feat
        return meta


@torch.no_grad()
def build_feature_table(model, clips: Iterable, n_rows: int, path, tokens: int = 25, width: int = 1024,
                        device: str | torch.device = "cuda", autocast_dtype=torch.bfloat16) -> dict:
    """Drive `model.encode_vision` over the RAG table's clips.

    model: the reference ActionTransformer (anything with encode_vision([b, k, T, C, H, W]) ->
    [b, k, L, C]); clips: iterable of (row_ids [n], videos [n, T, C, H, W]) batches in any order —
    decoding is the reference's own dataset code (src/data/dataset.py:205-240). Row ids are the
    embedding store's row numbers (the `id` column of tools/build_rag_database.py:40)."""
    dev = torch.device(device)
    w = FeatureTableWriter(path, n_rows, tokens, width)
    first = None
    use_amp = dev.type == "cuda" and autocast_dtype is not None
    for row_ids, videos in clips:
        videos = videos.to(dev, non_blocking=True)
        if first is None:
            first = videos[:1]
        with torch.autocast("cuda", dtype=autocast_dtype, enabled=use_amp):
            feats = model.encode_vision(videos[:, None])[:, 0]
        w.write(row_ids, feats)
    if first is None:
        raise ValueError("no clips")
    with torch.autocast("cuda", dtype=autocast_dtype, enabled=use_amp):
        w.set_uncond_row(model.encode_vision(torch.zeros_like(first)[:, None])[0, 0])   # module.py:327-329
    return w.close()


def load_feature_rows(path, rows: tuple[int, int] | None = None) -> tuple[torch.Tensor, torch.Tensor, dict]:
    """-> (features bf16 [hi-lo, L, C] on the host, uncond_row bf16 [L, C], meta); `rows` = the
    [lo, hi) shard of a row-sharded table (parallel.shard_range). Only that slice is read."""
    p = Path(path)
    meta = json.loads((p / "meta.json").read_text())
    if meta.get("format") != FORMAT_VERSION:
        raise ValueError(f"unknown feature table format {meta.get('format')!r}")
    mm = np.load(p / "features.npy", mmap_mode="r")
    lo, hi = (0, meta["n_rows"]) if rows is None else (max(0, int(rows[0])), min(meta["n_rows"], int(rows[1])))
    return _from_bits(mm[lo:hi]), _from_bits(np.load(p / "uncond_row.npy")), meta
