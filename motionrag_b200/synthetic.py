"""Seeded synthetic data of the shapes named in SURVEY.md §8(d) (there is no network for
OpenVid-1M captions or checkpoints). Generation is chunked: with device="cuda" nothing large
ever touches the host. CPU and CUDA generators produce different streams for the same seed,
so parity tests generate once and hand the same tensor to both sides.

  database  [N, dim] fp32, L2-normalised (tools/build_rag_database.py:31-37 stores normalised
            gte-base-en-v1.5 vectors): "iid" = N(0,1) rows (worst-case near-ties), "clustered"
            = 4096 unit centroids + N(0, 0.3^2/dim) noise (score spread like real captions)
  queries   [Q, dim] fp32, NOT normalised (src/data/datamodule.py:300-302): a database row
            plus noise, scaled by U(5, 15); or plain N(0,1) for "iid"
  groups    int32 [N] = row // 3 (about three clips per source video) for `video != own`
  features  [N, 25, 1024] bf16 N(0,1) motion tokens (Resampler output is LayerNorm-scaled,
            src/projects/condition/encoders/resampler.py:142,166)
"""
from __future__ import annotations

import torch

CHUNK_ROWS = 1 << 16


def _gen(device, seed: int) -> torch.Generator:
    return torch.Generator(device=device).manual_seed(int(seed))


def database_chunks(n_rows: int, dim: int = 768, kind: str = "clustered", seed: int = 0,
                    device: str | torch.device = "cuda", first_row: int = 0, n_centroids: int = 4096):
    """Yield (row_offset, rows[<=CHUNK_ROWS, dim] fp32 normalised) for global rows
    [first_row, first_row + n_rows); a chunk's values depend only on (seed, global chunk id),
    so shards of one table can be generated independently on different GPUs."""
    device = torch.device(device)
    cent = None
    if kind == "clustered":
        cent = torch.nn.functional.normalize(
            torch.randn(n_centroids, dim, generator=_gen(device, seed * 7919 + 1), device=device), dim=-1)
    elif kind != "iid":
        raise ValueError(kind)
    assert first_row % CHUNK_ROWS == 0 or n_rows <= CHUNK_ROWS, "shards must start on chunk boundaries"
    done = 0
    while done < n_rows:
        gchunk = (first_row + done) // CHUNK_ROWS
        rows = min(CHUNK_ROWS, n_rows - done)
        g = _gen(device, seed * 1_000_003 + 17 * gchunk + 5)
        x = torch.randn(CHUNK_ROWS, dim, generator=g, device=device)
        if cent is not None:
            which = torch.randint(0, n_centroids, (CHUNK_ROWS,), generator=g, device=device)
            x = cent[which] + (0.3 / dim ** 0.5) * x
        x = torch.nn.functional.normalize(x[:rows], dim=-1)
        yield done, x
        done += rows


def database(n_rows: int, dim: int = 768, kind: str = "clustered", seed: int = 0,
             device: str | torch.device = "cpu") -> torch.Tensor:
    out = torch.empty(n_rows, dim, dtype=torch.float32, device=device)
    for off, rows in database_chunks(n_rows, dim, kind, seed, device):
        out[off:off + rows.shape[0]] = rows
    return out


def fill_store(store, n_rows: int, kind: str = "clustered", seed: int = 0, first_row: int = 0) -> None:
    """Generate rows on the store's GPU and append them (rows are already unit-norm; the
    store re-normalises, which is idempotent up to rounding)."""
    for _, rows in database_chunks(n_rows, store.dim, kind, seed, store.device, first_row):
        store.append(rows, normalise=True)


def groups(n_rows: int, first_row: int = 0, device: str | torch.device = "cpu") -> torch.Tensor:
    return ((torch.arange(n_rows, device=device) + first_row) // 3).to(torch.int32)


def queries_from_rows(rows: torch.Tensor, seed: int = 1, noise: float = 0.05) -> torch.Tensor:
    """Un-normalised queries near given database rows: (row + noise) * U(5, 15)."""
    g = _gen(rows.device, seed)
    q = rows + (noise / rows.shape[1] ** 0.5) * torch.randn(rows.shape, generator=g, device=rows.device)
    scale = 5 + 10 * torch.rand(rows.shape[0], 1, generator=g, device=rows.device)
    return (q * scale).to(torch.float32).contiguous()


def queries_iid(nq: int, dim: int = 768, seed: int = 1, device: str | torch.device = "cpu") -> torch.Tensor:
    return torch.randn(nq, dim, generator=_gen(torch.device(device), seed), device=device)


def features(n_rows: int, L: int = 25, Cdim: int = 1024, dtype=torch.bfloat16, seed: int = 2,
             device: str | torch.device = "cuda", first_row: int = 0, out: torch.Tensor | None = None):
    """[n_rows, L, C] feature rows, chunked so a 50 GB table never needs fp32 staging."""
    device = torch.device(device)
    if out is None:
        out = torch.empty(n_rows, L, Cdim, dtype=dtype, device=device)
    step = 4096
    for s in range(0, n_rows, step):
        r = min(step, n_rows - s)
        g = _gen(device, seed * 999_983 + (first_row + s) // step)
        out[s:s + r] = torch.randn(step, L, Cdim, generator=g, device=device)[:r].to(dtype)
    return out
