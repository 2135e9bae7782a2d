"""HBM-resident stores of the retrieval hot path.

`EmbeddingStore`  — the text-embedding column of the RAG table written by the reference's
                    tools/build_rag_database.py:35-50 (fp32[dim], L2-normalised), held as an
                    fp32 master + bf16 shadow in device memory owned by libmrag.
`FeatureTable`    — the precomputed motion-feature rows [N, L, C] that stand in for the
                    reference's per-sample decode + VideoMAE + Resampler
                    (src/data/dataset.py:285-312, src/projects/condition/module.py:264-268).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi
from ._cabi import FILTER, METRIC, PATH, PlanInfo, SearchParams, StoreInfo, check


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@dataclass
class SearchResult:
    """Device tensors [nq, k]; unused slots hold (+inf, -1, -1)."""
    distance: torch.Tensor  # float32, ascending
    index: torch.Tensor     # int64 global row ids
    group: torch.Tensor     # int32 group id of each hit
    margin: torch.Tensor | None = None  # float32 [nq] exactness certificate (see EPS / mrag.h)


# |bf16 scan score - true q.d| <= EPS[path] * |q| |d| (worst case: every element rounds the same way
# and q is parallel to the rounding error): margin > EPS * max|d| PROVES the top-k is exact
EPS = {"stream_f32": 0.0, "stream_bf16": 2.0 ** -9, "tensor_bf16": 2.0 ** -8}
PATH_NAME = {1: "stream_f32", 2: "stream_bf16", 3: "tensor_bf16"}


def eps_typical(path: str, dim: int, sigmas: float = 6.0) -> float:
    """`sigmas` standard deviations of the same error when rounding errors behave like
    independent uniform noise and the vectors are not concentrated in a few coordinates:
    sigma = 2^-9 / sqrt(12) / sqrt(dim) per rounded operand (2.0e-5 |q| at dim 768) — about
    100x below the worst case. margin > eps_typical is a statistical, not a rigorous, pass."""
    if path == "stream_f32":
        return 0.0
    rounded = 2.0 if path == "tensor_bf16" else 1.0
    return sigmas * (2.0 ** -9) / (12.0 ** 0.5) / (dim ** 0.5) * (rounded ** 0.5)


def margin_threshold(path: str, dim: int, strict: bool, max_norm_deviation: float = 0.0) -> float:
    """What an exactness margin must exceed: the worst-case bound (strict) or 6 sigma of the rounding
    noise, both scaled by the largest row norm the store holds."""
    eps = EPS[path] if strict else eps_typical(path, dim)
    return eps * (1.0 + max(0.0, float(max_norm_deviation))) ** 0.5


class EmbeddingStore:
    def __init__(self, dim: int, capacity_rows: int, device: int | str | torch.device = 0):
        self._lib = _cabi.load()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise _cabi.MragError(-3, f"EmbeddingStore lives in HBM; got device {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else 0)
        self.dim = int(dim)
        h = C.c_void_p()
        check(self._lib.mrag_store_create(self.dim, int(capacity_rows), self.device.index, C.byref(h)))
        self._h = h
        self._ws: dict[int, torch.Tensor] = {}   # scratch per CUDA stream: searches on different streams may overlap
        self._need: dict[tuple, int] = {}        # workspace bytes per call shape (dropped by append / set_groups)
        self._info: StoreInfo | None = None
        self._host_calls: dict[tuple, tuple] = {}   # prebuilt parameter blocks + result arrays of small host calls

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.mrag_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- contents ---------------------------------------------------------------------------
    def info(self) -> StoreInfo:
        if self._info is None:
            inf = StoreInfo()
            check(self._lib.mrag_store_get_info(self._h, C.byref(inf)))
            self._info = inf
        return self._info

    def poll_error(self) -> None:
        """Raise if a kernel of this store reported a device-side error (a peer exchange that timed
        out) since the last poll. Call after synchronising the stream the search ran on."""
        code = C.c_int32()
        check(self._lib.mrag_store_poll_error(self._h, C.byref(code)))

    def __len__(self) -> int:
        return int(self.info().n_rows)

    def append(self, rows: torch.Tensor | np.ndarray, normalise: bool = True) -> None:
        """Append fp32 rows [n, dim] from host (numpy / CPU tensor) or device memory."""
        if isinstance(rows, np.ndarray):
            rows = np.ascontiguousarray(rows, dtype=np.float32)
            if not rows.flags.writeable:   # e.g. a read-only np.load(mmap_mode="r") slice
                rows = rows.copy()
            rows = torch.from_numpy(rows)
        rows = rows.to(torch.float32).contiguous()
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"rows must be [n, {self.dim}], got {tuple(rows.shape)}")
        on_dev = rows.is_cuda
        if on_dev and rows.device != self.device:
            rows = rows.to(self.device)
        self._need.clear()
        self._host_calls.clear()
        self._info = None
        check(self._lib.mrag_store_append(self._h, C.c_void_p(rows.data_ptr()), rows.shape[0],
                                          1 if on_dev else 0, 1 if normalise else 0,
                                          _stream_ptr(self.device)))
        if not on_dev:
            torch.cuda.current_stream(self.device).synchronize()  # pageable source must stay alive

    def set_groups(self, groups: torch.Tensor | np.ndarray) -> None:
        """int32 group id per row (rows of the same source video share an id)."""
        if isinstance(groups, np.ndarray):
            groups = torch.from_numpy(np.ascontiguousarray(groups, dtype=np.int32))
        groups = groups.to(torch.int32).contiguous()
        on_dev = groups.is_cuda
        self._need.clear()
        self._info = None
        check(self._lib.mrag_store_set_groups(self._h, C.c_void_p(groups.data_ptr()), groups.numel(),
                                              1 if on_dev else 0, _stream_ptr(self.device)))
        if not on_dev:
            torch.cuda.current_stream(self.device).synchronize()

    # -- shard files (fast reload; the bf16 shadow and norm statistics are rebuilt on load) ----
    def save(self, path) -> None:
        """Write this store (one shard) as <path>/{meta.json, rows_f32.npy[, groups.npy]}."""
        import json
        from pathlib import Path
        root = Path(path)
        root.mkdir(parents=True, exist_ok=True)
        inf = self.info()
        n = int(inf.n_rows)
        step = max(1, (256 << 20) // (self.dim * 4))
        out = np.lib.format.open_memmap(root / "rows_f32.npy", mode="w+", dtype=np.float32, shape=(n, self.dim))
        rows = self.rows_f32()
        for s0 in range(0, n, step):
            out[s0:s0 + step] = rows[s0:s0 + step].cpu().numpy()
        out.flush()
        if inf.has_groups:
            np.save(root / "groups.npy", _view(inf.groups_dev, (n,), torch.int32, self.device, self).cpu().numpy())
        (root / "meta.json").write_text(json.dumps({"format": "mrag-shard-1", "dim": self.dim, "n_rows": n,
                                                    "has_groups": bool(inf.has_groups)}))

    @classmethod
    def load(cls, path, device: int | str | torch.device = 0, capacity_rows: int | None = None) -> "EmbeddingStore":
        import json
        from pathlib import Path
        root = Path(path)
        meta = json.loads((root / "meta.json").read_text())
        if meta.get("format") != "mrag-shard-1":
            raise ValueError(f"{root} is not an mrag shard directory")
        n, dim = int(meta["n_rows"]), int(meta["dim"])
        st = cls(dim, max(capacity_rows or n, n, 1), device)
        rows = np.load(root / "rows_f32.npy", mmap_mode="r")
        step = max(1, (256 << 20) // (dim * 4))
        for s0 in range(0, n, step):
            st.append(np.ascontiguousarray(rows[s0:s0 + step]), normalise=False)
        if meta.get("has_groups"):
            st.set_groups(np.load(root / "groups.npy"))
        return st

    def rows_f32(self) -> torch.Tensor:
        """Zero-copy view of the fp32 master rows (for tests / re-use by torch code)."""
        inf = self.info()
        return _view(inf.rows_f32_dev, (int(inf.n_rows), self.dim), torch.float32, self.device, self)

    def rows_bf16(self) -> torch.Tensor:
        inf = self.info()
        return _view(inf.rows_bf16_dev, (int(inf.n_rows), self.dim), torch.bfloat16, self.device, self)

    # -- search -----------------------------------------------------------------------------
    def _params(self, k, metric, path, refine, filter_mode, index_base, list_len=0) -> SearchParams:
        return SearchParams(k=int(k), metric=METRIC[metric], path=PATH[path], refine=int(refine),
                            filter_mode=FILTER[filter_mode], list_len=int(list_len), index_base=int(index_base),
                            out_margin=None)

    def plan(self, nq: int, params: SearchParams | None = None, **kw) -> PlanInfo:
        """How a search of nq queries would run (path, grid, candidates, workspace, work)."""
        if params is None:
            params = self._params(kw.get("k", 12), kw.get("metric", "l2"), kw.get("path", "auto"),
                                  kw.get("refine", 0), kw.get("filter_mode", "none"),
                                  kw.get("index_base", 0), kw.get("list_len", 0))
        info = PlanInfo()
        check(self._lib.mrag_search_plan(self._h, int(nq), C.byref(params), C.byref(info)))
        return info

    def search(self, queries: torch.Tensor, k: int, *, metric: str = "l2", path: str = "auto",
               refine: int = 0, exclude_group: torch.Tensor | None = None,
               filter_mode: str = "post", index_base: int = 0,
               out: SearchResult | None = None, timings: list | None = None,
               exchange=None, certify: bool = False, list_len: int = 0) -> SearchResult:
        """Device-resident search: queries [nq, dim] fp32 on this store's GPU -> SearchResult.

        Asynchronous on the current stream; nothing is copied to the host.
        """
        if queries.device != self.device or queries.dtype != torch.float32 or not queries.is_contiguous():
            raise ValueError("queries must be a contiguous float32 tensor on the store's device")
        nq = queries.shape[0]
        if exclude_group is None:
            filter_mode = "none"
        elif exclude_group.device != self.device or exclude_group.dtype != torch.int32:
            raise ValueError("exclude_group must be int32 on the store's device")
        p = self._params(k, metric, path, refine, filter_mode, index_base, list_len)
        key = (nq, k, metric, path, refine, filter_mode, exchange is not None, list_len)
        need = self._need.get(key)
        if need is None:
            if exchange is not None and len(self) == 0:
                need = 1 << 20       # an empty shard still takes part in the exchange
            else:
                need = int(self.plan(nq, p).workspace_bytes)
            self._need[key] = need
        if certify:
            out_margin = torch.empty(nq, dtype=torch.float32, device=self.device)
            p.out_margin = out_margin.data_ptr()
        stream = _stream_ptr(self.device)
        ws = self._ws.get(stream.value or 0)
        if ws is None or ws.numel() < need:
            ws = self._ws[stream.value or 0] = torch.empty(need, dtype=torch.uint8, device=self.device)
        if out is None:
            out = SearchResult(torch.empty((nq, k), dtype=torch.float32, device=self.device),
                               torch.empty((nq, k), dtype=torch.int64, device=self.device),
                               torch.empty((nq, k), dtype=torch.int32, device=self.device))
        args = (self._h, C.c_void_p(queries.data_ptr()), nq, C.byref(p),
                C.c_void_p(exclude_group.data_ptr()) if exclude_group is not None else None,
                C.c_void_p(out.distance.data_ptr()), C.c_void_p(out.index.data_ptr()),
                C.c_void_p(out.group.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), stream)
        if exchange is not None:   # row-sharded: cross-GPU merge fused into the last kernel
            check(self._lib.mrag_search_sharded(*args[:-1], C.byref(exchange), args[-1]))
        elif timings is None:
            check(self._lib.mrag_search(*args))
        else:  # synchronous, event-bracketed variant: appends (scan_ms, total_ms)
            a, b = C.c_float(), C.c_float()
            check(self._lib.mrag_search_timed(*args, C.byref(a), C.byref(b)))
            timings.append((a.value, b.value))
        if certify:
            out.margin = out_margin
        return out

    def rescore(self, queries: torch.Tensor, cand_index: torch.Tensor, k_out: int, metric: str = "l2"):
        """Exact fp32 distances of the rows `cand_index` [nq, kc] (int64, -1 = none, kc <= 64) of this
        store against `queries` [nq, dim]: the second stage of text_image_search
        (src/data/rag.py:118-128) with the candidate rows as the 'temporary table'. Returns
        (distance f32 [nq, k_out], index i64 [nq, k_out]) sorted by (distance, candidate position)."""
        if queries.device != self.device or queries.dtype != torch.float32 or not queries.is_contiguous():
            raise ValueError("queries must be a contiguous float32 tensor on the store's device")
        if cand_index.device != self.device or cand_index.dtype != torch.int64 or cand_index.ndim != 2 \
                or cand_index.shape[0] != queries.shape[0]:
            raise ValueError("cand_index must be int64 [nq, kc] on the store's device")
        cand_index = cand_index.contiguous()
        nq, kc = cand_index.shape
        dist = torch.empty((nq, k_out), dtype=torch.float32, device=self.device)
        idx = torch.empty((nq, k_out), dtype=torch.int64, device=self.device)
        check(self._lib.mrag_rescore_rows(self._h, C.c_void_p(queries.data_ptr()), nq,
                                          C.c_void_p(cand_index.data_ptr()), kc, METRIC[metric], int(k_out),
                                          C.c_void_p(dist.data_ptr()), C.c_void_p(idx.data_ptr()),
                                          _stream_ptr(self.device)))
        return dist, idx

    def search_host(self, queries: np.ndarray, k: int, *, metric: str = "l2", path: str = "auto",
                    refine: int = 0, exclude_group: np.ndarray | None = None,
                    filter_mode: str = "post", index_base: int = 0, certify: bool = False, exchange=None,
                    list_len: int = 0, reuse: bool = False):
        """Host-buffer search through `mrag_search_host` (copies inside, synchronous); with
        `exchange` (an mrag_exchange descriptor) the row-sharded variant whose last kernel merges the
        shards over peer memory (`mrag_search_sharded_host`).

        Returns (distance f32 [nq,k], index i64 [nq,k], group i32 [nq,k]) numpy arrays, plus the
        float32 [nq] exactness margin when certify=True.
        """
        q = queries if (type(queries) is np.ndarray and queries.dtype == np.float32 and queries.flags.c_contiguous) \
            else np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [nq, {self.dim}]")
        nq = q.shape[0]
        ex = None
        if exclude_group is None:
            filter_mode = "none"
        else:
            ex = exclude_group if (type(exclude_group) is np.ndarray and exclude_group.dtype == np.int32
                                   and exclude_group.flags.c_contiguous) else np.ascontiguousarray(exclude_group, dtype=np.int32)
        if reuse:
            # small repeated calls (the reference's one-query-per-call pattern): parameter block, result
            # arrays and argument pointers are built once per call shape; the RETURNED ARRAYS ARE REUSED by
            # the next call of the same shape (callers consume them immediately)
            key = (nq, k, metric, path, refine, filter_mode, index_base, list_len, certify, exchange is None)
            c = self._host_calls.get(key)
            if c is None:
                p = self._params(k, metric, path, refine, filter_mode, index_base, list_len)
                dist = np.empty((nq, k), dtype=np.float32)
                idx = np.empty((nq, k), dtype=np.int64)
                grp = np.empty((nq, k), dtype=np.int32)
                margin = np.empty(nq, dtype=np.float32) if certify else None
                if certify:
                    p.out_margin = margin.ctypes.data
                c = self._host_calls[key] = (p, C.byref(p), dist, idx, grp, margin, dist.ctypes.data, idx.ctypes.data,
                                             grp.ctypes.data, _stream_ptr(self.device))
            p, pref, dist, idx, grp, margin, dp, ip, gp, stream = c
            qp = q.ctypes.data
            xp = ex.ctypes.data if ex is not None else None
            if exchange is None:
                rc = self._lib.mrag_search_host(self._h, qp, nq, pref, xp, dp, ip, gp, stream)
            else:
                rc = self._lib.mrag_search_sharded_host(self._h, qp, nq, pref, xp, dp, ip, gp, C.byref(exchange), stream)
            if rc != 0:
                check(rc)
            return (dist, idx, grp, margin) if certify else (dist, idx, grp)
        p = self._params(k, metric, path, refine, filter_mode, index_base, list_len)
        dist = np.empty((nq, k), dtype=np.float32)
        idx = np.empty((nq, k), dtype=np.int64)
        grp = np.empty((nq, k), dtype=np.int32)
        margin = None
        if certify:
            margin = np.empty(nq, dtype=np.float32)
            p.out_margin = margin.ctypes.data
        args = (self._h, q.ctypes.data, nq, C.byref(p), ex.ctypes.data if ex is not None else None,
                dist.ctypes.data, idx.ctypes.data, grp.ctypes.data)
        if exchange is None:
            check(self._lib.mrag_search_host(*args, _stream_ptr(self.device)))
        else:
            check(self._lib.mrag_search_sharded_host(*args, C.byref(exchange), _stream_ptr(self.device)))
        if certify:
            return dist, idx, grp, margin
        return dist, idx, grp


def merge_topk(cand_dist: torch.Tensor, cand_idx: torch.Tensor, cand_group: torch.Tensor | None,
               k_out: int, exclude_group: torch.Tensor | None = None,
               filter_mode: str = "post", shard_stride_bytes: int = 0) -> SearchResult:
    """Merge per-shard results laid out [nshards, nq, k_in] (the all-gather layout); with
    shard_stride_bytes the three fields may be views into one packed per-rank record."""
    lib = _cabi.load()
    nshards, nq, k_in = cand_dist.shape
    dev = cand_dist.device
    out = SearchResult(torch.empty((nq, k_out), dtype=torch.float32, device=dev),
                       torch.empty((nq, k_out), dtype=torch.int64, device=dev),
                       torch.empty((nq, k_out), dtype=torch.int32, device=dev))
    if exclude_group is None:
        filter_mode = "none"
    check(lib.mrag_merge_topk(
        C.c_void_p(cand_dist.data_ptr()), C.c_void_p(cand_idx.data_ptr()),
        C.c_void_p(cand_group.data_ptr()) if cand_group is not None else None,
        int(shard_stride_bytes), nshards, nq, k_in, k_out,
        C.c_void_p(exclude_group.data_ptr()) if exclude_group is not None else None,
        FILTER[filter_mode], C.c_void_p(out.distance.data_ptr()), C.c_void_p(out.index.data_ptr()),
        C.c_void_p(out.group.data_ptr()), _stream_ptr(dev)))
    return out


def _view(ptr: int, shape, dtype, device, owner) -> torch.Tensor:
    """Wrap raw device memory owned by libmrag as a torch tensor (no copy)."""
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return torch.empty(shape, dtype=dtype, device=device)
    itemsize = torch.empty((), dtype=dtype).element_size()
    typestr = {torch.float32: "<f4", torch.bfloat16: "<u2", torch.int32: "<i4"}[dtype]

    class _Holder:
        pass

    holder = _Holder()
    holder.owner = owner
    holder.__cuda_array_interface__ = {
        "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
        "strides": None,
    }
    with torch.cuda.device(device):
        t = torch.as_tensor(holder, device=device)
    if dtype == torch.bfloat16:
        t = t.view(torch.bfloat16)
    assert t.element_size() == itemsize
    return t


class FeatureTable:
    """Motion-feature rows [n_rows, L, C] in HBM, optionally row-sharded over ranks.

    Global row r lives in shard r // rows_per_shard. `shard_ptrs` is the device array of
    base pointers handed to the gather kernel: the local block plus, after
    `parallel.open_peer_tables`, the peer-mapped blocks of the other ranks (read over NVLink).
    """

    def __init__(self, local: torch.Tensor, rows_per_shard: int | None = None, shard_rank: int = 0,
                 n_shards: int = 1, n_rows: int | None = None):
        if local.ndim != 3 or not local.is_cuda or not local.is_contiguous():
            raise ValueError("local feature block must be a contiguous CUDA tensor [rows, L, C]")
        if local.dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("feature dtype must be bfloat16 or float32")
        self.local = local
        self.device = local.device
        self.L, self.Cdim = int(local.shape[1]), int(local.shape[2])
        self.rows_per_shard = int(rows_per_shard if rows_per_shard is not None else local.shape[0])
        self.shard_rank, self.n_shards = int(shard_rank), int(n_shards)
        # rows the whole table really has (the last shard may be short): the gather kernel maps any
        # index at or beyond it to the uncond row instead of dereferencing it
        self.n_rows = int(n_rows) if n_rows is not None else (
            int(local.shape[0]) if self.n_shards == 1 else self.rows_per_shard * self.n_shards)
        if self.n_rows > self.rows_per_shard * self.n_shards:
            raise ValueError("n_rows exceeds n_shards * rows_per_shard")
        ptrs = [0] * self.n_shards
        ptrs[self.shard_rank] = local.data_ptr()
        self._peer_ptrs: list[int] = ptrs
        self.shard_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    def set_peer_ptr(self, shard: int, ptr: int) -> None:
        self._peer_ptrs[shard] = int(ptr)
        self.shard_ptrs = torch.tensor(self._peer_ptrs, dtype=torch.int64, device=self.device)

    @property
    def complete(self) -> bool:
        return all(p != 0 for p in self._peer_ptrs)
