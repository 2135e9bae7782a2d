"""Row-sharded retrieval over the GPUs of one NVSwitch box: one process per GPU.

Partition: shard g owns the contiguous global rows
[g * rows_per_shard, (g+1) * rows_per_shard); embedding store and feature table use the same
split. A search is: every rank scans its shard for the (replicated) query batch -> ONE
all-gather of the packed per-rank record {distance f32, group i32, index i64}[nq, k] over
NCCL/NVLink -> every rank merges G*k -> k (mrag_merge_topk; ties by lowest global index; the
`video != x` post-filter is applied after the global top-k, exactly where a single table
would apply it). Retrieved feature rows are then read where they live through peer-mapped
pointers (CUDA IPC) by the gather kernel itself — no second collective.

The search / merge callables are injectable so the exchange and layout logic can run under
`gloo` on CPU in tests (with the oracle standing in for the kernels); the defaults are the
CUDA paths.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _cabi
from ._cabi import check
from .store import EmbeddingStore, FeatureTable, SearchResult, merge_topk


def shard_range(n_rows: int, n_shards: int, rank: int) -> tuple[int, int, int]:
    """(rows_per_shard, first_row, last_row_exclusive) of `rank` for an n_rows table."""
    rps = (n_rows + n_shards - 1) // n_shards
    lo = min(n_rows, rank * rps)
    hi = min(n_rows, lo + rps)
    return rps, lo, hi


@dataclass
class PackedLayout:
    """Byte layout of one rank's record block: dist f32 | group i32 | idx i64, each [nq, k]."""
    nq: int
    k: int

    @property
    def n(self) -> int:
        return self.nq * self.k

    @property
    def off_dist(self) -> int:
        return 0

    @property
    def off_group(self) -> int:
        return 4 * self.n

    @property
    def off_idx(self) -> int:
        return 8 * self.n

    @property
    def nbytes(self) -> int:
        return 16 * self.n

    def views(self, buf: torch.Tensor, n_blocks: int):
        """Typed strided views [n_blocks, nq, k] into a uint8 buffer of n_blocks records."""
        assert buf.dtype == torch.uint8 and buf.numel() == n_blocks * self.nbytes
        b = buf.view(n_blocks, self.nbytes)
        d = b[:, self.off_dist:self.off_group].view(torch.float32).view(n_blocks, self.nq, self.k)
        g = b[:, self.off_group:self.off_idx].view(torch.int32).view(n_blocks, self.nq, self.k)
        i = b[:, self.off_idx:].view(torch.int64).view(n_blocks, self.nq, self.k)
        return d, g, i


class PeerExchange:
    """Exchange buffers for the fused cross-GPU merge (mrag_search_sharded): one cudaMalloc'ed,
    zero-initialised block per rank, mapped into every peer with CUDA IPC, plus the device table
    of the `world` base pointers and the call epoch. Collective constructor (all ranks)."""

    def __init__(self, rank: int, world: int, device: torch.device, nq_cap: int = 4096, k_cap: int = 32,
                 group: dist.ProcessGroup | None = None):
        from .store import _view
        lib = _cabi.load()
        self.rank, self.world, self.device = rank, world, device
        self.nq_cap, self.k_cap = int(nq_cap), int(k_cap)
        nbytes = int(lib.mrag_exchange_bytes(world, self.nq_cap, self.k_cap))
        if nbytes == 0:
            raise ValueError("bad exchange shape (world <= 8, k_cap <= 32)")
        self._lib, self._own, self._opened, self.epoch = lib, None, [], 0
        # Every step that can fail locally is followed by a collective vote, so that all ranks
        # raise (or none does) and nobody is left waiting in a collective.
        mine, err = None, None
        try:
            ptr = C.c_void_p()
            check(lib.mrag_device_alloc(device.index, nbytes, C.byref(ptr)))
            self._own = ptr
            self.buf = _view(ptr.value, (nbytes // 4,), torch.int32, device, self)
            self.buf.zero_()
            torch.cuda.synchronize(device)
            handle = (C.c_ubyte * 64)()
            check(lib.mrag_ipc_export(ptr, handle))
            mine = bytes(handle)
        except Exception as e:  # noqa: BLE001
            err = e
        handles: list = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        ptrs = []
        if err is None and all(h is not None for h in handles):
            try:
                for r, h in enumerate(handles):
                    if r == rank:
                        ptrs.append(self._own.value)
                        continue
                    p = C.c_void_p()
                    check(lib.mrag_ipc_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(p)))
                    self._opened.append(p)
                    ptrs.append(p.value)
            except Exception as e:  # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("a peer rank could not export its exchange buffer")
        oks: list = [None] * world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            self.close()
            raise RuntimeError(f"peer exchange setup failed on rank(s) {[r for r, o in enumerate(oks) if not o]}"
                               + (f": {err}" if err is not None else ""))
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=device)
        dist.barrier(group=group)   # every rank has zeroed and mapped before the first use

    def next(self, timeout_ms: int = 0) -> "_cabi.Exchange":
        """Descriptor of the next call (epoch + 1). Every rank must issue the same call sequence;
        validate anything that can fail locally BEFORE taking a descriptor."""
        self.epoch += 1
        return _cabi.Exchange(world=self.world, rank=self.rank, nq_cap=self.nq_cap, k_cap=self.k_cap,
                              epoch=self.epoch, timeout_ms=int(timeout_ms), bufs_dev=self.table.data_ptr())

    def reset(self, group: dist.ProcessGroup | None = None) -> None:
        """Collective re-synchronisation after a failed call (a timed-out exchange leaves the ranks'
        epochs out of step): drain, zero every buffer, restart the epoch count."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)
        self.buf.zero_()
        self.epoch = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    def close(self) -> None:
        for p in getattr(self, "_opened", []):
            self._lib.mrag_ipc_close(p)
        self._opened = []
        if getattr(self, "_own", None):
            self._lib.mrag_device_free(self.device.index, self._own)
            self._own = None


class ShardedRetriever:
    def __init__(self, store: EmbeddingStore | None, rank: int, world: int, rows_per_shard: int,
                 group: dist.ProcessGroup | None = None, local_search=None, merge=None,
                 device: torch.device | None = None, exchange: PeerExchange | None = None):
        """exchange: a PeerExchange makes the cross-GPU merge part of the search's last kernel
        (NVLink peer stores + flags); without it the per-shard records travel in one NCCL
        all-gather followed by a merge kernel."""
        self.store, self.rank, self.world = store, rank, world
        self.rows_per_shard = int(rows_per_shard)
        self.group = group
        self.device = device if device is not None else store.device
        self._local_search = local_search or self._cuda_search
        self._merge = merge or self._cuda_merge
        self.exchange = exchange
        self._bufs: dict[tuple[int, int], tuple[torch.Tensor, torch.Tensor]] = {}

    # default (product) implementations ------------------------------------------------------
    def _cuda_search(self, queries, k, metric, path, refine, exclude_group, filter_mode, index_base, out,
                     certify=False, list_len=0):
        return self.store.search(queries, k, metric=metric, path=path, refine=refine,
                                 exclude_group=exclude_group, filter_mode=filter_mode,
                                 index_base=index_base, out=out, certify=certify, list_len=list_len)

    @staticmethod
    def _cuda_merge(dist_v, idx_v, grp_v, k_out, exclude_group, filter_mode, stride):
        return merge_topk(dist_v, idx_v, grp_v, k_out, exclude_group, filter_mode, stride)

    def _validate(self, nq: int, k: int, path: str) -> None:
        """Everything that can fail on ONE rank only is checked before the exchange epoch moves, so
        a bad call raises on every rank instead of leaving the others waiting for a peer."""
        if nq < 1 or not (1 <= k <= 32):
            raise ValueError(f"need nq >= 1 and 1 <= k <= 32 (got nq {nq}, k {k})")
        if path not in ("auto", "stream_f32", "stream_bf16", "tensor_bf16"):
            raise ValueError(f"unknown path {path!r}")
        if path.startswith("stream") and self.store is not None and self.store.dim not in (256, 512, 768, 1024):
            raise ValueError("streaming paths need dim in {256, 512, 768, 1024}")

    def search(self, queries: torch.Tensor, k: int, *, metric: str = "l2", path: str = "auto",
               refine: int = 0, exclude_group: torch.Tensor | None = None,
               filter_mode: str = "post", certify: bool = False, list_len: int = 0) -> SearchResult:
        """Same contract as EmbeddingStore.search, over the union of all shards. Every rank
        passes the same queries and gets the same (global-index) result; with certify the
        exactness margin of the GLOBAL result (see mrag.h) comes back in `.margin`."""
        extra = {**({"certify": True} if certify else {}), **({"list_len": list_len} if list_len else {})}
        if self.world == 1:   # one shard: the local search already is the answer (no exchange)
            return self._local_search(queries, k, metric, path, refine, exclude_group,
                                      filter_mode if exclude_group is not None else "none", 0, None, **extra)
        nq = queries.shape[0]
        self._validate(nq, k, path)
        if self.exchange is not None and nq <= self.exchange.nq_cap and k <= self.exchange.k_cap:
            return self.store.search(queries, k, metric=metric, path=path, refine=refine,
                                     exclude_group=exclude_group,
                                     filter_mode=filter_mode if exclude_group is not None else "none",
                                     index_base=self.rank * self.rows_per_shard,
                                     exchange=self.exchange.next(), certify=certify, list_len=list_len)
        lay = PackedLayout(nq, k)
        key = (nq, k)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty(lay.nbytes, dtype=torch.uint8, device=self.device),
                               torch.empty(self.world * lay.nbytes, dtype=torch.uint8, device=self.device))
        send, recv = self._bufs[key]
        d, g, i = lay.views(send, 1)
        local = SearchResult(d[0], i[0], g[0])
        # post-filter happens after the GLOBAL top-k; pre-filter can be applied per shard
        local_filter = "pre" if (filter_mode == "pre" and exclude_group is not None) else "none"
        local = self._local_search(queries, k, metric, path, refine,
                                   exclude_group if local_filter == "pre" else None, local_filter,
                                   self.rank * self.rows_per_shard, local, **extra)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        dv, gv, iv = lay.views(recv, self.world)
        out = self._merge(dv, iv, gv, k, exclude_group,
                          filter_mode if exclude_group is not None else "none", lay.nbytes)
        if certify:
            # the global k-th result is at least as good as any shard's own k-th, so the smallest
            # per-shard margin is a (conservative) margin of the merged result
            m = local.margin.contiguous()
            allm = torch.empty(self.world * nq, dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(allm, m, group=self.group)
            out.margin = allm.view(self.world, nq).amin(0)
        return out

    def search_host(self, queries, k: int, *, metric: str = "l2", path: str = "auto", refine: int = 0,
                    exclude_group=None, filter_mode: str = "post", certify: bool = False, list_len: int = 0,
                    reuse: bool = False):
        """Host buffers in, host buffers out (numpy), every rank with the same queries: the
        reference-facing entry of a row-sharded table. With a PeerExchange small calls replay ONE
        captured graph per rank (H2D copy, scan, fused select / exchange / merge writing straight to
        pinned host memory); otherwise it falls back to device tensors + the NCCL transport."""
        import numpy as np
        q = queries if (type(queries) is np.ndarray and queries.dtype == np.float32 and queries.flags.c_contiguous) \
            else np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        if self.world == 1:
            return self.store.search_host(q, k, metric=metric, path=path, refine=refine, exclude_group=exclude_group,
                                          filter_mode=filter_mode, certify=certify, list_len=list_len, reuse=reuse)
        self._validate(nq, k, path)
        if self.exchange is not None and nq <= self.exchange.nq_cap and k <= self.exchange.k_cap:
            return self.store.search_host(q, k, metric=metric, path=path, refine=refine,
                                          exclude_group=exclude_group, filter_mode=filter_mode,
                                          index_base=self.rank * self.rows_per_shard, certify=certify,
                                          exchange=self.exchange.next(), list_len=list_len, reuse=reuse)
        ex = None if exclude_group is None else torch.from_numpy(np.ascontiguousarray(exclude_group, dtype=np.int32)).to(self.device)
        r = self.search(torch.from_numpy(q).to(self.device), k, metric=metric, path=path, refine=refine,
                        exclude_group=ex, filter_mode=filter_mode, certify=certify, list_len=list_len)
        out = (r.distance.cpu().numpy(), r.index.cpu().numpy(), r.group.cpu().numpy())
        return out + ((r.margin.cpu().numpy(),) if certify else ())


# --- peer-mapped feature tables ------------------------------------------------------------------
def alloc_feature_block(rows: int, L: int, Cdim: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """cudaMalloc'ed (IPC-exportable) [rows, L, C] block wrapped as a torch tensor."""
    from .store import _view
    lib = _cabi.load()
    esz = 2 if dtype == torch.bfloat16 else 4
    ptr = C.c_void_p()
    check(lib.mrag_device_alloc(device.index, rows * L * Cdim * esz, C.byref(ptr)))

    class _Owner:
        def __init__(self, p, dev):
            self.p, self.dev = p, dev

        def __del__(self):
            try:
                lib.mrag_device_free(self.dev, self.p)
            except Exception:
                pass

    return _view(ptr.value, (rows, L, Cdim), dtype, device, _Owner(ptr, device.index))


def open_peer_tables(table: FeatureTable, group: dist.ProcessGroup | None = None) -> FeatureTable:
    """Exchange CUDA IPC handles of the local feature blocks and map every peer block, so the
    gather kernel can read any global row over NVLink. The local block must come from
    alloc_feature_block (IPC handles name whole cudaMalloc allocations)."""
    if table.n_shards == 1:
        return table
    lib = _cabi.load()
    handle = (C.c_ubyte * 64)()
    check(lib.mrag_ipc_export(C.c_void_p(table.local.data_ptr()), handle))
    mine = bytes(handle)
    handles: list = [None] * table.n_shards
    dist.all_gather_object(handles, mine, group=group)
    for r, h in enumerate(handles):
        if r == table.shard_rank:
            continue
        buf = (C.c_ubyte * 64).from_buffer_copy(h)
        ptr = C.c_void_p()
        check(lib.mrag_ipc_open(buf, C.byref(ptr)))
        table.set_peer_ptr(r, ptr.value)
    return table
