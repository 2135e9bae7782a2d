"""ctypes binding of libmrag.so — the only way Python reaches the CUDA kernels.

There is deliberately no fallback: if the shared library is missing or a call fails the
caller gets an exception (`MragError`), never a CPU substitute.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "_lib" / "libmrag.so"

ABI_VERSION = 4

# enums of include/mrag.h
METRIC = {"l2": 0, "cosine": 1, "dot": 2}
PATH = {"auto": 0, "stream_f32": 1, "stream_bf16": 2, "tensor_bf16": 3}
FILTER = {"none": 0, "post": 1, "pre": 2}
STATUS = {0: "MRAG_OK", -1: "MRAG_ERR_ARG", -2: "MRAG_ERR_CUDA", -3: "MRAG_ERR_DEVICE",
          -4: "MRAG_ERR_CAPACITY", -5: "MRAG_ERR_UNSUPPORTED"}


class MragError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class StoreInfo(C.Structure):
    _fields_ = [("dim", C.c_int32), ("device", C.c_int32), ("n_rows", C.c_int64),
                ("capacity_rows", C.c_int64), ("has_groups", C.c_int32), ("sm_count", C.c_int32),
                ("rows_f32_dev", C.c_void_p), ("rows_bf16_dev", C.c_void_p),
                ("groups_dev", C.c_void_p), ("max_norm_deviation", C.c_float), ("reserved", C.c_int32),
                ("zero_rows", C.c_int64), ("row_bias_dev", C.c_void_p)]


class SearchParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("metric", C.c_int32), ("path", C.c_int32),
                ("refine", C.c_int32), ("filter_mode", C.c_int32), ("list_len", C.c_int32),
                ("index_base", C.c_int64), ("out_margin", C.c_void_p)]


class PlanInfo(C.Structure):
    _fields_ = [("path", C.c_int32), ("grid", C.c_int32), ("cands_per_query", C.c_int32),
                ("rerank", C.c_int32), ("m_tiles", C.c_int32), ("n_tiles", C.c_int32),
                ("chunks", C.c_int32), ("tiles_per_chunk", C.c_int32), ("scan_bytes", C.c_int64),
                ("scan_flops", C.c_int64), ("workspace_bytes", C.c_size_t), ("fused_tail", C.c_int32),
                ("row_bias", C.c_int32)]


class Exchange(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("nq_cap", C.c_int32), ("k_cap", C.c_int32),
                ("epoch", C.c_uint32), ("timeout_ms", C.c_int32), ("bufs_dev", C.c_void_p)]


class CamaLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_qkv", "b_qkv", "w_o", "b_o", "w_1", "b_1", "w_2", "b_2",
                                          "ln1_g", "ln1_b", "ln2_g", "ln2_b")]


# name -> (restype, argtypes); mirrors include/mrag.h one to one (tests check the list)
SIGNATURES = {
    "mrag_abi_version": (C.c_int, []),
    "mrag_last_error": (C.c_char_p, []),
    "mrag_launch_count": (C.c_int64, []),
    "mrag_store_create": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]),
    "mrag_store_destroy": (C.c_int, [C.c_void_p]),
    "mrag_store_append": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "mrag_store_set_groups": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "mrag_store_get_info": (C.c_int, [C.c_void_p, C.POINTER(StoreInfo)]),
    "mrag_store_poll_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "mrag_search_plan": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(SearchParams), C.POINTER(PlanInfo)]),
    "mrag_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(SearchParams), C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mrag_exchange_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "mrag_search_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(SearchParams), C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.POINTER(Exchange), C.c_void_p]),
    "mrag_search_timed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(SearchParams), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                    C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "mrag_search_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(SearchParams), C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mrag_search_sharded_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(SearchParams), C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Exchange), C.c_void_p]),
    "mrag_rescore_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "mrag_merge_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "mrag_gather_context": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "mrag_cama_create": (C.c_int, [C.c_int32, C.POINTER(CamaLayer), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "mrag_cama_destroy": (C.c_int, [C.c_void_p]),
    "mrag_cama_io": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "mrag_cama_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "mrag_cama_predict": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mrag_linear": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                              C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mrag_device_alloc": (C.c_int, [C.c_int32, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mrag_device_free": (C.c_int, [C.c_int32, C.c_void_p]),
    "mrag_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mrag_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mrag_ipc_close": (C.c_int, [C.c_void_p]),
}

_lib = None


def load(path: os.PathLike | str | None = None) -> C.CDLL:
    """Load libmrag.so (once). Raises if it has not been built — there is no other path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path is not None else LIB_PATH
    if not p.exists():
        raise MragError(-3, f"{p} not found: build it with `python -m motionrag_b200.build` "
                            "(the CUDA extension is mandatory, there is no CPU fallback)")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    got = lib.mrag_abi_version()
    if got != ABI_VERSION:
        raise MragError(-5, f"libmrag ABI {got} != binding ABI {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().mrag_last_error()
        raise MragError(rc, msg.decode() if msg else "")


def launch_count() -> int:
    return int(load().mrag_launch_count())
