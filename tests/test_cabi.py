"""The C-ABI library builds without a GPU, loads, and exports exactly what include/mrag.h declares."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "mrag.h").read_text()
    return sorted(set(re.findall(r"MRAG_API\s+[\w\s\*]+?\b(mrag_\w+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = _declared()
    for must in ("mrag_store_create", "mrag_store_append", "mrag_search", "mrag_search_host",
                 "mrag_merge_topk", "mrag_gather_context", "mrag_search_plan", "mrag_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(libmrag):
    from motionrag_b200 import _cabi
    out = subprocess.run(["nm", "-D", "--defined-only", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(mrag_\w+)", out))
    declared = set(_declared())
    assert declared <= exported, declared - exported
    assert exported <= declared, f"undeclared exports: {exported - declared}"
    assert set(_cabi.SIGNATURES) == declared           # the ctypes table mirrors the header
    for name in declared:
        assert getattr(libmrag, name) is not None
    assert libmrag.mrag_abi_version() == _cabi.ABI_VERSION


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mrag.h"\nint main(void){ mrag_search_params p; p.k = 1; return p.k - 1 + (MRAG_OK); }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src),
                        "-o", str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every struct as gcc sees include/mrag.h == the ctypes mirror."""
    from motionrag_b200 import _cabi
    structs = {"mrag_search_params": _cabi.SearchParams, "mrag_store_info": _cabi.StoreInfo,
               "mrag_plan_info": _cabi.PlanInfo, "mrag_exchange": _cabi.Exchange, "mrag_cama_layer": _cabi.CamaLayer}
    lines = []
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mrag.h"\nint main(void){\n' + "\n".join(lines) + "\nreturn 0;}\n")
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c99", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), (cname, got[cname], C.sizeof(ct))
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)


def test_library_contains_blackwell_sass(libmrag):
    """tcgen05 / TMEM / TMA must be in the shipped cubin (not a recompiled legacy path)."""
    from motionrag_b200 import _cabi
    sass = subprocess.run(["cuobjdump", "-sass", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "HGMMA" not in sass
    # warp-level mma.sync is allowed in one place only: the 25-row block-causal attention tiles of K6
    # (far below a tcgen05 tile); the scan and every GEMM must be on tcgen05
    fn = None
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif "HMMA.16816" in line:
            assert fn is not None and "k6_attention" in fn, fn
    elf = subprocess.run(["cuobjdump", "-lelf", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in elf


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_calls_fail_loudly_without_gpu(libmrag):
    from motionrag_b200 import EmbeddingStore, MragError, RAGDatabase
    h = C.c_void_p()
    rc = libmrag.mrag_store_create(768, 1000, 0, C.byref(h))
    assert rc == -3 and b"no CPU path" in libmrag.mrag_last_error()
    with pytest.raises(MragError):
        EmbeddingStore(768, 1000, 0)
    with pytest.raises(MragError):
        RAGDatabase(None, None, columns={"text_embedding": [[0.0] * 768], "video": ["a"]})


def test_missing_library_is_an_error(tmp_path):
    from motionrag_b200 import MragError, _cabi
    with pytest.raises(MragError):
        _cabi.load(tmp_path / "nope.so")


def test_product_package_never_imports_the_oracle():
    for f in (ROOT / "motionrag_b200").rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(), re.M), f


def _build_c_host(tmp_path):
    from motionrag_b200 import _cabi
    exe = tmp_path / "search_host"
    r = subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", str(ROOT / "tests" / "c_host" / "search_host.c"),
                        "-I", str(ROOT / "include"), "-L", str(_cabi.LIB_PATH.parent), "-lmrag", "-lm",
                        f"-Wl,-rpath,{_cabi.LIB_PATH.parent}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_pure_c_host_links_and_fails_loudly_without_gpu(libmrag, tmp_path):
    """include/mrag.h + libmrag.so are enough for a host in another language: a C99 program links
    against them; with no GPU the library refuses with MRAG_ERR_DEVICE instead of computing on the CPU."""
    r = subprocess.run([str(_build_c_host(tmp_path))], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU path" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_pure_c_host_search_matches_brute_force(libmrag, tmp_path):
    r = subprocess.run([str(_build_c_host(tmp_path))], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "0 mismatches" in r.stdout
