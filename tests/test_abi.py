"""-m "not gpu": libnezha_gpu.so loads and exports every symbol include/nezha_gpu.h
declares; without a GPU the product fails loudly instead of falling back."""
import os
import re

import pytest

import nezha_b200 as nz
from nezha_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "nezha_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nz_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    L = capi.load()
    declared = _declared()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == declared, (set(declared) ^ set(capi.SYMBOLS))
    assert L.nz_abi_version() == 2


def test_strerror_and_shard_of():
    L = capi.load()
    assert L.nz_strerror(capi.NZ_ENOTFOUND) == b"key not found (miss)"
    import ctypes as C
    key = bytes(range(16))
    k = (C.c_uint8 * 16).from_buffer_copy(key)
    want = int.from_bytes(key[:8], "little") % 8
    assert L.nz_shard_of(k, 8) == want and L.nz_shard_of(k, 1) == 0 and L.nz_shard_of(k, 0) == 0


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(capi.Config) == 40 and C.sizeof(capi.Extent) == 40 and C.sizeof(capi.Seg) == 24
    assert C.sizeof(capi.GetReq) == 48 and C.sizeof(capi.GetRes) == 32 and C.sizeof(capi.Stats) == 22 * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nz.NezhaError) as ei:
        nz.Zone(device=0, arena_bytes=1 << 20)
    assert ei.value.code == capi.NZ_ECUDA


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under nezha_b200/ may reference it."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "nezha_b200")):
        if "_build" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".cpp", ".h", ".hpp", ".go", "Makefile")):
                t = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"nz_oracle|from oracle|import oracle|libnz_oracle|nzo_", t):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_crc32c_combine_matches_oracle():
    import numpy as np
    from oracle import oracle as o
    L = capi.load()
    rng = np.random.default_rng(5)
    for n1, n2 in [(1, 1), (100, 65536), (65536, 1), (12345, 678901)]:
        a = rng.integers(0, 256, n1, dtype=np.uint8)
        b = rng.integers(0, 256, n2, dtype=np.uint8)
        assert L.nz_crc32c_combine(o.crc32c(a), o.crc32c(b), n2) == o.crc32c(np.concatenate([a, b]))
    assert L.nz_crc32c_combine(0x12345678, 0, 0) == 0x12345678
    assert L.nz_crc32c_combine(o.crc32c(b"x"), o.crc32c(bytes(1 << 20)), (1 << 33) + 5) == o.crc32c_combine(
        o.crc32c(b"x"), o.crc32c(bytes(1 << 20)), (1 << 33) + 5)


