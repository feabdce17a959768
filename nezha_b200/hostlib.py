"""nezha_b200/hostlib.py — ctypes view of libnezha_host.so (nezha_b200/host/nz_http.cc),
the product's host-side HTTP semantics (key, method gate, Range, conditionals, plan)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libnezha_host.so")


class Request(C.Structure):
    _fields_ = [("method", C.c_char_p), ("http_minor", C.c_int), ("range", C.c_char_p), ("if_range", C.c_char_p),
                ("if_none_match", C.c_char_p), ("if_match", C.c_char_p), ("if_modified_since", C.c_char_p),
                ("if_unmodified_since", C.c_char_p), ("accept_encoding", C.c_char_p)]


class Object(C.Structure):
    _fields_ = [("stored_length", C.c_int64), ("raw_length", C.c_int64), ("content_type", C.c_char_p),
                ("etag", C.c_char_p), ("last_modified", C.c_char_p), ("content_encoding", C.c_char_p),
                ("gunzip_enabled", C.c_int), ("no_ranges", C.c_int)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        L.nzh_cache_key.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.nzh_method_allowed.argtypes = [C.c_char_p]
        L.nzh_parse_ranges.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int)]
        L.nzh_parse_http_date.argtypes = [C.c_char_p, C.c_size_t]
        L.nzh_parse_http_date.restype = C.c_int64
        L.nzh_md5.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.nzh_plan_text.argtypes = [C.POINTER(Request), C.POINTER(Object), C.c_uint64, C.c_char_p, C.c_size_t]
        L.nzh_admission.argtypes = [C.c_int, C.c_char_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _e(s):
    return s.encode("latin-1") if s is not None else None


def cache_key(host: Optional[str], target: str) -> Optional[bytes]:
    out = C.create_string_buffer(16384)
    n = load().nzh_cache_key(_e(host), _e(target), out, len(out))
    return out.raw[:n] if n >= 0 else None


def method_allowed(m: str) -> bool:
    return bool(load().nzh_method_allowed(_e(m)))


def parse_ranges(spec: str, length: int, max_pairs: int = 64):
    arr = (C.c_int64 * (2 * max_pairs))()
    n = C.c_int(0)
    rc = load().nzh_parse_ranges(_e(spec), length, arr, max_pairs, C.byref(n))
    return rc, [(arr[2 * i], arr[2 * i + 1]) for i in range(n.value)]


def parse_http_date(s: str) -> int:
    b = _e(s)
    return load().nzh_parse_http_date(b, len(b))


def md5(data: bytes) -> bytes:
    out = (C.c_uint8 * 16)()
    load().nzh_md5(data, len(data), out)
    return bytes(out)


def plan_text(method="GET", http_minor=1, range=None, if_range=None, if_none_match=None, if_match=None,
              if_modified_since=None, if_unmodified_since=None, accept_encoding=None, *, stored_length=0, raw_length=-1,
              content_type="", etag=None, last_modified=None, content_encoding=None, gunzip_enabled=False,
              boundary=1, accept_ranges=True) -> str:
    r = Request(_e(method), http_minor, _e(range), _e(if_range), _e(if_none_match), _e(if_match), _e(if_modified_since),
                _e(if_unmodified_since), _e(accept_encoding))
    o = Object(stored_length, raw_length, _e(content_type or ""), _e(etag), _e(last_modified), _e(content_encoding),
               1 if gunzip_enabled else 0, 0 if accept_ranges else 1)
    out = C.create_string_buffer(1 << 16)
    n = load().nzh_plan_text(C.byref(r), C.byref(o), boundary, out, len(out))
    if n < 0:
        raise RuntimeError("plan overflow")
    return out.raw[:n].decode("latin-1")


def admission(status: int, headers, now: int, default_valid_s: int = 86400):
    """SPEC H2 (nginx.conf:57): (cacheable, valid_until) of an upstream response; headers in arrival order."""
    blob = "".join(f"{k}: {v}\n" for k, v in headers).encode("latin-1")
    vu = C.c_int64()
    c = load().nzh_admission(status, blob, now, default_valid_s, C.byref(vu))
    return bool(c), vu.value


def parse_head_text(raw: bytes) -> str:
    """SPEC P: nzh::parse_head(raw).to_text() (raw = the bytes before the blank line)"""
    out = C.create_string_buffer(len(raw) + 4096)
    n = load().nzh_parse_head_text(raw, len(raw), out, len(out))
    if n < 0:
        raise RuntimeError("head overflow")
    return out.raw[:n].decode("latin-1")
