"""oracle/oracle.py — ctypes loader for the CPU restatement (oracle/nz_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs; never by nezha_b200/.  PARITY UNPINNED: the
reference has no tests for this path and nginx cannot run here (oracle/SPEC.md).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnz_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nz_oracle.c")
    hdr = os.path.join(_HERE, "nz_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class _Range(C.Structure):
    _fields_ = [("start", C.c_int64), ("end", C.c_int64)]


class _Request(C.Structure):
    _fields_ = [
        ("method", C.c_char_p),
        ("http_minor", C.c_int),
        ("range", C.c_char_p),
        ("if_range", C.c_char_p),
        ("if_none_match", C.c_char_p),
        ("if_match", C.c_char_p),
        ("if_modified_since", C.c_char_p),
        ("if_unmodified_since", C.c_char_p),
        ("accept_encoding", C.c_char_p),
    ]


class _Object(C.Structure):
    _fields_ = [
        ("stored_length", C.c_int64),
        ("raw_length", C.c_int64),
        ("content_type", C.c_char_p),
        ("etag", C.c_char_p),
        ("last_modified", C.c_char_p),
        ("content_encoding", C.c_char_p),
        ("gunzip_enabled", C.c_int),
        ("no_ranges", C.c_int),
    ]


class Job(C.Structure):
    _fields_ = [("fd", C.c_int), ("off", C.c_int64), ("len", C.c_int64), ("dst", C.c_void_p)]


class ZJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("n", C.c_int64), ("dst", C.c_void_p), ("cap", C.c_int64), ("out", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.nzo_md5.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.nzo_crc32_ieee.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        L.nzo_crc32_ieee.restype = C.c_uint32
        for f in (L.nzo_crc32c, L.nzo_crc32c_bitwise):
            f.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
            f.restype = C.c_uint32
        L.nzo_crc32c_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.nzo_crc32c_combine.restype = C.c_uint32
        L.nzo_crc32c_chunks.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.nzo_crc32c_chunks.restype = C.c_uint32
        L.nzo_cache_key.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.nzo_cache_key.restype = C.c_int
        L.nzo_vhost_select.argtypes = [C.c_char_p]
        L.nzo_vhost_name.argtypes = [C.c_int]
        L.nzo_vhost_name.restype = C.c_char_p
        L.nzo_vhost_proxy_host.argtypes = [C.c_int]
        L.nzo_vhost_proxy_host.restype = C.c_char_p
        L.nzo_method_gate.argtypes = [C.c_char_p]
        L.nzo_range_parse.argtypes = [C.c_char_p, C.c_int64, C.POINTER(_Range), C.c_int, C.POINTER(C.c_int)]
        L.nzo_parse_http_time.argtypes = [C.c_char_p, C.c_size_t]
        L.nzo_parse_http_time.restype = C.c_int64
        L.nzo_admission.argtypes = [C.c_int, C.c_char_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        L.nzo_plan.argtypes = [C.POINTER(_Request), C.POINTER(_Object), C.c_uint64, C.c_char_p, C.c_size_t]
        L.nzo_gunzip.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.nzo_gunzip.restype = C.c_int64
        L.nzo_gzip_isize.argtypes = [C.c_void_p, C.c_size_t]
        L.nzo_gzip_isize.restype = C.c_int64
        L.nzo_serve_slice.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        L.nzo_serve_slice.restype = C.c_int64
        L.nzo_serve_slice_mem.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.nzo_serve_slice_mem.restype = C.c_int64
        L.nzo_serve_jobs.argtypes = [C.POINTER(Job), C.c_int, C.c_int]
        L.nzo_serve_jobs.restype = C.c_double
        L.nzo_gunzip_jobs.argtypes = [C.POINTER(ZJob), C.c_int, C.c_int]
        L.nzo_gunzip_jobs.restype = C.c_double
        _lib = L
    return _lib


def _buf(b) -> Tuple[C.c_void_p, int, object]:
    """bytes / bytearray / numpy uint8 array -> (pointer, length, keepalive)"""
    if isinstance(b, (bytes, bytearray)):
        arr = (C.c_char * len(b)).from_buffer_copy(b) if isinstance(b, bytes) else (C.c_char * len(b)).from_buffer(b)
        return C.cast(arr, C.c_void_p), len(b), arr
    import numpy as np

    a = np.ascontiguousarray(b).view(np.uint8).reshape(-1)
    return C.c_void_p(a.ctypes.data), a.size, a


def md5(data) -> bytes:
    p, n, _k = _buf(data)
    out = (C.c_uint8 * 16)()
    lib().nzo_md5(p, n, out)
    return bytes(out)


def crc32_ieee(data, crc: int = 0) -> int:
    p, n, _k = _buf(data)
    return lib().nzo_crc32_ieee(crc, p, n)


def crc32c(data, crc: int = 0) -> int:
    p, n, _k = _buf(data)
    return lib().nzo_crc32c(crc, p, n)


def crc32c_bitwise(data, crc: int = 0) -> int:
    p, n, _k = _buf(data)
    return lib().nzo_crc32c_bitwise(crc, p, n)


def crc32c_combine(c1: int, c2: int, len2: int) -> int:
    return lib().nzo_crc32c_combine(c1, c2, len2)


def crc32c_chunks(data, chunk: int = 65536):
    import numpy as np

    p, n, _k = _buf(data)
    nch = (n + chunk - 1) // chunk
    out = np.zeros(max(nch, 1), dtype=np.uint32)
    whole = lib().nzo_crc32c_chunks(p, n, chunk, C.c_void_p(out.ctypes.data))
    return out[:nch], whole


def crc32c_header(crc: int) -> str:
    """SPEC X3: x-goog-hash form."""
    import base64

    return base64.b64encode(crc.to_bytes(4, "big")).decode()


def cache_key(host_hdr: Optional[str], target: str, server_name: Optional[str] = None, proxy_host: Optional[str] = None):
    """SPEC K1-K4.  Returns the key bytes or None (400).  With server_name/proxy_host
    omitted the vhost is selected from the conf's table (SPEC V1)."""
    L = lib()
    if server_name is None or proxy_host is None:
        hsrc = host_hdr or ""
        tl = target.lower()
        if tl.startswith("http://") or tl.startswith("https://"):   # absolute-form: request line wins
            hsrc = target.split("://", 1)[1].split("/", 1)[0].split("?", 1)[0]
        h = hsrc.split(":")[0].rstrip(".").lower().encode()
        idx = L.nzo_vhost_select(h)
        server_name = L.nzo_vhost_name(idx).decode()
        proxy_host = L.nzo_vhost_proxy_host(idx).decode()
    out = C.create_string_buffer(16384)
    n = L.nzo_cache_key(
        host_hdr.encode("latin-1") if host_hdr is not None else None,
        target.encode("latin-1"),
        server_name.encode(),
        proxy_host.encode(),
        out,
        len(out),
    )
    return out.raw[:n] if n >= 0 else None


def key128(key: bytes) -> bytes:
    return md5(key)


def method_gate(method: str) -> bool:
    return bool(lib().nzo_method_gate(method.encode()))


def range_parse(spec: str, length: int, max_out: int = 64):
    """spec = Range value after 'bytes='.  Returns (rc, [(start,end),...])."""
    arr = (_Range * max_out)()
    n = C.c_int(0)
    rc = lib().nzo_range_parse(spec.encode("latin-1"), length, arr, max_out, C.byref(n))
    return rc, [(arr[i].start, arr[i].end) for i in range(n.value)] if rc == 0 else []


def parse_http_time(s: str) -> int:
    b = s.encode("latin-1")
    return lib().nzo_parse_http_time(b, len(b))


@dataclass
class Obj:
    stored_length: int
    content_type: str = "application/octet-stream"
    etag: Optional[str] = None
    last_modified: Optional[str] = None
    content_encoding: Optional[str] = None
    raw_length: int = -1
    gunzip_enabled: bool = False
    accept_ranges: bool = True        # SPEC R0: did the stored upstream response carry an Accept-Ranges header?


@dataclass
class Req:
    method: str = "GET"
    http_minor: int = 1
    range: Optional[str] = None
    if_range: Optional[str] = None
    if_none_match: Optional[str] = None
    if_match: Optional[str] = None
    if_modified_since: Optional[str] = None
    if_unmodified_since: Optional[str] = None
    accept_encoding: Optional[str] = None


def _e(s):
    return s.encode("latin-1") if s is not None else None


def plan(rq: Req, ob: Obj, boundary: int = 1) -> str:
    r = _Request(_e(rq.method), rq.http_minor, _e(rq.range), _e(rq.if_range), _e(rq.if_none_match), _e(rq.if_match),
                 _e(rq.if_modified_since), _e(rq.if_unmodified_since), _e(rq.accept_encoding))
    o = _Object(ob.stored_length, ob.raw_length, _e(ob.content_type or ""), _e(ob.etag), _e(ob.last_modified),
                _e(ob.content_encoding), 1 if ob.gunzip_enabled else 0, 0 if ob.accept_ranges else 1)
    out = C.create_string_buffer(1 << 16)
    n = lib().nzo_plan(C.byref(r), C.byref(o), boundary, out, len(out))
    if n < 0:
        raise RuntimeError("plan overflow")
    return out.raw[:n].decode("latin-1")


@dataclass
class Plan:
    status: int
    head: bool
    repr: str
    headers: List[Tuple[str, str]]
    body: List[Tuple[str, object]]  # ("lit", bytes) | ("slice", (start, len))
    bypass: bool = False

    def body_bytes(self, representation: bytes) -> bytes:
        out = bytearray()
        for kind, v in self.body:
            if kind == "lit":
                out += v
            else:
                s, l = v
                out += representation[s:s + l]
        return bytes(out)


def parse_plan(text: str) -> Plan:
    status, head, rep, hdrs, body, bypass = 0, False, "identity", [], [], False
    for line in text.split("\n"):
        if not line or line == "end":
            continue
        k, _, v = line.partition(" ")
        if k == "status":
            status = int(v)
        elif k == "head":
            head = v == "1"
        elif k == "repr":
            rep = v
        elif k == "bypass":
            bypass = True
        elif k == "hdr":
            n, _, val = v.partition(": ")
            hdrs.append((n, val))
        elif k == "lit":
            body.append(("lit", bytes.fromhex(v)))
        elif k == "slice":
            a, b = v.split()
            body.append(("slice", (int(a), int(b))))
    return Plan(status, head, rep, hdrs, body, bypass)


def gunzip(data) -> Optional[bytes]:
    """SPEC G2/G4: inflated bytes, or None when declined."""
    p, n, _k = _buf(data)
    isz = lib().nzo_gzip_isize(p, n)
    if isz < 0:
        return None
    # zlib verifies ISIZE, so for a valid member (< 4 GiB) isz is the exact output size
    out = C.create_string_buffer(isz + 1)
    r = lib().nzo_gunzip(p, n, out, isz + 1)
    return out.raw[:r] if r >= 0 else None


def serve_slice_mem(obj, off: int, length: int) -> bytes:
    p, n, _k = _buf(obj)
    assert 0 <= off and off + length <= n
    out = C.create_string_buffer(max(length, 1))
    r = lib().nzo_serve_slice_mem(p, off, length, out)
    assert r == length
    return out.raw[:length]


def serve_slice_fd(fd: int, off: int, length: int) -> bytes:
    out = C.create_string_buffer(max(length, 1))
    r = lib().nzo_serve_slice(fd, off, length, out)
    if r != length:
        raise IOError("short read")
    return out.raw[:length]


def admission(status: int, headers, now: int, default_valid_s: int = 86400):
    """SPEC H2: (cacheable, valid_until) for an upstream response; headers = [(name, value)] in arrival order."""
    blob = "".join(f"{k}: {v}\n" for k, v in headers).encode("latin-1")
    vu = C.c_int64()
    c = lib().nzo_admission(status, blob, now, default_valid_s, C.byref(vu))
    return bool(c), vu.value


def parse_head_text(raw: bytes) -> str:
    """SPEC P: nzo_parse_head's text form (raw = the bytes before the blank line)"""
    out = C.create_string_buffer(len(raw) + 4096)
    n = lib().nzo_parse_head(raw, len(raw), out, len(out))
    if n < 0:
        raise RuntimeError("head overflow")
    return out.raw[:n].decode("latin-1")
