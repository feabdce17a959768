"""nezha_b200/capi.py — ctypes view of include/nezha_gpu.h (libnezha_gpu.so).

This is the same C ABI the Go proxy binds with cgo; Python is
only a test/bench harness over it.  There is no CPU fallback: if the shared
library is missing, or CUDA is, the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libnezha_gpu.so")

NZ_OK, NZ_EINVAL, NZ_ENOMEM, NZ_ECUDA, NZ_ENOTFOUND = 0, -1, -2, -3, -4
NZ_ECORRUPT, NZ_ETIMEOUT, NZ_EEXIST, NZ_EFULL, NZ_ENOSYS, NZ_ERANGE = -5, -6, -7, -8, -9, -10
NZ_ENC_IDENTITY, NZ_ENC_GZIP, NZ_ENC_STRIPE, NZ_ENC_STRIPED, NZ_ENC_INFLATABLE = 0, 1, 4, 8, 16
NZ_IPC_HANDLE_BYTES = 64
NZ_GET_INFLATE, NZ_GET_VERIFY, NZ_GET_URGENT = 1, 2, 4
NZ_BUF_PINNED_HOST, NZ_BUF_DEVICE = 0, 1
NZ_CRC_CHUNK = 65536

# every symbol include/nezha_gpu.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "nz_abi_version", "nz_strerror", "nz_last_error", "nz_init", "nz_shutdown",
    "nz_buf_alloc", "nz_buf_wrap_device", "nz_buf_ptr", "nz_buf_devptr", "nz_buf_size", "nz_buf_free", "nz_buf_read", "nz_buf_write",
    "nz_put_begin", "nz_put_write", "nz_put_sync", "nz_put_write_device", "nz_put_file", "nz_cufile_status", "nz_numa_node", "nz_put_commit", "nz_put_commit_many", "nz_put_packed", "nz_put_abort",
    "nz_evict", "nz_lookup", "nz_get_submit", "nz_get_wait", "nz_get_poll", "nz_ticket_elapsed_ms",
    "nz_ticket_gather_ms", "nz_ticket_inflate_ms", "nz_ticket_free", "nz_get_stats", "nz_set_profiling", "nz_set_host_rate", "nz_crc32c_device", "nz_debug_poke", "nz_shard_of",
    "nz_arena_ipc_export", "nz_peer_attach", "nz_extent_location", "nz_put_striped", "nz_crc32c_combine",
]


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("arena_bytes", C.c_uint64),
        ("table_slots", C.c_uint64), ("max_objects", C.c_uint32), ("shard_index", C.c_uint32),
        ("shard_count", C.c_uint32), ("flags", C.c_uint32),
    ]


class Extent(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("flags", C.c_uint32), ("stored_length", C.c_uint64), ("raw_length", C.c_uint64),
        ("crc32c", C.c_uint32), ("desc_index", C.c_uint32), ("user_meta", C.c_uint64),
    ]


class Seg(C.Structure):
    _fields_ = [("src_off", C.c_uint64), ("len", C.c_uint64), ("dst_off", C.c_uint64)]


class GetReq(C.Structure):
    _fields_ = [
        ("key", C.c_uint8 * 16), ("flags", C.c_uint32), ("nseg", C.c_uint32), ("segs", C.POINTER(Seg)),
        ("dst", C.c_void_p), ("dst_base", C.c_uint64),
    ]


class GetRes(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("flags", C.c_uint32), ("bytes", C.c_uint64), ("length", C.c_uint64),
        ("crc32c", C.c_uint32), ("reserved", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "objects", "arena_bytes", "arena_used", "table_slots", "table_live", "gets", "get_bytes", "hits",
        "misses", "puts", "put_bytes", "evictions", "corrupt", "kernel_launches", "file_cufile_bytes", "file_bounce_bytes",
        "verify_chunks", "verify_fused_chunks", "host_dma_waves", "inflate_segments", "inflate_whole_members",
        "index_bytes")]


class Stripe(C.Structure):
    _fields_ = [("shard", C.c_uint32), ("reserved", C.c_uint32), ("arena_off", C.c_uint64), ("len", C.c_uint64),
                ("crc_tab_off", C.c_uint64)]


class NezhaError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(what)


_lib = None


def load() -> C.CDLL:
    """dlopen libnezha_gpu.so.  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NezhaError(NZ_ENOSYS, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(or make -C nezha_b200/csrc); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u8p = C.c_void_p, C.POINTER(C.c_uint8)
    L.nz_abi_version.restype = C.c_int32
    L.nz_strerror.argtypes = [C.c_int32]
    L.nz_strerror.restype = C.c_char_p
    L.nz_last_error.restype = C.c_char_p
    L.nz_numa_node.restype = C.c_int32
    L.nz_numa_node.argtypes = [C.c_void_p]
    L.nz_cufile_status.restype = C.c_char_p
    L.nz_cufile_status.argtypes = []
    L.nz_init.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.nz_shutdown.argtypes = [vp]
    L.nz_buf_alloc.argtypes = [vp, C.c_uint64, C.c_int32, C.POINTER(vp)]
    L.nz_buf_wrap_device.argtypes = [vp, vp, C.c_uint64, C.POINTER(vp)]
    L.nz_buf_ptr.argtypes = [vp]
    L.nz_buf_ptr.restype = vp
    L.nz_buf_devptr.argtypes = [vp]
    L.nz_buf_devptr.restype = vp
    L.nz_buf_size.argtypes = [vp]
    L.nz_buf_size.restype = C.c_uint64
    L.nz_buf_free.argtypes = [vp]
    L.nz_buf_read.argtypes = [vp, C.c_uint64, vp, C.c_uint64]
    L.nz_buf_write.argtypes = [vp, C.c_uint64, vp, C.c_uint64]
    L.nz_put_begin.argtypes = [vp, u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(vp)]
    L.nz_put_write.argtypes = [vp, vp, C.c_uint64, C.c_uint64]
    L.nz_put_sync.argtypes = [vp]
    L.nz_put_write_device.argtypes = [vp, vp, C.c_uint64, C.c_uint64]
    L.nz_put_file.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]
    L.nz_put_commit.argtypes = [vp, C.c_uint32, C.c_int32, C.POINTER(C.c_uint32)]
    L.nz_put_packed.argtypes = [vp, C.c_uint32, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
    L.nz_put_commit_many.argtypes = [C.POINTER(vp), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
    L.nz_put_abort.argtypes = [vp]
    L.nz_evict.argtypes = [vp, u8p]
    L.nz_lookup.argtypes = [vp, vp, C.c_uint32, C.POINTER(Extent)]
    L.nz_get_submit.argtypes = [vp, C.POINTER(GetReq), C.c_uint32, vp, C.POINTER(vp)]
    L.nz_get_wait.argtypes = [vp, C.c_int64, C.POINTER(GetRes)]
    L.nz_get_poll.argtypes = [vp]
    L.nz_ticket_elapsed_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.nz_ticket_gather_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.nz_ticket_inflate_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.nz_ticket_free.argtypes = [vp]
    L.nz_debug_poke.argtypes = [vp, u8p, C.c_uint64, C.c_uint8]
    L.nz_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.nz_set_profiling.argtypes = [vp, C.c_int32]
    L.nz_set_host_rate.argtypes = [vp, C.c_double, C.c_int32]
    L.nz_crc32c_device.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint32), vp]
    L.nz_arena_ipc_export.argtypes = [vp, u8p, C.POINTER(C.c_uint64)]
    L.nz_peer_attach.argtypes = [vp, C.c_uint32, u8p]
    L.nz_extent_location.argtypes = [vp, u8p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.nz_put_striped.argtypes = [vp, u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(Stripe), C.c_uint32, C.c_uint64]
    L.nz_crc32c_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    L.nz_crc32c_combine.restype = C.c_uint32
    L.nz_shard_of.argtypes = [u8p, C.c_uint32]
    L.nz_shard_of.restype = C.c_uint32
    for name in SYMBOLS:
        f = getattr(L, name)
        if f.restype is C.c_int and name not in ("nz_abi_version",):
            f.restype = C.c_int32
    _lib = L
    return L


def check(rc: int, what: str = "") -> int:
    if rc != NZ_OK:
        L = load()
        msg = L.nz_strerror(rc).decode()
        detail = L.nz_last_error().decode() if rc == NZ_ECUDA else ""
        raise NezhaError(rc, f"{what}: {msg}" + (f" [{detail}]" if detail else ""))
    return rc
