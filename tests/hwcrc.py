"""The x86 `crc32` instruction (SSE4.2) as an independent witness of CRC-32C: Intel's implementation of the Castagnoli
polynomial, not a line of it written here.  Compiled on first use into the pytest temp dir; None where gcc or the
instruction is missing (the tests that use it skip)."""
import ctypes as C
import os
import subprocess
import tempfile

SRC = r"""
#include <stddef.h>
#include <stdint.h>
#include <string.h>
uint32_t hw_crc32c(uint32_t crc, const unsigned char* p, size_t n) {
    uint64_t c = (uint32_t)~crc;
    while (n >= 8) { uint64_t v; memcpy(&v, p, 8); c = __builtin_ia32_crc32di(c, v); p += 8; n -= 8; }
    uint32_t c32 = (uint32_t)c;
    while (n--) c32 = __builtin_ia32_crc32qi(c32, *p++);
    return ~c32;
}
"""
_lib = None
_tried = False


def lib():
    global _lib, _tried
    if _tried:
        return _lib
    _tried = True
    try:
        if "sse4_2" not in open("/proc/cpuinfo").read():
            return None
        d = tempfile.mkdtemp(prefix="nz_hwcrc_")
        src, so = os.path.join(d, "hwcrc.c"), os.path.join(d, "hwcrc.so")
        with open(src, "w") as f:
            f.write(SRC)
        subprocess.run(["gcc", "-O2", "-msse4.2", "-shared", "-fPIC", "-o", so, src], check=True, capture_output=True)
        L = C.CDLL(so)
        L.hw_crc32c.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        L.hw_crc32c.restype = C.c_uint32
        _lib = L
    except Exception:
        _lib = None
    return _lib


def crc32c(data, crc: int = 0) -> int:
    import numpy as np

    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    return int(lib().hw_crc32c(crc, C.c_void_p(a.ctypes.data) if a.size else None, a.size))
