"""nezha_b200/workloads.py — seeded synthetic inputs for BASELINE.json's configs.

SURVEY.md §8(d) fixes these (seeds included) so that tests, bench.py and the CPU
baseline all see the same objects and the same request log.  Nothing here
computes results; it only makes inputs.
"""
from __future__ import annotations

import hashlib
import math
import zlib
from typing import List, Optional, Tuple

import numpy as np

C1_LENGTH = 169_001_437            # /root/reference/README.md:94,116
C1_KEY = b"www.cs.toronto.edu/www.cs.toronto.edu/~kriz/cifar-100-python.tar.gz"   # nginx.conf:94 + kubeflow-test.yaml:22
C2_LENGTH = 1 << 30
C2_REQUESTS = 64
C3_OBJECTS = 10_000
C3_RAW = 1 << 20


def key128(key_string: bytes) -> bytes:
    """Cache identity = MD5("$host/$proxy_host$uri") (nginx.conf:38,56)."""
    return hashlib.md5(key_string).digest()


def object_key_string(config: str, index: int, host: str = "storage.googleapis.com") -> bytes:
    return f"{host}/{host}/nezha-bench/{config}/obj-{index:08d}.bin".encode()


def synth_bytes(n: int, seed: int) -> np.ndarray:
    """Incompressible bytes (like a .tar.gz), Philox counter-based so any size is cheap."""
    out = np.empty(n, dtype=np.uint8)
    g = np.random.Generator(np.random.Philox(key=seed))
    step = 1 << 26
    for off in range(0, n, step):
        m = min(step, n - off)
        out[off:off + m] = g.integers(0, 256, m, dtype=np.uint8)
    return out


def c2_requests(length: int = C2_LENGTH, n: int = C2_REQUESTS, seed: int = 22) -> List[Optional[Tuple[int, int]]]:
    """C2: 50 % full GETs (None), 50 % single ranges (start, len): start ~ U[0,L),
    len ~ log-uniform[4 KiB, 256 MiB] clamped to the object end."""
    g = np.random.Generator(np.random.Philox(key=seed))
    kinds = np.arange(n) % 2
    g.shuffle(kinds)
    out: List[Optional[Tuple[int, int]]] = []
    lo, hi = math.log(4 << 10), math.log(256 << 20)
    for k in kinds:
        if k == 0:
            out.append(None)
        else:
            start = int(g.integers(0, length))
            ln = int(math.exp(g.uniform(lo, hi)))
            ln = max(1, min(ln, length - start))
            out.append((start, ln))
    return out


_VOCAB_CACHE = {}


def _vocab():
    if "tab" not in _VOCAB_CACHE:
        vg = np.random.Generator(np.random.Philox(key=1234))
        letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
        nw, maxw = 4096, 12
        tab = np.zeros((nw, maxw), dtype=np.uint8)
        lens = vg.integers(2, 11, nw)
        for i in range(nw):
            tab[i, :lens[i]] = letters[np.minimum(vg.zipf(1.3, lens[i]) - 1, 25)]
        # separator appended to each token: ' ' (90 %), ', ' (5 %), '\n' (5 %) chosen per token later
        _VOCAB_CACHE["tab"], _VOCAB_CACHE["lens"] = tab, lens.astype(np.int64)
    return _VOCAB_CACHE["tab"], _VOCAB_CACHE["lens"]


def text_like(n: int, seed: int) -> np.ndarray:
    """Seeded 'text-like' bytes: Zipf-distributed words from a fixed 4096-word vocabulary,
    which zlib level 6 (gzip_comp_level 6, nginx.conf:23) packs about 3:1.  Vectorised."""
    tab, wlen = _vocab()
    g = np.random.Generator(np.random.Philox(key=seed))
    cnt = n // 4 + 16                                     # >= tokens needed (min token is 3 bytes)
    idx = np.minimum(g.zipf(1.15, cnt) - 1, tab.shape[0] - 1)
    sepk = g.integers(0, 20, cnt)
    seplen = np.where(sepk == 1, 2, 1)
    tl = wlen[idx] + seplen
    start = np.concatenate(([0], np.cumsum(tl)[:-1]))
    keep = start < n
    idx, sepk, seplen, tl, start = idx[keep], sepk[keep], seplen[keep], tl[keep], start[keep]
    out = np.full(int(start[-1] + tl[-1]) + 2, 0x20, dtype=np.uint8)
    wl_ = wlen[idx]
    for j in range(tab.shape[1]):
        m = wl_ > j
        out[start[m] + j] = tab[idx[m], j]
    send = start + wl_
    out[send[sepk == 0]] = 0x0A
    m1 = sepk == 1
    out[send[m1]] = 0x2C                                  # ", "
    return out[:n].copy()


def gzip_member(raw, level: int = 6) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, 31)
    return c.compress(bytes(raw)) + c.flush()


def zipf_ranks(draws: int, items: int, s: float = 0.99, seed: int = 3) -> np.ndarray:
    """Bounded Zipf(s) over ranks 0..items-1 by inverse CDF."""
    w = 1.0 / np.power(np.arange(1, items + 1, dtype=np.float64), s)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    g = np.random.Generator(np.random.Philox(key=seed))
    return np.searchsorted(cdf, g.random(draws), side="left").astype(np.int64)
