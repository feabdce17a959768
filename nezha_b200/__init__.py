"""nezha_b200 — B200-native cache-hit data path of fast-ml/nezha's dataset cache.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI of
include/nezha_gpu.h), `host/` (C++ HTTP front end above that ABI) and this thin
Python harness over the same ABI.  Importing the package does not touch CUDA;
using it without the built library or without a GPU raises NezhaError.
"""
from . import capi
from .capi import NezhaError
from .zone import Buf, Get, PreparedWave, Result, Ticket, Zone

__all__ = ["capi", "NezhaError", "Zone", "Buf", "Get", "Result", "Ticket", "PreparedWave"]
