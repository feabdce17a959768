"""nezha_b200/zone.py — host-side handle on one GPU's cache zone.

`Zone` is the Python spelling of what the Go proxy does over cgo: it owns an
`nz_ctx` (one per GPU = one `keys_zone` + cache directory of
/root/reference/examples/demo/nginx.conf:38), fills objects (`put`) and serves
waves of cached GETs (`submit`).  Everything that touches bytes runs in
libnezha_gpu.so's CUDA kernels; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

from . import capi
from .capi import NezhaError, check


class Buf:
    """C-owned I/O memory (nz_buf): pinned+mapped host, or HBM."""

    def __init__(self, zone: "Zone", nbytes: int, kind: int):
        self.zone, self.nbytes, self.kind = zone, nbytes, kind
        h = C.c_void_p()
        check(zone.L.nz_buf_alloc(zone.ctx, nbytes, kind, C.byref(h)), "nz_buf_alloc")
        self.h = h
        self.host = zone.L.nz_buf_ptr(h)
        self.dev = zone.L.nz_buf_devptr(h)

    def numpy(self):
        import numpy as np

        if not self.host:
            raise NezhaError(capi.NZ_EINVAL, "device buffer has no host view")
        return np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.host))

    def bytes(self, off: int = 0, n: Optional[int] = None) -> bytes:
        n = self.nbytes - off if n is None else n
        if self.host:
            return C.string_at(self.host + off, n)
        out = C.create_string_buffer(max(n, 1))
        check(self.zone.L.nz_buf_read(self.h, off, out, n), "nz_buf_read")
        return out.raw[:n]

    def write(self, off: int, data: bytes):
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        check(self.zone.L.nz_buf_write(self.h, off, buf, len(data)), "nz_buf_write")

    def free(self):
        if self.h:
            self.zone.L.nz_buf_free(self.h)
            self.h = None


@dataclass
class Get:
    key: bytes
    dst: Buf
    segs: Sequence[Tuple[int, int, int]] = ()   # (src_off, len, dst_off); () = whole object
    flags: int = 0
    dst_base: int = 0


@dataclass
class Result:
    status: int
    flags: int
    bytes: int
    length: int
    crc32c: int


class Ticket:
    def __init__(self, zone: "Zone", h, n: int, keep):
        self.zone, self.h, self.n, self._keep = zone, h, n, keep

    def wait(self, timeout_us: int = -1) -> List[Result]:
        arr = (capi.GetRes * self.n)()
        rc = self.zone.L.nz_get_wait(self.h, timeout_us, arr)
        if rc == capi.NZ_ETIMEOUT:
            raise TimeoutError("nz_get_wait")
        check(rc, "nz_get_wait")
        return [Result(r.status, r.flags, r.bytes, r.length, r.crc32c) for r in arr]

    def done(self) -> bool:
        return self.zone.L.nz_get_poll(self.h) == capi.NZ_OK

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        check(self.zone.L.nz_ticket_elapsed_ms(self.h, C.byref(ms)), "nz_ticket_elapsed_ms")
        return ms.value

    def gather_ms(self) -> float:
        ms = C.c_float()
        check(self.zone.L.nz_ticket_gather_ms(self.h, C.byref(ms)), "nz_ticket_gather_ms")
        return ms.value

    def inflate_ms(self) -> float:
        ms = C.c_float()
        check(self.zone.L.nz_ticket_inflate_ms(self.h, C.byref(ms)), "nz_ticket_inflate_ms")
        return ms.value

    def free(self):
        if self.h:
            self.zone.L.nz_ticket_free(self.h)
            self.h = None


class PreparedWave:
    """A wave whose nz_get_req array is built once and can be submitted repeatedly."""

    def __init__(self, zone: "Zone", gets: Sequence[Get]):
        self.zone, self.n = zone, len(gets)
        self.reqs = (capi.GetReq * self.n)()
        self._keep = []
        for i, g in enumerate(gets):
            r = self.reqs[i]
            C.memmove(r.key, g.key, 16)
            r.flags = g.flags
            r.nseg = len(g.segs)
            if g.segs:
                sa = (capi.Seg * len(g.segs))(*[capi.Seg(a, b, c) for a, b, c in g.segs])
                self._keep.append(sa)
                r.segs = sa
            r.dst = g.dst.h
            r.dst_base = g.dst_base

    def submit(self, stream: Optional[int] = None) -> Ticket:
        t = C.c_void_p()
        check(self.zone.L.nz_get_submit(self.zone.ctx, self.reqs, self.n, C.c_void_p(stream) if stream else None,
                                        C.byref(t)), "nz_get_submit")
        return Ticket(self.zone, t, self.n, self)


class Zone:
    def __init__(self, device: int = 0, arena_bytes: int = 1 << 30, max_objects: int = 1 << 14,
                 table_slots: int = 0, shard_index: int = 0, shard_count: int = 1):
        self.L = capi.load()
        self.device = device
        cfg = capi.Config(C.sizeof(capi.Config), device, arena_bytes, table_slots, max_objects, shard_index,
                          shard_count, 0)
        ctx = C.c_void_p()
        check(self.L.nz_init(C.byref(cfg), C.byref(ctx)), "nz_init")
        self.ctx = ctx
        self.shard_index, self.shard_count = shard_index, shard_count

    # -- lifecycle ---------------------------------------------------------
    def close(self):
        if self.ctx:
            self.L.nz_shutdown(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- memory ------------------------------------------------------------
    def numa_node(self) -> int:
        """host memory node next to this zone's GPU (-1 unknown); pinned buffers are placed there"""
        return int(self.L.nz_numa_node(self.ctx))

    def alloc_host(self, nbytes: int) -> Buf:
        return Buf(self, nbytes, capi.NZ_BUF_PINNED_HOST)

    def alloc_device(self, nbytes: int) -> Buf:
        return Buf(self, nbytes, capi.NZ_BUF_DEVICE)

    # -- fill side ----------------------------------------------------------
    def owns(self, key: bytes) -> bool:
        k = (C.c_uint8 * 16).from_buffer_copy(key)
        return self.L.nz_shard_of(k, self.shard_count) == self.shard_index

    def put(self, key: bytes, data, enc: int = capi.NZ_ENC_IDENTITY, raw_len: Optional[int] = None,
            expected_crc32c: Optional[int] = None, user_meta: int = 0) -> int:
        """Store one object.  `data`: bytes / numpy uint8 / a CUDA torch uint8 tensor.
        Returns the CRC-32C the device computed."""
        k = (C.c_uint8 * 16).from_buffer_copy(key)
        is_torch = hasattr(data, "data_ptr")
        if is_torch:
            n = data.numel()
        elif isinstance(data, (bytes, bytearray)):
            n = len(data)
        else:
            import numpy as np

            data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
            n = data.size
        p = C.c_void_p()
        check(self.L.nz_put_begin(self.ctx, k, n, raw_len if raw_len is not None else n, enc, user_meta, C.byref(p)),
              "nz_put_begin")
        try:
            if is_torch:
                if not data.is_cuda:
                    rc = self.L.nz_put_write(p, data.data_ptr(), 0, n)
                else:
                    import torch

                    torch.cuda.current_stream(data.device).synchronize()
                    rc = self.L.nz_put_write_device(p, data.data_ptr(), 0, n)
            elif isinstance(data, (bytes, bytearray)):
                buf = (C.c_char * n).from_buffer_copy(data) if n else None
                rc = self.L.nz_put_write(p, C.cast(buf, C.c_void_p) if n else None, 0, n)
            else:
                rc = self.L.nz_put_write(p, data.ctypes.data, 0, n)
            check(rc, "nz_put_write")
        except Exception:
            self.L.nz_put_abort(p)
            raise
        crc = C.c_uint32()
        check(self.L.nz_put_commit(p, expected_crc32c or 0, 1 if expected_crc32c is not None else 0, C.byref(crc)),
              "nz_put_commit")
        return crc.value

    def put_many(self, objects, enc: int = capi.NZ_ENC_IDENTITY) -> List[int]:
        """Bulk load: `objects` = [(key, data[, raw_len])], data bytes / numpy uint8.  One nz_put_commit_many for the
        lot (batched CRC / index / publish kernels).  Returns the CRC-32C of every object; raises on any failure."""
        import numpy as np

        n = len(objects)
        if n == 0:
            return []
        handles = (C.c_void_p * n)()
        keep = []
        made = 0
        try:
            for i, obj in enumerate(objects):
                key, data = obj[0], obj[1]
                raw_len = obj[2] if len(obj) > 2 else None
                if isinstance(data, (bytes, bytearray)):
                    arr = np.frombuffer(data, dtype=np.uint8)
                else:
                    arr = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
                m = arr.size
                k = (C.c_uint8 * 16).from_buffer_copy(key)
                p = C.c_void_p()
                check(self.L.nz_put_begin(self.ctx, k, m, raw_len if raw_len is not None else (0 if enc & capi.NZ_ENC_GZIP else m),
                                          enc, 0, C.byref(p)), "nz_put_begin")
                handles[i] = p
                made = i + 1
                keep.append(arr)
                check(self.L.nz_put_write(p, arr.ctypes.data if m else None, 0, m), "nz_put_write")
        except Exception:
            for i in range(made):
                self.L.nz_put_abort(handles[i])
            raise
        crcs = (C.c_uint32 * n)()
        sts = (C.c_int32 * n)()
        check(self.L.nz_put_commit_many(handles, n, None, None, crcs, sts), "nz_put_commit_many")
        for i in range(n):
            check(sts[i], f"nz_put_commit_many[{i}]")
        return list(crcs)

    def put_packed(self, objects, window: "Buf" = None) -> List[int]:
        """Bulk load of many small identity objects through ONE pinned window (nz_put_packed): `objects` = [(key, data)],
        packed back to back at arbitrary byte offsets.  Returns the CRC-32C of every object; raises on any failure."""
        import numpy as np

        n = len(objects)
        if n == 0:
            return []
        arrs = [np.frombuffer(d, dtype=np.uint8) if isinstance(d, (bytes, bytearray)) else np.ascontiguousarray(d).view(np.uint8).reshape(-1) for _, d in objects]
        total = sum(a.size for a in arrs)
        own = window is None
        if own:
            window = self.alloc_host(max(total, 1))
        try:
            view = window.numpy()
            offs, lens, pos = (C.c_uint64 * n)(), (C.c_uint64 * n)(), 0
            for i, a in enumerate(arrs):
                view[pos:pos + a.size] = a
                offs[i], lens[i] = pos, a.size
                pos += a.size
            keys = ((C.c_uint8 * 16) * n)()
            for i, (k, _) in enumerate(objects):
                keys[i][:] = k
            crcs, sts = (C.c_uint32 * n)(), (C.c_int32 * n)()
            check(self.L.nz_put_packed(self.ctx, n, keys, offs, lens, window.h, None, crcs, sts), "nz_put_packed")
            for i in range(n):
                check(sts[i], f"nz_put_packed[{i}]")
            return list(crcs)
        finally:
            if own:
                window.free()

    def put_file(self, key: bytes, fd: int, length: int, file_off: int = 0, enc: int = capi.NZ_ENC_IDENTITY,
                 raw_len: Optional[int] = None, expected_crc32c: Optional[int] = None) -> int:
        k = (C.c_uint8 * 16).from_buffer_copy(key)
        p = C.c_void_p()
        check(self.L.nz_put_begin(self.ctx, k, length, raw_len if raw_len is not None else length, enc, 0, C.byref(p)),
              "nz_put_begin")
        rc = self.L.nz_put_file(p, fd, file_off, 0, length)
        if rc != capi.NZ_OK:
            self.L.nz_put_abort(p)
            check(rc, "nz_put_file")
        crc = C.c_uint32()
        check(self.L.nz_put_commit(p, expected_crc32c or 0, 1 if expected_crc32c is not None else 0, C.byref(crc)),
              "nz_put_commit")
        return crc.value

    def evict(self, key: bytes) -> int:
        k = (C.c_uint8 * 16).from_buffer_copy(key)
        return self.L.nz_evict(self.ctx, k)

    # -- hit side -----------------------------------------------------------
    def lookup(self, keys: Sequence[bytes]) -> List[capi.Extent]:
        n = len(keys)
        if n == 0:
            return []
        blob = (C.c_uint8 * (16 * n)).from_buffer_copy(b"".join(keys))
        out = (capi.Extent * n)()
        check(self.L.nz_lookup(self.ctx, C.cast(blob, C.c_void_p), n, out), "nz_lookup")
        return list(out)

    def submit(self, gets: Sequence[Get], stream: Optional[int] = None) -> Ticket:
        """One wave of GETs.  `stream`: a cudaStream_t handle (e.g.
        torch.cuda.current_stream().cuda_stream) or None for the zone's pool."""
        return PreparedWave(self, gets).submit(stream)

    def prepare(self, gets: Sequence[Get]) -> "PreparedWave":
        """Marshal a wave once; resubmit it many times (bench loops, streaming windows)."""
        return PreparedWave(self, gets)

    def get(self, key: bytes, segs: Sequence[Tuple[int, int]] = (), flags: int = 0, device_dst: bool = False) -> Tuple[Result, bytes]:
        """Convenience for tests: one GET; slices (src_off, len) are packed back to back."""
        if segs:
            total = sum(l for _, l in segs)
        else:
            ext = self.lookup([key])[0]
            total = 0 if ext.status != capi.NZ_OK else (
                ext.raw_length if (flags & capi.NZ_GET_INFLATE and ext.flags & capi.NZ_ENC_GZIP) else ext.stored_length)
        buf = (self.alloc_device if device_dst else self.alloc_host)(max(total, 1))
        try:
            packed, off = [], 0
            for s, l in segs:
                packed.append((s, l, off))
                off += l
            t = self.submit([Get(key, buf, packed, flags)])
            try:
                res = t.wait()[0]
            finally:
                t.free()
            body = buf.bytes(0, res.bytes) if res.status == capi.NZ_OK and res.bytes else b""
            return res, body
        finally:
            buf.free()

    # -- cross-shard objects (K5) ---------------------------------------------
    def ipc_export(self) -> bytes:
        h = (C.c_uint8 * capi.NZ_IPC_HANDLE_BYTES)()
        n = C.c_uint64()
        check(self.L.nz_arena_ipc_export(self.ctx, h, C.byref(n)), "nz_arena_ipc_export")
        return bytes(h)

    def peer_attach(self, shard: int, handle: bytes):
        h = (C.c_uint8 * capi.NZ_IPC_HANDLE_BYTES).from_buffer_copy(handle)
        check(self.L.nz_peer_attach(self.ctx, shard, h), "nz_peer_attach")

    def extent_location(self, key: bytes):
        k = (C.c_uint8 * 16).from_buffer_copy(key)
        off, ln, tab, crc = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32()
        check(self.L.nz_extent_location(self.ctx, k, C.byref(off), C.byref(ln), C.byref(tab), C.byref(crc)),
              "nz_extent_location")
        return off.value, ln.value, tab.value, crc.value

    def put_striped(self, key: bytes, total_len: int, stripe_bytes: int, stripes, crc32c: int, user_meta: int = 0):
        """stripes: [(shard, arena_off, len, crc_tab_off)] in object order."""
        k = (C.c_uint8 * 16).from_buffer_copy(key)
        arr = (capi.Stripe * len(stripes))(*[capi.Stripe(s, 0, o, l, t) for s, o, l, t in stripes])
        check(self.L.nz_put_striped(self.ctx, k, total_len, stripe_bytes, len(stripes), arr, crc32c, user_meta),
              "nz_put_striped")

    def crc32c_combine(self, crc_a: int, crc_b: int, len_b: int) -> int:
        return self.L.nz_crc32c_combine(crc_a, crc_b, len_b)

    def debug_poke(self, key: bytes, off: int, xor_mask: int = 0x01) -> int:
        k = (C.c_uint8 * 16).from_buffer_copy(key)
        return self.L.nz_debug_poke(self.ctx, k, off, xor_mask)

    def set_profiling(self, on: bool = True):
        check(self.L.nz_set_profiling(self.ctx, 1 if on else 0), "nz_set_profiling")

    def set_host_rate(self, total_gbs: float, ctas: int = 0) -> None:
        check(self.L.nz_set_host_rate(self.ctx, float(total_gbs), int(ctas)), "nz_set_host_rate")

    def stats(self) -> capi.Stats:
        s = capi.Stats()
        check(self.L.nz_get_stats(self.ctx, C.byref(s)), "nz_get_stats")
        return s

    def crc32c_device(self, devptr: int, n: int, want_chunks: bool = False):
        import numpy as np

        crc = C.c_uint32()
        nch = (n + capi.NZ_CRC_CHUNK - 1) // capi.NZ_CRC_CHUNK
        chunks = np.zeros(max(nch, 1), dtype=np.uint32) if want_chunks else None
        check(self.L.nz_crc32c_device(self.ctx, devptr, n, C.byref(crc), chunks.ctypes.data if want_chunks else None),
              "nz_crc32c_device")
        return (crc.value, chunks[:nch]) if want_chunks else crc.value
