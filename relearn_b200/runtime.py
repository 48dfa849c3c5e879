"""Context and device-memory helpers over the C ABI (numpy in, numpy out)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class Context:
    """One per process/GPU: owns the stream and the optional NCCL data-parallel group."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = L.lib()
        h = C.c_void_p()
        L.check(self._lib.rl_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.device = device
        self.rank, self.world_size = 0, 1

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if self.handle:
            self._lib.rl_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        L.check(self._lib.rl_ctx_synchronize(self.handle), self.handle)

    @property
    def launch_count(self) -> int:
        return int(self._lib.rl_ctx_launch_count(self.handle))

    def device_info(self) -> dict:
        sm, maj, mnr, mem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
        L.check(self._lib.rl_ctx_device_info(self.handle, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)),
                self.handle)
        return {"sm_count": sm.value, "cc": (maj.value, mnr.value), "total_mem": mem.value}

    # -- timing -------------------------------------------------------------------------------
    def event(self) -> "Event":
        return Event(self)

    def fp32_peak_tflops(self) -> float:
        v = C.c_double()
        L.check(self._lib.rl_probe_fp32_tflops(self.handle, C.byref(v)), self.handle)
        return v.value

    # -- data-parallel group ----------------------------------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, world_size: int):
        assert len(unique_id) == L.RL_NCCL_UNIQUE_ID_BYTES
        buf = C.create_string_buffer(unique_id, L.RL_NCCL_UNIQUE_ID_BYTES)
        L.check(self._lib.rl_ctx_comm_init(self.handle, buf, rank, world_size), self.handle)
        self.rank, self.world_size = rank, world_size

    def comm_peer_info(self) -> dict:
        """Whether the update's reductions use the NVLink peer mailboxes (one fused kernel per pass) and whether a
        wait for a peer ever timed out.  Synchronises the stream."""
        on, bad = C.c_int32(), C.c_int32()
        L.check(self._lib.rl_ctx_comm_peer_info(self.handle, C.byref(on), C.byref(bad)), self.handle)
        return {"peer_mailboxes": bool(on.value), "timed_out": bool(bad.value)}

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(L.RL_NCCL_UNIQUE_ID_BYTES)
        L.check(L.lib().rl_nccl_unique_id(buf))
        return buf.raw

    # -- memory -----------------------------------------------------------------------------
    def alloc(self, nbytes: int) -> "DeviceBuffer":
        return DeviceBuffer(self, nbytes)

    def pinned_array(self, shape, dtype) -> np.ndarray:
        """numpy array over page-locked host memory (kept alive by the context)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        L.check(self._lib.rl_malloc_host(self.handle, n, C.byref(p)), self.handle)
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned = getattr(self, "_pinned", []) + [p]
        return arr

    def to_device(self, array: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(array)
        buf = DeviceBuffer(self, a.nbytes)
        buf.upload(a)
        return buf

    def read(self, ptr, shape, dtype) -> np.ndarray:
        """Copy `shape` elements of `dtype` from a raw device pointer to a new numpy array."""
        out = np.empty(shape, dtype=dtype)
        if out.nbytes:
            L.check(self._lib.rl_memcpy_d2h(self.handle, out.ctypes.data_as(C.c_void_p), C.c_void_p(_addr(ptr)),
                                            out.nbytes), self.handle)
        return out


def _addr(ptr) -> int:
    if isinstance(ptr, DeviceBuffer):
        return ptr.ptr
    if isinstance(ptr, C.c_void_p):
        return ptr.value or 0
    return int(ptr or 0)


class DeviceBuffer:
    def __init__(self, ctx: Context, nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        L.check(ctx._lib.rl_malloc(ctx.handle, self.nbytes, C.byref(p)), ctx.handle)
        self.ptr = p.value

    def upload(self, array: np.ndarray):
        a = np.ascontiguousarray(array)
        assert a.nbytes <= self.nbytes
        if a.nbytes:
            L.check(self.ctx._lib.rl_memcpy_h2d(self.ctx.handle, C.c_void_p(self.ptr), a.ctypes.data_as(C.c_void_p),
                                                a.nbytes), self.ctx.handle)

    def download(self, shape, dtype) -> np.ndarray:
        return self.ctx.read(self.ptr, shape, dtype)

    def zero(self):
        L.check(self.ctx._lib.rl_memset(self.ctx.handle, C.c_void_p(self.ptr), 0, self.nbytes), self.ctx.handle)

    def free(self):
        if self.ptr and self.ctx.handle:
            self.ctx._lib.rl_free(self.ctx.handle, C.c_void_p(self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def c(self) -> C.c_void_p:
        return C.c_void_p(self.ptr)


class Event:
    """cudaEvent on the context stream."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = C.c_void_p()
        L.check(ctx._lib.rl_event_create(ctx.handle, C.byref(h)), ctx.handle)
        self.handle = h

    def record(self):
        L.check(self.ctx._lib.rl_event_record(self.handle), self.ctx.handle)
        return self

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        L.check(self.ctx._lib.rl_event_elapsed_ms(self.handle, stop.handle, C.byref(ms)), self.ctx.handle)
        return ms.value

    def __del__(self):
        try:
            if self.handle and self.ctx.handle:
                self.ctx._lib.rl_event_destroy(self.handle)
        except Exception:
            pass
