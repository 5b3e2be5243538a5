"""ctypes binding of libnsb.so (include/nsb.h) — plumbing only, no compute and NO CPU fallback.

The product is the C-ABI CUDA library; this module lets the tests, ``bench.py`` and
``__graft_entry__.smoke()`` drive it with numpy host buffers and raw device pointers.  If the
library is missing or no sm_100a device is present every entry point raises ``NsbError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnsb.so")

NUM_SQUARES = 81
POLICY_SIZE = 2187
FEATURE_CHANNELS = 86
MAX_LEGAL_MOVES = 593
CACHE_MAX_MOVES = 164
DECODE_PROBS = 0
DECODE_LOGITS = 1
DECODE_BOTH = 2            # self-play: raw logits to the cache, probabilities out (frame.cc:93-118)
DECODE_NAN_FALLBACK = 0x100  # feedResult<NaNFallbackEnabled = true>; off by default like src/context.h:103
ROW_SKIP_SOFTMAX = 1       # row flag: Gumbel root (frame.cc:116-118)

# 16-byte packed plane == nshogi ml::FeatureBitboard (reference src/cuda/extractbit.cu:20-37)
FEATURE_BITBOARD = np.dtype([("lo", "<u8"), ("hi", "<u8")])
# nsb_position (include/nsb.h), 108 bytes
POSITION = np.dtype(
    {
        "names": ["board", "side", "hands", "ply", "max_ply", "black_draw_value", "white_draw_value"],
        "formats": [("u1", (81,)), "u1", ("u1", (2, 7)), "<u2", "<u2", "<f4", "<f4"],
        "offsets": [0, 81, 82, 96, 98, 100, 104],
        "itemsize": 108,
    }
)


class NetDesc(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("channels", C.c_int32), ("blocks", C.c_int32),
                ("value_hidden", C.c_int32)]


class NsbError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/nsb.h declares
_P = C.c_void_p
SIGNATURES = {
    "nsb_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.POINTER(NetDesc)]),
    "nsb_destroy": (None, [_P]),
    "nsb_bind_thread": (C.c_int, [_P]),
    "nsb_weight_blob_floats": (C.c_size_t, [C.POINTER(NetDesc)]),
    "nsb_weight_blob_random": (C.c_int, [C.POINTER(NetDesc), C.c_uint64, _P]),
    "nsb_load_weights": (C.c_int, [_P, _P, C.c_size_t]),
    "nsb_eval_async": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, _P]),
    "nsb_eval_decode_async": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, C.c_int, _P, _P, _P, _P]),
    "nsb_eval_positions_async": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, _P]),
    "nsb_eval_positions_decode_async": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, C.c_int, _P, _P, _P, _P]),
    "nsb_await": (C.c_int, [_P, C.c_int]),
    "nsb_is_computing": (C.c_int, [_P, C.c_int]),
    "nsb_eval_device": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, _P]),
    "nsb_eval_decode_device": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, C.c_int, _P, _P, _P, _P, _P]),
    "nsb_extract_device": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.c_int, C.c_int, _P]),
    "nsb_pack_positions_device": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P]),
    "nsb_decode_device": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_size_t, _P, _P, C.c_int, _P, _P]),
    "nsb_decode_device_ex": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_size_t, _P, _P, C.c_int, _P, _P, _P, _P]),
    "nsb_cache_create": (C.c_int, [_P, C.c_size_t]),
    "nsb_cache_clear": (C.c_int, [_P]),
    "nsb_cache_num_bundles": (C.c_uint64, [_P]),
    "nsb_cache_store_device": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, _P, _P, _P, _P]),
    "nsb_cache_probe_device": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, _P, _P, _P, _P, _P]),
    "nsb_eval_cached_decode_async": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P]),
    "nsb_eval_positions_cached_decode_async": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, _P, C.c_int, _P, _P, _P, _P,
                                                         _P]),
    "nsb_eval_cached_decode_device": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P]),
    "nsb_stream": (_P, [_P, C.c_int]),
    "nsb_set_timing": (C.c_int, [_P, C.c_int]),
    "nsb_trunk_time": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "nsb_trunk_time_reset": (C.c_int, [_P]),
    "nsb_launch_count": (C.c_uint64, [_P]),
    "nsb_trunk_kernel_name": (C.c_char_p, [_P]),
    "nsb_host_register": (C.c_int, [_P, C.c_size_t]),
    "nsb_host_unregister": (C.c_int, [_P]),
    "nsb_eval_request_async": (C.c_int, [_P, C.c_int, _P]),
    "nsb_cache_attach": (C.c_int, [_P, _P]),
    "nsb_set_io_mode": (C.c_int, [_P, C.c_int]),
    "nsb_io_mode": (C.c_int, [_P]),
    "nsb_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "nsb_host_free": (C.c_int, [_P]),
    "nsb_host_alloc_near": (C.c_int, [C.POINTER(_P), C.c_size_t, C.c_int]),
    "nsb_gpu_numa_node": (C.c_int, [C.c_int]),
    "nsb_numa_bind_thread": (C.c_int, [C.c_int]),
    "nsb_device_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "nsb_device_free": (C.c_int, [_P]),
    "nsb_memcpy_h2d": (C.c_int, [_P, _P, C.c_size_t]),
    "nsb_memcpy_d2h": (C.c_int, [_P, _P, C.c_size_t]),
    "nsb_memset_device": (C.c_int, [_P, C.c_int, C.c_size_t]),
    "nsb_device_sync": (C.c_int, []),
    "nsb_last_error": (C.c_char_p, []),
    "nsb_version": (C.c_char_p, []),
    "nsb_device_count": (C.c_int, []),
    "nsb_umma_selftest": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "nsb_event_create": (C.c_int, [C.POINTER(_P)]),
    "nsb_event_destroy": (C.c_int, [_P]),
    "nsb_event_record": (C.c_int, [_P, _P, C.c_int]),
    "nsb_event_sync": (C.c_int, [_P]),
    "nsb_event_elapsed_ms": (C.c_int, [_P, _P, C.POINTER(C.c_float)]),
}


# the additional symbols of the diagnostic build (include/nsb_diag.h, libnsb_diag.so)
DIAG_SIGNATURES = {
    "nsb_debug_trunk_timeline": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, C.c_size_t]),
    "nsb_debug_trunk_timeline_positions": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, C.c_size_t]),
    "nsb_debug_umma_probe": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                       C.POINTER(C.c_double)]),
    "nsb_debug_bulk_rate_probe": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
}
DIAG_LIB_PATH = os.path.join(HERE, "libnsb_diag.so")
_diag = None


def diag_lib() -> C.CDLL:
    """The diagnostic build of the library (-DNSB_DIAG): every product symbol plus include/nsb_diag.h.  Tools and the
    tests that compare against the superseded kernels load it; the product path never does."""
    global _diag
    if _diag is None:
        if not os.path.exists(DIAG_LIB_PATH):
            raise NsbError(f"{DIAG_LIB_PATH} not built")
        l = C.CDLL(DIAG_LIB_PATH)
        for name, (res, args) in {**SIGNATURES, **DIAG_SIGNATURES}.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _diag = l
    return _diag


def lib() -> C.CDLL:
    """Load libnsb.so (built in-tree by ``__graft_entry__.build()`` / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NsbError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().nsb_last_error().decode()
        if _diag is not None and _diag.nsb_last_error():   # the call may have gone to the diagnostic build
            msg = msg or _diag.nsb_last_error().decode()
        raise NsbError(f"{what} failed ({rc}): {msg}")


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "host buffers must be C-contiguous"
        return a.ctypes.data
    return int(a)


def device_count() -> int:
    return lib().nsb_device_count()


def net_desc(channels: int, blocks: int, value_hidden: int = 256, in_channels: int = FEATURE_CHANNELS) -> NetDesc:
    return NetDesc(in_channels, channels, blocks, value_hidden)


def random_blob(desc: NetDesc, seed: int) -> np.ndarray:
    n = lib().nsb_weight_blob_floats(C.byref(desc))
    blob = np.empty(n, dtype=np.float32)
    _check(lib().nsb_weight_blob_random(C.byref(desc), seed, blob.ctypes.data), "nsb_weight_blob_random")
    return blob


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (reference pins with cudaHostRegister,
    src/evaluate/evaluator.cc:95-106)."""

    def __init__(self, shape, dtype, gpu: Optional[int] = None):
        """gpu: place the pages on that GPU's NUMA node (nsb_host_alloc_near)."""
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = _P()
        if gpu is None:
            _check(lib().nsb_host_alloc(C.byref(p), nbytes), "nsb_host_alloc")
        else:
            _check(lib().nsb_host_alloc_near(C.byref(p), nbytes, gpu), "nsb_host_alloc_near")
        self._p = p
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p is not None:
            self.array = None
            lib().nsb_host_free(self._p)
            self._p = None


class DecodeRequest(C.Structure):
    """include/nsb.h nsb_decode_request"""
    _fields_ = [("features", _P), ("positions", _P), ("n", C.c_size_t), ("hashes", _P), ("move_off", _P), ("move_idx", _P),
                ("mode", C.c_int), ("legal_out", _P), ("order_out", _P), ("win", _P), ("draw", _P), ("nan_flag", _P),
                ("hit_flag", _P), ("row_flags", _P), ("logits_out", _P)]


def host_register(arr: np.ndarray) -> bool:
    """True if this call page-locked the array, False if it already was (NSB_HOST_ALREADY_LOCKED)."""
    rc = lib().nsb_host_register(arr.ctypes.data, arr.nbytes)
    if rc < 0:
        _check(rc, "nsb_host_register")
    return rc == 0


def host_unregister(arr: np.ndarray):
    _check(lib().nsb_host_unregister(arr.ctypes.data), "nsb_host_unregister")


class DeviceBuffer:
    def __init__(self, nbytes: int):
        p = _P()
        _check(lib().nsb_device_alloc(C.byref(p), nbytes), "nsb_device_alloc")
        self.ptr = p.value
        self.nbytes = nbytes

    @classmethod
    def from_host(cls, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        d = cls(max(a.nbytes, 1))
        if a.nbytes:
            _check(lib().nsb_memcpy_h2d(d.ptr, a.ctypes.data, a.nbytes), "nsb_memcpy_h2d")
        return d

    def to_host(self, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        if out.nbytes:
            _check(lib().nsb_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes), "nsb_memcpy_d2h")
        return out

    def fill(self, byte: int):
        _check(lib().nsb_memset_device(self.ptr, byte, self.nbytes), "nsb_memset_device")

    def free(self):
        if self.ptr:
            lib().nsb_device_free(self.ptr)
            self.ptr = None


def gpu_numa_node(gpu: int = 0) -> int:
    return lib().nsb_gpu_numa_node(gpu)


def numa_bind_thread(gpu: int = 0) -> bool:
    """Pin the calling thread to the CPUs of the GPU's NUMA node; False when there is nothing to do here."""
    return lib().nsb_numa_bind_thread(gpu) == 0


def device_sync():
    _check(lib().nsb_device_sync(), "nsb_device_sync")


def umma_selftest(n_cols: int, k_elems: int, shift_rows: int, gpu: int = 0):
    """Returns (max |D - ref|, max |epilogue record - ref|); both 0.0 when the tcgen05 path is right."""
    err, epi = C.c_float(-1.0), C.c_float(-1.0)
    _check(lib().nsb_umma_selftest(gpu, n_cols, k_elems, shift_rows, C.byref(err), C.byref(epi)),
           "nsb_umma_selftest")
    return float(err.value), float(epi.value)


def umma_probe(n_cols, k_elems, shift_rows, layout, iters=20, gpu=0):
    err, cyc = C.c_float(-1.0), C.c_double(0)
    _check(diag_lib().nsb_debug_umma_probe(gpu, n_cols, k_elems, shift_rows, layout, iters, C.byref(err), C.byref(cyc)),
           "nsb_debug_umma_probe")
    return float(err.value), float(cyc.value)


def bulk_rate_probe(ctas, tile_bytes, stages, split, gpu=0):
    v = C.c_double(0)
    _check(diag_lib().nsb_debug_bulk_rate_probe(gpu, ctas, tile_bytes, stages, split, C.byref(v)), "nsb_debug_bulk_rate_probe")
    return float(v.value)


class Event:
    """CUDA event recorded on a slot's stream (timing from the host side of the ABI)."""

    def __init__(self):
        p = _P()
        _check(lib().nsb_event_create(C.byref(p)), "nsb_event_create")
        self._p = p

    def record(self, ctx: "Context", slot: int = 0):
        _check(ctx._l.nsb_event_record(self._p, ctx._h, slot), "nsb_event_record")

    def sync(self):
        _check(lib().nsb_event_sync(self._p), "nsb_event_sync")

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float(0)
        _check(lib().nsb_event_elapsed_ms(self._p, stop._p, C.byref(ms)), "nsb_event_elapsed_ms")
        return float(ms.value)

    def destroy(self):
        if self._p is not None:
            lib().nsb_event_destroy(self._p)
            self._p = None


class Context:
    """One nsb_ctx: the C-ABI twin of one reference ``infer::Infer`` instance
    (reference src/infer/infer.h:19-32), with ``slots`` independent in-flight batches."""

    def __init__(self, desc: NetDesc, batch_max: int, slots: int = 1, gpu: int = 0,
                 blob: Optional[np.ndarray] = None, seed: Optional[int] = None, diag: bool = False):
        self.desc = desc
        self.batch_max = batch_max
        self.slots = slots
        self._l = diag_lib() if diag else lib()   # diag: the -DNSB_DIAG build (timeline stamps, superseded kernels)
        h = _P()
        _check(self._l.nsb_create(C.byref(h), gpu, batch_max, slots, C.byref(desc)), "nsb_create")
        self._h = h
        if blob is None and seed is not None:
            blob = random_blob(desc, seed)
        if blob is not None:
            self.load_weights(blob)

    def close(self):
        if self._h is not None:
            self._l.nsb_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def load_weights(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        _check(self._l.nsb_load_weights(self._h, blob.ctypes.data, blob.size), "nsb_load_weights")

    def bind_thread(self):
        _check(self._l.nsb_bind_thread(self._h), "nsb_bind_thread")

    # -- host-buffer calls (the Infer contract) ------------------------------------------------
    def eval_async(self, slot, features, n, policy, win, draw):
        _check(self._l.nsb_eval_async(self._h, slot, _ptr(features), n, _ptr(policy), _ptr(win), _ptr(draw)),
               "nsb_eval_async")

    def eval_decode_async(self, slot, features, n, move_off, move_idx, mode, legal_out, win, draw, nan_flag=None):
        _check(self._l.nsb_eval_decode_async(self._h, slot, _ptr(features), n, _ptr(move_off), _ptr(move_idx), mode,
                                           _ptr(legal_out), _ptr(win), _ptr(draw), _ptr(nan_flag)),
               "nsb_eval_decode_async")

    def eval_positions_async(self, slot, positions, n, policy, win, draw):
        _check(self._l.nsb_eval_positions_async(self._h, slot, _ptr(positions), n, _ptr(policy), _ptr(win),
                                              _ptr(draw)), "nsb_eval_positions_async")

    def eval_positions_decode_async(self, slot, positions, n, move_off, move_idx, mode, legal_out, win, draw,
                                    nan_flag=None):
        _check(self._l.nsb_eval_positions_decode_async(self._h, slot, _ptr(positions), n, _ptr(move_off),
                                                     _ptr(move_idx), mode, _ptr(legal_out), _ptr(win), _ptr(draw),
                                                     _ptr(nan_flag)), "nsb_eval_positions_decode_async")

    def await_(self, slot=0):
        _check(self._l.nsb_await(self._h, slot), "nsb_await")

    def is_computing(self, slot=0) -> bool:
        rc = self._l.nsb_is_computing(self._h, slot)
        if rc < 0:
            _check(rc, "nsb_is_computing")
        return rc == 1

    # -- device-pointer calls -------------------------------------------------------------------
    def eval_device(self, slot, d_features, n, d_policy, d_win, d_draw):
        _check(self._l.nsb_eval_device(self._h, slot, _ptr(d_features), n, _ptr(d_policy), _ptr(d_win), _ptr(d_draw)),
               "nsb_eval_device")

    def eval_decode_device(self, slot, d_features, n, d_off, d_idx, mode, d_policy, d_legal, d_win, d_draw, d_flag):
        _check(self._l.nsb_eval_decode_device(self._h, slot, _ptr(d_features), n, _ptr(d_off), _ptr(d_idx), mode,
                                            _ptr(d_policy), _ptr(d_legal), _ptr(d_win), _ptr(d_draw), _ptr(d_flag)),
               "nsb_eval_decode_device")

    def extract_device(self, slot, d_features, n, channels, channels_first, d_planes):
        _check(self._l.nsb_extract_device(self._h, slot, _ptr(d_features), n, channels, int(channels_first),
                                        _ptr(d_planes)), "nsb_extract_device")

    def pack_positions_device(self, slot, d_positions, n, d_features):
        _check(self._l.nsb_pack_positions_device(self._h, slot, _ptr(d_positions), n, _ptr(d_features)),
               "nsb_pack_positions_device")

    def decode_device(self, slot, d_policy, d_win, d_draw, n, d_off, d_idx, mode, d_legal, d_flag):
        _check(self._l.nsb_decode_device(self._h, slot, _ptr(d_policy), _ptr(d_win), _ptr(d_draw), n, _ptr(d_off),
                                       _ptr(d_idx), mode, _ptr(d_legal), _ptr(d_flag)), "nsb_decode_device")

    def decode_device_ex(self, slot, d_policy, d_win, d_draw, n, d_off, d_idx, mode, d_row_flags, d_legal, d_logits, d_flag):
        _check(self._l.nsb_decode_device_ex(self._h, slot, _ptr(d_policy), _ptr(d_win), _ptr(d_draw), n, _ptr(d_off),
                                          _ptr(d_idx), mode, _ptr(d_row_flags), _ptr(d_legal), _ptr(d_logits), _ptr(d_flag)),
               "nsb_decode_device_ex")

    # -- device-resident evaluation cache ------------------------------------------------------------
    def cache_create(self, memory_mb: int):
        _check(self._l.nsb_cache_create(self._h, memory_mb), "nsb_cache_create")

    def cache_clear(self):
        _check(self._l.nsb_cache_clear(self._h), "nsb_cache_clear")

    def cache_num_bundles(self) -> int:
        return int(self._l.nsb_cache_num_bundles(self._h))

    def cache_store_device(self, slot, d_hashes, n, d_off, d_legal, d_win, d_draw, d_skip=None, d_stored=None):
        _check(self._l.nsb_cache_store_device(self._h, slot, _ptr(d_hashes), n, _ptr(d_off), _ptr(d_legal), _ptr(d_win),
                                            _ptr(d_draw), _ptr(d_skip), _ptr(d_stored)), "nsb_cache_store_device")

    def cache_probe_device(self, slot, d_hashes, n, d_off, d_legal, d_win, d_draw, d_hit, d_miss_idx, d_miss_count):
        _check(self._l.nsb_cache_probe_device(self._h, slot, _ptr(d_hashes), n, _ptr(d_off), _ptr(d_legal), _ptr(d_win),
                                            _ptr(d_draw), _ptr(d_hit), _ptr(d_miss_idx), _ptr(d_miss_count)),
               "nsb_cache_probe_device")

    def eval_cached_decode_async(self, slot, features, n, hashes, move_off, move_idx, mode, legal_out, win, draw,
                                 nan_flag=None, hit_flag=None):
        _check(self._l.nsb_eval_cached_decode_async(self._h, slot, _ptr(features), n, _ptr(hashes), _ptr(move_off),
                                                  _ptr(move_idx), mode, _ptr(legal_out), _ptr(win), _ptr(draw),
                                                  _ptr(nan_flag), _ptr(hit_flag)), "nsb_eval_cached_decode_async")

    def eval_positions_cached_decode_async(self, slot, positions, n, hashes, move_off, move_idx, mode, legal_out, win,
                                           draw, nan_flag=None, hit_flag=None):
        _check(self._l.nsb_eval_positions_cached_decode_async(self._h, slot, _ptr(positions), n, _ptr(hashes),
                                                            _ptr(move_off), _ptr(move_idx), mode, _ptr(legal_out),
                                                            _ptr(win), _ptr(draw), _ptr(nan_flag), _ptr(hit_flag)),
               "nsb_eval_positions_cached_decode_async")

    def eval_cached_decode_device(self, slot, d_features, n, d_hashes, d_off, d_idx, mode, d_legal, d_win, d_draw,
                                  d_flag, d_hit):
        _check(self._l.nsb_eval_cached_decode_device(self._h, slot, _ptr(d_features), n, _ptr(d_hashes), _ptr(d_off),
                                                   _ptr(d_idx), mode, _ptr(d_legal), _ptr(d_win), _ptr(d_draw),
                                                   _ptr(d_flag), _ptr(d_hit)), "nsb_eval_cached_decode_device")

    def debug_trunk_timeline(self, slot, d_features, n, positions=False):
        nl = 2 * self.desc.blocks + 2
        out = np.zeros(nl * 4 + 16 + 3 * 1024, dtype=np.uint64)
        fn = self._l.nsb_debug_trunk_timeline_positions if positions else self._l.nsb_debug_trunk_timeline
        _check(fn(self._h, slot, _ptr(d_features), n, out.ctypes.data, out.size), "nsb_debug_trunk_timeline")
        return out[:nl * 4].reshape(nl, 4), out[nl * 4:]

    def eval_request_async(self, slot, n, move_off, move_idx, mode, legal_out, win, draw, features=None, positions=None,
                           hashes=None, order_out=None, nan_flag=None, hit_flag=None, row_flags=None, logits_out=None):
        """nsb_eval_request_async: the general fused call (cache when `hashes`, rank order when `order_out`,
        per-row flags and raw logits beside the probabilities for DECODE_BOTH)."""
        r = DecodeRequest(_ptr(features), _ptr(positions), n, _ptr(hashes), _ptr(move_off), _ptr(move_idx), mode,
                          _ptr(legal_out), _ptr(order_out), _ptr(win), _ptr(draw), _ptr(nan_flag), _ptr(hit_flag),
                          _ptr(row_flags), _ptr(logits_out))
        _check(self._l.nsb_eval_request_async(self._h, slot, C.byref(r)), "nsb_eval_request_async")

    def cache_attach(self, owner: "Context"):
        _check(self._l.nsb_cache_attach(self._h, owner._h), "nsb_cache_attach")

    def set_io_mode(self, direct: bool):
        _check(self._l.nsb_set_io_mode(self._h, 1 if direct else 0), "nsb_set_io_mode")

    def io_mode(self) -> str:
        return "direct" if self._l.nsb_io_mode(self._h) == 1 else "staged"

    def stream(self, slot=0) -> int:
        return self._l.nsb_stream(self._h, slot) or 0

    def set_timing(self, on: bool):
        _check(self._l.nsb_set_timing(self._h, int(on)), "nsb_set_timing")

    def trunk_time(self):
        s, n = C.c_double(0), C.c_uint64(0)
        _check(self._l.nsb_trunk_time(self._h, C.byref(s), C.byref(n)), "nsb_trunk_time")
        return float(s.value), int(n.value)

    def trunk_time_reset(self):
        _check(self._l.nsb_trunk_time_reset(self._h), "nsb_trunk_time_reset")

    def trunk_kernel_name(self) -> str:
        return self._l.nsb_trunk_kernel_name(self._h).decode()

    def launch_count(self) -> int:
        return int(self._l.nsb_launch_count(self._h))
