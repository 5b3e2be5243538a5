"""ctypes binding of the CPU oracle (oracle/libnsb_oracle.so) and of the reference's own sources
compiled in place (oracle/_ref/*.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libnsb_oracle.so")
REF_RANDOM = os.path.join(HERE, "_ref", "libnsb_ref_random.so")
REF_EXTRACT = os.path.join(HERE, "_ref", "libnsb_ref_extractbit.so")
REF_EVALCACHE = os.path.join(HERE, "_ref", "libnsb_ref_evalcache.so")

POLICY_SIZE = 2187
FEATURE_CHANNELS = 86
_P = C.c_void_p
_lib = None
_ref_random = None


class NetDesc(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("channels", C.c_int32), ("blocks", C.c_int32),
                ("value_hidden", C.c_int32)]


def build(force: bool = False) -> None:
    if force or not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", HERE, "-s", "all"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB)
        l.nsb_oracle_expand.argtypes = [_P, C.c_size_t, C.c_int, C.c_int, _P]
        l.nsb_oracle_expand.restype = None
        l.nsb_oracle_pack.argtypes = [_P, C.c_size_t, _P]
        l.nsb_oracle_pack.restype = None
        l.nsb_oracle_decode.argtypes = [_P, _P, _P, C.c_size_t, _P, _P, C.c_int, _P, _P]
        l.nsb_oracle_decode.restype = None
        l.nsb_oracle_decode_ex.argtypes = [_P, _P, _P, C.c_size_t, _P, _P, C.c_int, _P, _P, _P, _P]
        l.nsb_oracle_decode_ex.restype = None
        l.nsb_oracle_value_fallback.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_double, C.c_double,
                                                C.c_uint64]
        l.nsb_oracle_value_fallback.restype = C.c_int
        l.nsb_oracle_dirichlet_mix.argtypes = [_P, _P, C.c_uint32]
        l.nsb_oracle_dirichlet_mix.restype = None
        l.nsb_oracle_forward.argtypes = [C.POINTER(NetDesc), _P, _P, C.c_size_t, C.c_int, _P, _P, _P]
        l.nsb_oracle_forward.restype = None
        l.nsb_oracle_rng_create.argtypes = [C.c_uint64]
        l.nsb_oracle_rng_create.restype = _P
        l.nsb_oracle_rng_destroy.argtypes = [_P]
        l.nsb_oracle_rng_destroy.restype = None
        l.nsb_oracle_random_fill.argtypes = [_P, C.c_size_t, _P, _P, _P]
        l.nsb_oracle_random_fill.restype = None
        l.nsb_oracle_cpu_path.argtypes = [_P, C.c_size_t, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P,
                                          C.POINTER(C.c_double)]
        l.nsb_oracle_cpu_path.restype = C.c_double
        l.nsb_oracle_bf16_round.argtypes = [C.c_float]
        l.nsb_oracle_bf16_round.restype = C.c_float
        _lib = l
    return _lib


def expand(fb: np.ndarray, n: int, channels: int = FEATURE_CHANNELS, channels_first: bool = True) -> np.ndarray:
    fb = np.ascontiguousarray(fb)
    shape = (n, channels, 81) if channels_first else (n, 81, channels)
    out = np.empty(shape, dtype=np.float32)
    lib().nsb_oracle_expand(fb.ctypes.data, n, channels, int(channels_first), out.ctypes.data)
    return out


def pack(pos: np.ndarray) -> np.ndarray:
    pos = np.ascontiguousarray(pos)
    fb = np.zeros(len(pos) * FEATURE_CHANNELS, dtype=np.dtype([("lo", "<u8"), ("hi", "<u8")]))
    lib().nsb_oracle_pack(pos.ctypes.data, len(pos), fb.ctypes.data)
    return fb


def decode(policy, win, draw, off, idx, mode: int):
    policy = np.ascontiguousarray(policy, dtype=np.float32)
    win = np.ascontiguousarray(win, dtype=np.float32)
    draw = np.ascontiguousarray(draw, dtype=np.float32)
    off = np.ascontiguousarray(off, dtype=np.uint32)
    idx = np.ascontiguousarray(idx, dtype=np.uint16)
    n = len(win)
    out = np.zeros(int(off[n]), dtype=np.float32)
    flag = np.zeros(n, dtype=np.uint8)
    lib().nsb_oracle_decode(policy.ctypes.data, win.ctypes.data, draw.ctypes.data, n, off.ctypes.data,
                            idx.ctypes.data, mode, out.ctypes.data, flag.ctypes.data)
    return out, flag


def decode_ex(policy, win, draw, off, idx, mode: int, row_flags=None, want_logits: bool = False):
    """nsb_oracle_decode_ex: returns (legal, logits or None, nan_flag)."""
    policy = np.ascontiguousarray(policy, dtype=np.float32)
    win = np.ascontiguousarray(win, dtype=np.float32)
    draw = np.ascontiguousarray(draw, dtype=np.float32)
    off = np.ascontiguousarray(off, dtype=np.uint32)
    idx = np.ascontiguousarray(idx, dtype=np.uint16)
    n = len(win)
    out = np.zeros(int(off[n]), dtype=np.float32)
    logits = np.zeros(int(off[n]), dtype=np.float32) if want_logits else None
    flag = np.zeros(n, dtype=np.uint8)
    rf = None if row_flags is None else np.ascontiguousarray(row_flags, dtype=np.uint8)
    lib().nsb_oracle_decode_ex(policy.ctypes.data, win.ctypes.data, draw.ctypes.data, n, off.ctypes.data, idx.ctypes.data,
                               mode, None if rf is None else rf.ctypes.data, out.ctypes.data,
                               None if logits is None else logits.ctypes.data, flag.ctypes.data)
    return out, logits, flag


def value_fallback(win: float, draw: float, has_parent: bool, parent_win_acc: float = 0.0, parent_draw_acc: float = 0.0,
                   parent_visits: int = 1):
    """feedworker.cc:58-85; returns (win, draw, NaNFound)."""
    w, d = C.c_float(win), C.c_float(draw)
    f = lib().nsb_oracle_value_fallback(C.byref(w), C.byref(d), int(has_parent), parent_win_acc, parent_draw_acc,
                                        parent_visits)
    return float(w.value), float(d.value), bool(f)


def dirichlet_mix(probs, noise):
    """frame.cc:121-133."""
    p = np.ascontiguousarray(probs, dtype=np.float32).copy()
    nz = np.ascontiguousarray(noise, dtype=np.float64)
    assert len(nz) >= len(p)
    lib().nsb_oracle_dirichlet_mix(p.ctypes.data, nz.ctypes.data, len(p))
    return p


def rank_rows(legal, off, flags=None):
    """Edge order after reference src/mcts/node.h:163-168 (Node::sort: std::sort of a node's edges by decreasing
    probability, called from feedworker.cc:129): per CSR row the permutation `order` with legal[order] non-increasing.
    std::sort leaves the order of equal elements unspecified; the restatement fixes it (lower index first = a stable
    sort), which is one of the results std::sort may produce.  A uniform row (the NaN-logit fallback) therefore keeps
    the generation order, and so does - by definition of the executor - a row that contains NaNs."""
    legal = np.asarray(legal, dtype=np.float32)
    off = np.asarray(off, dtype=np.int64)
    order = np.zeros(int(off[-1]), dtype=np.uint16)
    for b in range(len(off) - 1):
        row = legal[off[b]:off[b + 1]]
        if np.isnan(row).any():     # std::sort over NaNs is undefined in the reference; the executor defines: identity
            order[off[b]:off[b + 1]] = np.arange(len(row), dtype=np.uint16)
        else:
            order[off[b]:off[b + 1]] = np.argsort(-row.astype(np.float64), kind="stable").astype(np.uint16)
    return order


def forward(desc, blob: np.ndarray, planes: np.ndarray, emulate_bf16: bool):
    """fp32 definition of the canonical net; planes [n][in_channels][81] fp32."""
    d = NetDesc(desc.in_channels, desc.channels, desc.blocks, desc.value_hidden)
    blob = np.ascontiguousarray(blob, dtype=np.float32)
    planes = np.ascontiguousarray(planes, dtype=np.float32)
    n = planes.shape[0]
    policy = np.empty((n, POLICY_SIZE), dtype=np.float32)
    win = np.empty(n, dtype=np.float32)
    draw = np.empty(n, dtype=np.float32)
    lib().nsb_oracle_forward(C.byref(d), blob.ctypes.data, planes.ctypes.data, n, int(emulate_bf16),
                             policy.ctypes.data, win.ctypes.data, draw.ctypes.data)
    return policy, win, draw


def random_fill(seed: int, n: int):
    """Port of reference src/infer/random.cc:28-42."""
    r = lib().nsb_oracle_rng_create(seed)
    policy = np.empty((n, POLICY_SIZE), dtype=np.float32)
    win = np.empty(n, dtype=np.float32)
    draw = np.empty(n, dtype=np.float32)
    lib().nsb_oracle_random_fill(r, n, policy.ctypes.data, win.ctypes.data, draw.ctypes.data)
    lib().nsb_oracle_rng_destroy(r)
    return policy, win, draw


def have_ref_random() -> bool:
    return os.path.exists(REF_RANDOM)


def have_ref_extract() -> bool:
    return os.path.exists(REF_EXTRACT)


def ref_random():
    """The REFERENCE's own Random executor compiled in place from /root/reference (oracle/Makefile)."""
    global _ref_random
    if _ref_random is None:
        l = C.CDLL(REF_RANDOM)
        l.nsb_ref_random_make.argtypes = [C.c_uint64]
        l.nsb_ref_random_make.restype = _P
        l.nsb_ref_random_free.argtypes = [_P]
        l.nsb_ref_random_free.restype = None
        l.nsb_ref_random_fill.argtypes = [_P, C.c_size_t, _P, _P, _P]
        l.nsb_ref_random_fill.restype = None
        _ref_random = l
    return _ref_random


def ref_random_fill(seed: int, n: int):
    l = ref_random()
    h = l.nsb_ref_random_make(seed)
    policy = np.empty((n, POLICY_SIZE), dtype=np.float32)
    win = np.empty(n, dtype=np.float32)
    draw = np.empty(n, dtype=np.float32)
    l.nsb_ref_random_fill(h, n, policy.ctypes.data, win.ctypes.data, draw.ctypes.data)
    l.nsb_ref_random_free(h)
    return policy, win, draw


def ref_extract_device(d_dest: int, d_src: int, batch: int, channels: int, channels_first: bool) -> int:
    """The REFERENCE's own extractbit.cu kernels (needs a GPU); pointers are device addresses."""
    l = C.CDLL(REF_EXTRACT)
    l.nsb_ref_extract_bits.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int]
    l.nsb_ref_extract_bits.restype = C.c_int
    return l.nsb_ref_extract_bits(d_dest, d_src, batch, channels, int(channels_first))


def cpu_path(pos, off, idx, batch: int, threads: int, batches_per_thread: int, with_expand: bool,
             use_reference_random: bool):
    """Config-1 CPU path (pack [+expand] + Random fill + decode) on `threads` host threads.
    Returns (samples_per_second, seconds)."""
    pos = np.ascontiguousarray(pos)
    off = np.ascontiguousarray(off, dtype=np.uint32)
    idx = np.ascontiguousarray(idx, dtype=np.uint16)
    fill = mk = None
    if use_reference_random and have_ref_random():
        l = ref_random()
        fill = C.cast(l.nsb_ref_random_fill, _P)
        mk = C.cast(l.nsb_ref_random_make, _P)
    sec = C.c_double(0)
    v = lib().nsb_oracle_cpu_path(pos.ctypes.data, len(pos), off.ctypes.data, idx.ctypes.data, batch, threads,
                                  batches_per_thread, int(with_expand), fill, mk, C.byref(sec))
    return float(v), float(sec.value)


class Cache:
    """Restatement of reference src/mcts/evalcache.{h,cc} (single-threaded)."""

    def __init__(self, num_bundles: int):
        l = lib()
        l.nsb_oracle_cache_create.restype = _P
        l.nsb_oracle_cache_create.argtypes = [C.c_uint64]
        l.nsb_oracle_cache_destroy.argtypes = [_P]
        l.nsb_oracle_cache_store.argtypes = [_P, C.c_uint64, C.c_uint32, _P, C.c_float, C.c_float]
        l.nsb_oracle_cache_load.argtypes = [_P, C.c_uint64, C.c_uint32, _P, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        self._h = l.nsb_oracle_cache_create(num_bundles)

    def store(self, h: int, row: np.ndarray, win: float, draw: float) -> bool:
        row = np.ascontiguousarray(row, dtype=np.float32)
        return bool(lib().nsb_oracle_cache_store(self._h, h, len(row), row.ctypes.data, win, draw))

    def load(self, h: int, expected_n: int):
        row = np.zeros(max(expected_n, 1), dtype=np.float32)
        w, d = C.c_float(0), C.c_float(0)
        ok = bool(lib().nsb_oracle_cache_load(self._h, h, expected_n, row.ctypes.data, C.byref(w), C.byref(d)))
        return ok, row[:expected_n], float(w.value), float(d.value)

    def close(self):
        if self._h:
            lib().nsb_oracle_cache_destroy(self._h)
            self._h = None


def have_ref_evalcache() -> bool:
    return os.path.exists(REF_EVALCACHE)


class RefCache:
    """The reference's own EvalCache (src/mcts/evalcache.cc compiled in place into oracle/_ref)."""

    def __init__(self, memory_mb: int):
        l = C.CDLL(REF_EVALCACHE)
        l.nsb_ref_evalcache_make.restype = _P
        l.nsb_ref_evalcache_make.argtypes = [C.c_size_t]
        l.nsb_ref_evalcache_free.argtypes = [_P]
        l.nsb_ref_evalcache_num_bundles.restype = C.c_uint64
        l.nsb_ref_evalcache_num_bundles.argtypes = [_P]
        l.nsb_ref_evalcache_store.argtypes = [_P, C.c_uint64, C.c_uint32, _P, C.c_float, C.c_float]
        l.nsb_ref_evalcache_load.argtypes = [_P, C.c_uint64, C.c_uint32, _P, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        self._l = l
        self._h = l.nsb_ref_evalcache_make(memory_mb)
        self.num_bundles = int(l.nsb_ref_evalcache_num_bundles(self._h))

    def store(self, h: int, row: np.ndarray, win: float, draw: float) -> bool:
        row = np.ascontiguousarray(row, dtype=np.float32)
        return bool(self._l.nsb_ref_evalcache_store(self._h, h, len(row), row.ctypes.data, win, draw))

    def load(self, h: int, expected_n: int):
        row = np.zeros(max(expected_n, 164), dtype=np.float32)
        w, d = C.c_float(0), C.c_float(0)
        ok = bool(self._l.nsb_ref_evalcache_load(self._h, h, expected_n, row.ctypes.data, C.byref(w), C.byref(d)))
        return ok, row[:expected_n], float(w.value), float(d.value)

    def close(self):
        if self._h:
            self._l.nsb_ref_evalcache_free(self._h)
            self._h = None
