#!/usr/bin/env python
"""Tiny end-to-end case for compute-sanitizer (memcheck / racecheck): both trunk kernels with the
fused decode, the stage kernels and the device cache, a few positions each."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package(); nb, synth = pkg.binding, pkg.synth
for C in (128, 256):
    desc = nb.net_desc(C, 1)
    n = 5
    pos = synth.random_positions(n, seed=7)
    off, idx = synth.random_legal_moves(n, seed=7)
    legal = np.zeros(int(off[-1]), dtype=np.float32)
    win = np.zeros(n, dtype=np.float32); draw = np.zeros(n, dtype=np.float32); flag = np.zeros(n, dtype=np.uint8)
    hit = np.zeros(n, dtype=np.uint8)
    hashes = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    with nb.Context(desc, batch_max=8, slots=1, seed=3) as ctx:
        ctx.eval_positions_decode_async(0, pos, n, off, idx, nb.DECODE_PROBS, legal, win, draw, flag)
        ctx.await_(0)
        ctx.cache_create(1)
        for _ in range(2):
            ctx.eval_positions_cached_decode_async(0, pos, n, hashes, off, idx, nb.DECODE_PROBS, legal, win, draw, flag, hit)
            ctx.await_(0)
        print(C, "hits", int(hit.sum()), "win", win[:2])
        d_pos = nb.DeviceBuffer.from_host(pos); d_fb = nb.DeviceBuffer(n * 86 * 16); d_pl = nb.DeviceBuffer(n * 86 * 81 * 4)
        ctx.pack_positions_device(0, d_pos.ptr, n, d_fb.ptr)
        ctx.extract_device(0, d_fb.ptr, n, 86, True, d_pl.ptr)
        ctx.extract_device(0, d_fb.ptr, n, 86, False, d_pl.ptr)
        ctx.await_(0)
print("sanitize_case ok")
