#!/usr/bin/env python
"""One launch of the PRODUCT library's trunk_duo_kernel with two co-resident CTAs per SM (592 positions = 296 CTAs), for
an ncu capture of the kernel in the state the pipeline keeps it in (profiles/r2_duo128_coresident_ncu_summary.txt).
usage: coresident.py [positions] [launches]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nb, synth = pkg.binding, pkg.synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 6
desc = nb.net_desc(128, 10)
ctx = nb.Context(desc, batch_max=B, slots=2, seed=1)
assert ctx.trunk_kernel_name().startswith("trunk_duo_kernel")
fb = synth.random_feature_bitboards(B * 86, seed=1)
off, idx = synth.random_legal_moves(B, seed=2, edge_rows=False)
d_fb, d_off, d_idx = (nb.DeviceBuffer.from_host(a) for a in (fb, off, idx))
d_legal, d_win, d_draw, d_flag = nb.DeviceBuffer(int(off[-1]) * 4), nb.DeviceBuffer(B * 4), nb.DeviceBuffer(B * 4), nb.DeviceBuffer(B)
e0, e1 = nb.Event(), nb.Event()
for i in range(launches):
    e0.record(ctx, 0)
    ctx.eval_decode_device(0, d_fb.ptr, B, d_off.ptr, d_idx.ptr, nb.DECODE_PROBS, None, d_legal.ptr, d_win.ptr, d_draw.ptr, d_flag.ptr)
    e1.record(ctx, 0)
    e1.sync()
    ctx.await_(0)
    print(f"launch {i}: {B} positions, {e0.elapsed_ms(e1) * 1e3:.1f} us")
