#!/usr/bin/env python
"""Duration of the trunk kernel itself (CUDA events around the launch) and of a whole submit -> await step, with
staged and with direct host I/O, one batch in flight.  usage: io_kernel_time.py [repo root to load the library from]
(a second checkout, e.g. `git worktree add _old <commit>` + make, gives an A/B on the same box)."""
import os, sys, time
import numpy as np
ROOT = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package(); nb, synth = pkg.binding, pkg.synth
B=256
desc = nb.net_desc(128, 10)
ctx = nb.Context(desc, batch_max=B, slots=1, seed=1234)
pos = synth.random_positions(B, seed=1)
off, idx = synth.random_legal_moves(B, seed=20240203, edge_rows=False)
nm = int(off[-1])
P = nb.PinnedArray
h_pos=P((B,), nb.POSITION); h_pos.array[:]=pos
h_off=P((B+1,), np.uint32); h_off.array[:]=off
h_idx=P((nm,), np.uint16); h_idx.array[:]=idx
h_legal=P((nm,), np.float32); h_win=P((B,), np.float32); h_draw=P((B,), np.float32); h_flag=P((B,), np.uint8)
for direct in (False, True):
    ctx.set_io_mode(direct)
    for _ in range(50):
        ctx.eval_positions_decode_async(0, h_pos.array, B, h_off.array, h_idx.array, 0, h_legal.array, h_win.array, h_draw.array, h_flag.array); ctx.await_(0)
    ctx.set_timing(True); ctx.trunk_time_reset()
    t0=time.perf_counter()
    for _ in range(500):
        ctx.eval_positions_decode_async(0, h_pos.array, B, h_off.array, h_idx.array, 0, h_legal.array, h_win.array, h_draw.array, h_flag.array); ctx.await_(0)
    dt=(time.perf_counter()-t0)/500*1e6
    s,n = ctx.trunk_time()
    ctx.set_timing(False)
    print(ROOT, "direct" if direct else "staged", f"kernel {s/n*1e3:.1f} us   step {dt:.1f} us (timing on)")
