#!/usr/bin/env python
"""Per-layer timeline of CTA 0 of the fused trunk (clock64 stamps) — where do the cycles go?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nb, synth = pkg.binding, pkg.synth
C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 10
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
desc = nb.net_desc(C, blocks)
ctx = nb.Context(desc, batch_max=B, seed=1234, diag=True)   # the clock64 stamps exist in libnsb_diag.so only
POSITIONS = os.environ.get("NSB_TIMELINE_POSITIONS") == "1"   # feed packed positions: stage 1 in the prologue
fb = synth.random_positions(B, seed=1) if POSITIONS else synth.random_feature_bitboards(B * 86, seed=1)
d_fb = nb.DeviceBuffer.from_host(fb)
for _ in range(3):
    t, ph = ctx.debug_trunk_timeline(0, d_fb.ptr, B, positions=POSITIONS)
    t, ph = t.astype(np.int64), ph.astype(np.int64)
t0 = t[0, 0]
print(f"C={C} blocks={blocks} B={B}  (cycles; CTA 0, pass 0)")
print("layer  mma_start  mma_issue  acc_wait(after issue)  epilogue  act_wait(next mma start - epi end)  layer_total")
for L in range(len(t)):
    ms, me, es, ee = t[L]
    nxt = t[L + 1, 0] if L + 1 < len(t) else 0
    print(f"{L:3d} {ms - t0:9d} {me - ms:9d} {es - me if es else -1:9d} {ee - es if ee else -1:9d} "
          f"{nxt - ee if (nxt and ee) else -1:9d} {nxt - ms if nxt else -1:9d}")
print("total cycles (layer 0 start -> last mma issued):", t[-1, 1] - t0)
names = ["entry", "setup done", "features expanded+arrived", "features loaded+barrier", "heads read", "policy written",
         "value MLP done", "decode done", "pass start", "loads done (thread)", "expansion done (thread)", "role start"]
order = [0, 1, 11, 8, 9, 3, 10, 2, 4, 5, 6, 7]
print("phase stamps relative to kernel entry (cycles):")
for k in order:
    print(f"  {names[k]:20s} {ph[k] - ph[0]:9d}")
print("  layer-0 MMA start    ", t0 - ph[0], "   last MMA issued", t[-1, 1] - ph[0])
if len(ph) > 12 and ph[12]:
    print("  MMA warp: cycles spent waiting for weight stages (whole launch):", ph[12])
