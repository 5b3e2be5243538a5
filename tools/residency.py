#!/usr/bin/env python
"""Do the two CTAs of an SM really run side by side in trunk_duo_kernel?  Every CTA records its SM id and the
globaltimer at start and end; this prints, per SM, the CTAs and how much their lifetimes overlap.
usage: residency.py [batch]   (592 positions = 296 CTAs = 2 per SM)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NSB_TRUNK128"] = "duo"
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nb, synth = pkg.binding, pkg.synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
desc = nb.net_desc(128, 10)
ctx = nb.Context(desc, batch_max=B, slots=2, seed=1234, diag=True)
fb = synth.random_feature_bitboards(B * 86, seed=1)
d_fb = nb.DeviceBuffer.from_host(fb)
for _ in range(3):
    t, ph = ctx.debug_trunk_timeline(0, d_fb.ptr, B)
rec = ph[16:16 + 3 * 1024].reshape(1024, 3).astype(np.int64)
n_cta = min((B + 1) // 2, 1024)
rec = rec[:n_cta]
t0 = rec[:, 1].min()
by_sm = {}
for i, (sm, a, b) in enumerate(rec):
    by_sm.setdefault(int(sm), []).append((i, (a - t0) / 1e3, (b - t0) / 1e3))
print(f"{ctx.trunk_kernel_name()}  B={B}: {n_cta} CTAs on {len(by_sm)} SMs; kernel span {(rec[:, 2].max() - t0) / 1e3:.1f} us")
overlaps = []
for sm, ctas in sorted(by_sm.items()):
    if len(ctas) >= 2:
        (i0, a0, b0), (i1, a1, b1) = ctas[0], ctas[1]
        ov = max(0.0, min(b0, b1) - max(a0, a1))
        overlaps.append(ov / max(b0 - a0, 1e-9))
ran = rec[rec[:, 2] > 0]
print(f"CTAs that ran: {len(ran)} of {n_cta}")
t0 = ran[:, 1].min()
by_sm = {}
for i, (sm, a, b) in enumerate(rec):
    if b > 0:
        by_sm.setdefault(int(sm), []).append((i, (a - t0) / 1e3, (b - t0) / 1e3))
overlaps = []
for sm, ctas in sorted(by_sm.items()):
    if len(ctas) >= 2:
        (i0, a0, b0), (i1, a1, b1) = ctas[0], ctas[1]
        overlaps.append(max(0.0, min(b0, b1) - max(a0, a1)) / max(b0 - a0, 1e-9))
for sm in list(sorted(by_sm))[:4]:
    print(f"  SM {sm}: " + ", ".join(f"CTA {i}: {a:.1f}-{b:.1f} us" for i, a, b in by_sm[sm]))
if overlaps:
    print(f"SMs with two CTAs: {len(overlaps)}; lifetime overlap of the pair: median {100 * np.median(overlaps):.0f} %, min {100 * min(overlaps):.0f} %")
print(f"CTA lifetime: median {np.median((ran[:, 2] - ran[:, 1]) / 1e3):.1f} us; kernel span {(ran[:, 2].max() - t0) / 1e3:.1f} us")
