#!/usr/bin/env python
"""trunk_fused_kernel<128> in clusters that share the weight stream by multicast (NSB_TRUNK128=mc2 / mc4) against the
same kernel without clusters: outputs must be bit-identical for every batch size (odd sizes and sizes that leave
cluster CTAs without positions included); then launch durations by CUDA events.

    python tools/mcast_check.py [blocks] [mode ...]        # default: 10 blocks, modes mc2 mc4
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nb, synth = pkg.binding, pkg.synth
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 10
modes = sys.argv[2:] or ["mc2", "mc4"]
desc = nb.net_desc(128, blocks)
blob = nb.random_blob(desc, 99)
SIZES = [1, 2, 3, 4, 5, 7, 8, 64, 255, 256, 296, 300, 512, 1024]
NMAX = max(SIZES)
pos = synth.random_positions(NMAX, seed=128)
off, idx = synth.random_legal_moves(NMAX, seed=5, edge_rows=True)


def run(mode):
    os.environ["NSB_TRUNK128"] = mode
    out = {}
    with nb.Context(desc, batch_max=NMAX, blob=blob, slots=1) as ctx:
        d_pos = nb.DeviceBuffer.from_host(pos)
        d_fb = nb.DeviceBuffer(NMAX * 86 * 16)
        ctx.pack_positions_device(0, d_pos.ptr, NMAX, d_fb.ptr)
        ctx.await_(0)
        fb = d_fb.to_host((NMAX, 86), nb.FEATURE_BITBOARD)
        for n in SIZES:
            policy = np.zeros((n, nb.POLICY_SIZE), dtype=np.float32)
            win = np.zeros(n, dtype=np.float32)
            draw = np.zeros(n, dtype=np.float32)
            ctx.eval_async(0, fb[:n], n, policy, win, draw)
            ctx.await_(0)
            out[n] = (policy, win, draw)
            print(f"  {mode} n={n} ok", flush=True)
        # launch duration, back to back on one stream
        times = {}
        for n in (256, 296, 512, 1024):
            policy = np.zeros((n, nb.POLICY_SIZE), dtype=np.float32)
            win = np.zeros(n, dtype=np.float32)
            draw = np.zeros(n, dtype=np.float32)
            for _ in range(5):
                ctx.eval_async(0, fb[:n], n, policy, win, draw)
                ctx.await_(0)
            t0 = time.perf_counter()
            for _ in range(50):
                ctx.eval_async(0, fb[:n], n, policy, win, draw)
                ctx.await_(0)
            times[n] = (time.perf_counter() - t0) / 50 * 1e3
        print(f"{mode}: kernel {ctx.trunk_kernel_name()}; host-timed ms per batch (dense logits out): " +
              ", ".join(f"B={n}: {t:.3f}" for n, t in times.items()), flush=True)
    return out


ref = run("classic")
bad = 0
for m in modes:
    got = run(m)
    for n in SIZES:
        same = all(np.array_equal(a, b) for a, b in zip(ref[n], got[n]))
        if not same:
            bad += 1
            rows = [r for r in range(n) if not np.array_equal(ref[n][0][r], got[n][0][r])]
            nan_rows = [r for r in range(n) if np.isnan(got[n][0][r]).any()]
            print(f"MISMATCH {m} n={n}: rows that differ {rows[:12]}{'...' if len(rows) > 12 else ''} ({len(rows)} of {n}); rows with NaN {nan_rows[:12]} "
                  f"({len(nan_rows)}); first differing row: ref {ref[n][0][rows[0]][:4]} got {got[n][0][rows[0]][:4]}")
    print(f"{m}: {'bit-identical to the un-clustered kernel at every size' if not bad else 'DIFFERS'}")
sys.exit(1 if bad else 0)
