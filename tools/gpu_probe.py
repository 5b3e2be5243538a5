#!/usr/bin/env python
"""First-contact diagnostics on a GPU box: descriptor conventions, then a tiny forward."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
orc = graft.load_oracle()
nb, synth = pkg.binding, pkg.synth
print("devices:", nb.device_count(), nb.lib().nsb_version().decode(), flush=True)
for (n, k, sh) in ((192, 16, 0), (192, 64, 0), (192, 64, 11), (96, 128, 1)):
    print(f"umma N={n} K={k} shift={sh}: (mma_err, epilogue_err)={nb.umma_selftest(n, k, sh)}", flush=True)
for (C, blocks, n) in ((128, 1, 2), (128, 2, 5), (256, 1, 3)):
    desc = nb.net_desc(C, blocks)
    blob = nb.random_blob(desc, 1234)
    pos = synth.random_positions(n, seed=3)
    fb = orc.pack(pos)
    policy = np.zeros((n, 2187), np.float32)
    win = np.zeros(n, np.float32)
    draw = np.zeros(n, np.float32)
    with nb.Context(desc, batch_max=8, blob=blob) as ctx:
        ctx.eval_async(0, fb, n, policy, win, draw)
        ctx.await_(0)
    op, ow, od = orc.forward(desc, blob, orc.expand(fb, n), emulate_bf16=True)
    print(f"C={C} blocks={blocks} n={n}: max|dlogit|={np.max(np.abs(policy - op)):.3e} "
          f"max|dwin|={np.max(np.abs(win - ow)):.3e} logits range [{op.min():.2f},{op.max():.2f}] "
          f"gpu range [{policy.min():.2f},{policy.max():.2f}]", flush=True)
    if np.max(np.abs(policy - op)) > 1e-2:
        d = np.abs(policy - op)
        print("  per-position max:", d.max(axis=1), " per-plane max (pos0):", d[0].reshape(27, 81).max(axis=1)[:6], flush=True)
