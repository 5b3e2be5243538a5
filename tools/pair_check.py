#!/usr/bin/env python
"""CTA-pair 256-channel trunk vs the single-CTA kernel and the oracle (diagnostics)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package(); nb, synth = pkg.binding, pkg.synth
orc = graft.load_oracle()
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
desc = nb.net_desc(256, blocks)
blob = nb.random_blob(desc, 99)
pos = synth.random_positions(n, seed=256)
fb = orc.pack(pos)
def run(mode):
    if mode: os.environ["NSB_TRUNK256"] = mode
    else: os.environ.pop("NSB_TRUNK256", None)
    policy = np.zeros((n, nb.POLICY_SIZE), dtype=np.float32); win = np.zeros(n, dtype=np.float32); draw = np.zeros(n, dtype=np.float32)
    with nb.Context(desc, batch_max=max(n, 8), blob=blob, diag=bool(mode)) as ctx:   # "single": diagnostic build
        ctx.eval_async(0, fb, n, policy, win, draw); ctx.await_(0)
    return policy, win, draw
ps, ws, ds = run("single")
print("single done", flush=True)
pp, wp, dp = run(None)
print("pair done", flush=True)
print("pair vs single: max|dlogit| =", np.max(np.abs(pp - ps)), " bit-identical:", np.array_equal(pp, ps) and np.array_equal(wp, ws))
if n <= 64:
    op, ow, od = orc.forward(desc, blob, orc.expand(fb, n), emulate_bf16=True)
    print("pair vs oracle: max|dlogit| =", np.max(np.abs(pp - op)), " max|dwin| =", np.max(np.abs(wp - ow)))
    print("single vs oracle: max|dlogit| =", np.max(np.abs(ps - op)))
