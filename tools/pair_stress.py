#!/usr/bin/env python
"""Race hunt for the CTA-pair kernel: many launches of a deep 256-channel net over several streams,
every result compared bit for bit with the first one (and with the one-CTA kernel)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package(); nb, synth = pkg.binding, pkg.synth
orc = graft.load_oracle()
blocks, n, rounds = 20, 777, int(sys.argv[1]) if len(sys.argv) > 1 else 30
desc = nb.net_desc(256, blocks)
blob = nb.random_blob(desc, 99)
pos = synth.random_positions(n, seed=256)
fb = orc.pack(pos)
off, idx = synth.random_legal_moves(n, seed=4, edge_rows=False)
def run(ctx, slot):
    legal = np.zeros(int(off[-1]), dtype=np.float32); win = np.zeros(n, dtype=np.float32); draw = np.zeros(n, dtype=np.float32)
    ctx.eval_decode_async(slot, fb, n, off, idx, nb.DECODE_LOGITS, legal, win, draw, None)
    return legal, win, draw
os.environ["NSB_TRUNK256"] = "single"
with nb.Context(desc, batch_max=n, blob=blob, diag=True) as ctx:   # the one-CTA kernel lives in libnsb_diag.so
    ref = run(ctx, 0); ctx.await_(0)
os.environ.pop("NSB_TRUNK256")
bad = 0
with nb.Context(desc, batch_max=n, slots=4, blob=blob) as ctx:
    for r in range(rounds):
        outs = [run(ctx, s) for s in range(4)]
        for s in range(4):
            ctx.await_(s)
        for o in outs:
            if not (np.array_equal(o[0].view(np.uint32), ref[0].view(np.uint32)) and np.array_equal(o[1], ref[1])):
                bad += 1
print(f"pair_stress: {rounds * 4} launches of {blocks}x256 on {n} positions, mismatching launches: {bad}")
sys.exit(1 if bad else 0)
