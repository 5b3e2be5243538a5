#!/usr/bin/env python
"""Generates tests/golden/*.npz.  Run in the BUILD container (needs /root/reference for the
reference's own Random executor, compiled in place by oracle/Makefile into oracle/_ref/).

  random_ref.npz   : outputs of the REFERENCE's src/infer/random.cc (seed 0 and 7, 3 samples)
  expand_kat.npz   : 96 fuzz FeatureBitboards and their planes from an independent numpy
                     restatement of src/cuda/extractbit.cu (tests/helpers.py), NCHW and NHWC
  forward_small.npz: 4 seeded positions through the canonical 1 x 128 net: oracle fp32 and
                     oracle bf16-emulating outputs, after checking the fp32 one against PyTorch
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
import helpers  # noqa: E402

pkg = graft.load_package()
orc = graft.load_oracle()
nb, synth = pkg.binding, pkg.synth
out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

# 1. the reference's Random executor
assert orc.have_ref_random(), "oracle/_ref/libnsb_ref_random.so missing (needs /root/reference)"
rec = {}
for seed in (0, 7):
    p, w, d = orc.ref_random_fill(seed, 3)
    rec[f"policy_head_{seed}"] = p[:, :8].copy()
    rec[f"policy_tail_{seed}"] = p[:, -4:].copy()
    rec[f"win_{seed}"] = w
    rec[f"draw_{seed}"] = d
    rec[f"sha256_{seed}"] = np.frombuffer(hashlib.sha256(p.tobytes() + w.tobytes() + d.tobytes()).digest(),
                                          dtype=np.uint8)
np.savez(os.path.join(out, "random_ref.npz"), **rec)

# 2. expand known answers
fb = synth.random_feature_bitboards(96, seed=11)
np.savez(os.path.join(out, "expand_kat.npz"), lo=fb["lo"], hi=fb["hi"],
         nchw=helpers.expand_numpy(fb, 2, 48, True).view(np.uint32),
         nhwc=helpers.expand_numpy(fb, 2, 48, False).view(np.uint32))

# 3. forward of a small canonical net
desc = nb.net_desc(128, 1)
blob = nb.random_blob(desc, 4321)
pos = synth.random_positions(4, seed=99)
planes = orc.expand(orc.pack(pos), 4)
p32, w32, d32 = orc.forward(desc, blob, planes, emulate_bf16=False)
tp, tw, td = helpers.forward_torch(desc, blob, planes)
assert np.max(np.abs(p32 - tp)) < 2e-4 and np.max(np.abs(w32 - tw)) < 1e-5, "oracle fp32 != torch fp32"
p16, w16, d16 = orc.forward(desc, blob, planes, emulate_bf16=True)
np.savez_compressed(os.path.join(out, "forward_small.npz"), positions=pos.view(np.uint8).reshape(4, 108),
                    blob_seed=np.int64(4321), blob_sha256=np.frombuffer(hashlib.sha256(blob.tobytes()).digest(),
                                                                       dtype=np.uint8),
                    policy_fp32=p32, win_fp32=w32, draw_fp32=d32, policy_bf16=p16, win_bf16=w16, draw_bf16=d16)
print("golden written to", out, {f: os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)})

# 4. evaluation-cache trace from the REFERENCE's own src/mcts/evalcache.cc (compiled in place):
#    4000 operations over 18 keys that collide in 3 bundles, rows are a function of (win, j)
assert orc.have_ref_evalcache(), "oracle/_ref/libnsb_ref_evalcache.so missing (needs /root/reference)"
ref = orc.RefCache(1)
NB = ref.num_bundles
rng = np.random.default_rng(20240203)
keys = np.array([b + k * NB for b in (3, 4, 5) for k in range(6)], dtype=np.uint64)
n_ops = 4000
op = (rng.random(n_ops) < 0.5).astype(np.uint8)                       # 1 = store, 0 = load
key = keys[rng.integers(0, len(keys), size=n_ops)]
cnt = rng.choice(np.array([1, 3, 3, 3, 5, 164, 165], dtype=np.uint32), size=n_ops)
win = rng.random(n_ops).astype(np.float32)
draw = rng.random(n_ops).astype(np.float32)
res = np.zeros(n_ops, dtype=np.uint8)
got_win = np.zeros(n_ops, dtype=np.float32)
got_last = np.zeros(n_ops, dtype=np.float32)
for i in range(n_ops):
    if op[i]:
        row = (win[i] + np.arange(cnt[i], dtype=np.float32)).astype(np.float32)
        res[i] = ref.store(int(key[i]), row, float(win[i]), float(draw[i]))
    else:
        ok, row, w, d = ref.load(int(key[i]), int(cnt[i]))
        res[i] = ok
        if ok:
            got_win[i], got_last[i] = w, row[-1]
ref.close()
np.savez_compressed(os.path.join(out, "evalcache_trace.npz"), num_bundles=np.uint64(NB), op=op, key=key, cnt=cnt, win=win,
                    draw=draw, res=res, got_win=got_win, got_last=got_last)
print("evalcache trace:", n_ops, "ops,", int(res[op == 0].sum()), "hits of", int((op == 0).sum()), "loads")
