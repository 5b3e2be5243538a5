#!/usr/bin/env python
"""Soak run for races and rare hangs: for SECONDS, four host threads drive four contexts on one GPU at once -
10x128 duo (4 slots, staged), 10x128 classic (1 slot, direct I/O, ranked, shared cache), a second classic executor
attached to the same cache, and a 256-channel CTA-pair context - each re-evaluating a fixed batch and comparing
every result bit for bit with its first one.  usage: soak.py [seconds]"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nb, synth = pkg.binding, pkg.synth
SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
P = nb.PinnedArray
stop_at = time.time() + SECONDS
errors, counts = [], {}


def worker(name, channels, blocks, slots, n, cached_with=None, own_cache=False, ranked=False):
    desc = nb.net_desc(channels, blocks)
    blob = nb.random_blob(desc, 7)
    ctx = nb.Context(desc, batch_max=n, slots=slots, blob=blob)
    if own_cache:
        ctx.cache_create(64)
    if cached_with is not None:
        ctx.cache_attach(cached_with)
    handles[name] = ctx
    ready.wait()
    pos = synth.random_positions(n, seed=100 + sorted(NAMES).index(name))
    off, idx = synth.random_legal_moves(n, seed=3, edge_rows=False)
    total = int(off[-1])
    # distinct key ranges per worker: two workers that share the cache must not collide by construction (a collision
    # with an equal move count is, as in the reference, served as a hit - and would fail the bit-for-bit comparison)
    hashes = (np.arange(n, dtype=np.uint64) + np.uint64(1_000_003 * (1 + sorted(NAMES).index(name)))) * np.uint64(0x9E3779B97F4A7C15)
    bufs = []
    for s in range(slots):
        b = dict(pos=P((n,), nb.POSITION), off=P((n + 1,), np.uint32), idx=P((total,), np.uint16), legal=P((total,), np.float32),
                 order=P((total,), np.uint16), win=P((n,), np.float32), draw=P((n,), np.float32), flag=P((n,), np.uint8),
                 hit=P((n,), np.uint8), hashes=P((n,), np.uint64))
        b["pos"].array[:], b["off"].array[:], b["idx"].array[:], b["hashes"].array[:] = pos, off, idx, hashes
        bufs.append(b)
    use_cache = own_cache or cached_with is not None

    def submit(s):
        b = bufs[s]
        ctx.eval_request_async(s, n, b["off"].array, b["idx"].array, nb.DECODE_PROBS, b["legal"].array, b["win"].array,
                               b["draw"].array, positions=b["pos"].array, order_out=b["order"].array if ranked else None,
                               nan_flag=b["flag"].array, hashes=b["hashes"].array if use_cache else None,
                               hit_flag=b["hit"].array if use_cache else None)

    submit(0)
    ctx.await_(0)
    ref = (bufs[0]["legal"].array.copy(), bufs[0]["win"].array.copy(), bufs[0]["order"].array.copy())
    it = 0
    while time.time() < stop_at and not errors:
        for s in range(slots):
            submit(s)
        for s in range(slots):
            ctx.await_(s)
            b = bufs[s]
            ok = np.array_equal(b["legal"].array.view(np.uint32), ref[0].view(np.uint32)) and np.array_equal(b["win"].array, ref[1])
            if ranked:
                ok = ok and np.array_equal(b["order"].array, ref[2])
            if not ok:
                errors.append(f"{name}: iteration {it} slot {s} differs from the first result")
            it += 1
        if use_cache and it % 64 == 0 and own_cache:
            ctx.cache_clear()          # misses again: the trunk path keeps running under the shared cache
    counts[name] = it


NAMES = ["classic+cache", "classic+attached", "duo", "pair"]
handles, ready = {}, threading.Event()
t1 = threading.Thread(target=worker, args=("classic+cache", 128, 10, 1, 200), kwargs=dict(own_cache=True, ranked=True))
t1.start()
while "classic+cache" not in handles:
    time.sleep(0.01)
threads = [t1,
           threading.Thread(target=worker, args=("classic+attached", 128, 10, 1, 130), kwargs=dict(cached_with=handles["classic+cache"], ranked=True)),
           threading.Thread(target=worker, args=("duo", 128, 10, 4, int(sys.argv[2]) if len(sys.argv) > 2 else 700), kwargs=dict(ranked=True)),
           threading.Thread(target=worker, args=("pair", 256, 6, 2, 301))]
for t in threads[1:]:
    t.start()
while len(handles) < 4:
    time.sleep(0.01)
ready.set()
for t in threads:
    t.join()
print("soak", "FAILED" if errors else "ok", f"{SECONDS:.0f} s", counts, errors[:3])
sys.exit(1 if errors else 0)
