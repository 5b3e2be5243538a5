#!/usr/bin/env python
"""Differential fuzz of the host-buffer entry points: for random nets / batch sizes / move lists (rows of 0 moves, 1 move,
164, 165, 593 moves included), every way of asking for the same evaluation must return the same bits -
bitboards in (staged) == packed positions in (direct I/O, ranked) == through the device cache (all misses) == served from
the cache (hits, rows <= 164 moves) - and the rank order must be the oracle's.  usage: fuzz_modes.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nb, synth = pkg.binding, pkg.synth
orc = graft.load_oracle()
SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
P = nb.PinnedArray
t_end = time.time() + SECONDS
iters = 0
blobs = {}


def pinned(a):
    p = P(a.shape, a.dtype)
    p.array[...] = a
    return p


while time.time() < t_end:
    channels = int(rng.choice([128, 256]))
    blocks = int(rng.integers(1, 3))
    slots = int(rng.integers(1, 3))
    n = int(rng.choice([1, 2, 3, int(rng.integers(4, 64)), int(rng.integers(64, 700))]))
    mode = int(rng.choice([nb.DECODE_PROBS, nb.DECODE_LOGITS, nb.DECODE_BOTH]))
    both = mode == nb.DECODE_BOTH
    row_flags = (rng.random(n) < 0.2).astype(np.uint8) * nb.ROW_SKIP_SOFTMAX   # Gumbel roots (read in BOTH mode only)
    desc = nb.net_desc(channels, blocks)
    key = (channels, blocks)
    if key not in blobs:
        blobs[key] = nb.random_blob(desc, 100 + channels + blocks)
    blob = blobs[key]
    pos = synth.random_positions(n, seed=int(rng.integers(0, 1 << 30)))
    fb = orc.pack(pos)
    cnt = rng.choice(np.array([0, 1, 2, 30, 80, 164, 165, 593]), size=n, p=[.05, .05, .1, .3, .3, .08, .07, .05]).astype(np.int64)
    off = np.zeros(n + 1, dtype=np.uint32)
    off[1:] = np.cumsum(cnt)
    total = int(off[-1])
    idx = rng.integers(0, 2187, size=max(total, 1), dtype=np.uint16)[:total] if total else np.zeros(0, dtype=np.uint16)
    hashes = rng.integers(1, 1 << 62, size=n, dtype=np.uint64)
    tag = f"C={channels} blocks={blocks} slots={slots} n={n} mode={mode} total={total}"

    def outs():
        return dict(legal=P((max(total, 1),), np.float32), order=P((max(total, 1),), np.uint16), win=P((n,), np.float32),
                    draw=P((n,), np.float32), flag=P((n,), np.uint8), hit=P((n,), np.uint8), logits=P((max(total, 1),), np.float32))

    h = dict(fb=pinned(fb), pos=pinned(pos), off=pinned(off), idx=pinned(idx if total else np.zeros(1, dtype=np.uint16)),
             hashes=pinned(hashes), rf=pinned(row_flags))
    res = {}
    with nb.Context(desc, batch_max=n, slots=slots, blob=blob) as ctx:
        ctx.cache_create(8)
        s_last = slots - 1

        def call(name, direct, **kw):
            o = outs()
            for a in o.values():
                a.array[...] = 0
            ctx.set_io_mode(direct)
            ctx.eval_request_async(s_last, n, h["off"].array, h["idx"].array, mode, o["legal"].array, o["win"].array, o["draw"].array,
                                   nan_flag=o["flag"].array, row_flags=h["rf"].array if both else None,
                                   logits_out=o["logits"].array if both else None, **{k: (v(o) if callable(v) else v) for k, v in kw.items()})
            ctx.await_(s_last)
            res[name] = {k: a.array.copy() for k, a in o.items()}
            for a in o.values():
                a.free()

        call("bb_staged", False, features=h["fb"].array)
        call("pos_direct_ranked", True, positions=h["pos"].array, order_out=lambda o: o["order"].array)
        call("cached_miss", True, positions=h["pos"].array, hashes=h["hashes"].array, hit_flag=lambda o: o["hit"].array,
             order_out=lambda o: o["order"].array)
        call("cached_hit", False, features=h["fb"].array, hashes=h["hashes"].array, hit_flag=lambda o: o["hit"].array,
             order_out=lambda o: o["order"].array)
    base = res["bb_staged"]
    for name in ("pos_direct_ranked", "cached_miss", "cached_hit"):
        r = res[name]
        assert np.array_equal(r["legal"][:total].view(np.uint32), base["legal"][:total].view(np.uint32)), (tag, name, "legal")
        assert np.array_equal(r["win"], base["win"]) and np.array_equal(r["draw"], base["draw"]), (tag, name, "win/draw")
        assert np.array_equal(r["logits"][:total].view(np.uint32), base["logits"][:total].view(np.uint32)), (tag, name, "logits")
        want = orc.rank_rows(base["legal"][:total], off, base["flag"])
        assert np.array_equal(r["order"][:total], want), (tag, name, "order")
    assert not base["flag"].any(), tag
    if both:   # Frame::setEvaluation: softmax of the raw logits, skipped at Gumbel roots (frame.cc:116-118)
        for b in range(n):
            lg, pr = base["logits"][off[b]:off[b + 1]], base["legal"][off[b]:off[b + 1]]
            if row_flags[b]:
                assert np.array_equal(lg.view(np.uint32), pr.view(np.uint32)), (tag, b, "skip")
            elif len(lg):
                e = np.exp(lg.astype(np.float64) - lg.max())
                assert np.allclose(pr, e / e.sum(), rtol=2e-6, atol=1e-9), (tag, b, "softmax")
    assert res["cached_miss"]["hit"].sum() == 0, tag
    # rows of more than 164 moves are never stored; the others are hits the second time unless their store was dropped
    # because another warp held the bundle's lock at that moment (the reference's try_lock semantics, evalcache.cc:58-62)
    hit = res["cached_hit"]["hit"].astype(bool)
    assert not hit[cnt > 164].any(), (tag, "hit on an uncacheable row")
    cacheable = int((cnt <= 164).sum())
    assert hit.sum() >= 0.9 * cacheable - 2, (tag, "hits", int(hit.sum()), cacheable)
    for a in h.values():
        a.free()
    iters += 1
print(f"fuzz ok: {iters} configurations in {SECONDS:.0f} s")
