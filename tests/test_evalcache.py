"""Evaluation cache (reference src/mcts/evalcache.{h,cc}): the oracle restatement is pinned against
the reference's own evalcache.cc compiled in place (oracle/_ref) and against a committed trace
produced by it; the host mirror (EvalCacheB200) and the device-resident cache (cache_device.cuh)
are then compared with the oracle operation by operation."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "nshogi-engine_b200", "host")


def row_of(win, n):
    return (np.float32(win) + np.arange(n, dtype=np.float32)).astype(np.float32)


def random_ops(num_bundles, n_ops, seed, bundles=(3, 4, 5), ways=6):
    rng = np.random.default_rng(seed)
    keys = np.array([b + k * num_bundles for b in bundles for k in range(ways)], dtype=np.uint64)
    return ((rng.random(n_ops) < 0.5), keys[rng.integers(0, len(keys), size=n_ops)],
            rng.choice(np.array([1, 3, 3, 3, 5, 164, 165], dtype=np.uint32), size=n_ops),
            rng.random(n_ops).astype(np.float32), rng.random(n_ops).astype(np.float32))


def run_ops(cache, ops):
    """-> list of (result, win, last row element) per operation."""
    out = []
    for is_store, key, cnt, win, draw in zip(*ops):
        if is_store:
            out.append((cache.store(int(key), row_of(win, cnt), float(win), float(draw)), 0.0, 0.0))
        else:
            ok, row, w, _ = cache.load(int(key), int(cnt))
            out.append((ok, w if ok else 0.0, float(row[-1]) if ok else 0.0))
    return out


def test_oracle_matches_committed_reference_trace(orc, golden_dir):
    """tests/golden/evalcache_trace.npz was produced by the REFERENCE's evalcache.cc (tools/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "evalcache_trace.npz"))
    c = orc.Cache(int(g["num_bundles"]))
    got = run_ops(c, (g["op"].astype(bool), g["key"], g["cnt"], g["win"], g["draw"]))
    c.close()
    assert [int(r[0]) for r in got] == g["res"].tolist()
    loads = ~g["op"].astype(bool) & g["res"].astype(bool)
    assert np.array_equal(np.array([r[1] for r in got], dtype=np.float32)[loads], g["got_win"][loads])
    assert np.array_equal(np.array([r[2] for r in got], dtype=np.float32)[loads], g["got_last"][loads])
    assert loads.sum() > 100


def test_oracle_matches_reference_evalcache_live(orc):
    if not orc.have_ref_evalcache():
        pytest.skip("oracle/_ref/libnsb_ref_evalcache.so not built (needs /root/reference)")
    ref = orc.RefCache(1)
    c = orc.Cache(ref.num_bundles)
    for seed in (1, 2, 3):
        ops = random_ops(ref.num_bundles, 6000, seed)
        assert run_ops(ref, ops) == run_ops(c, ops)
    ref.close()
    c.close()


def test_reference_replacement_quirk(orc):
    """evalcache.cc:75-86 never repairs the old head's Prev: after three distinct stores in a bundle a
    hit on the last element does not move it, and the next new key overwrites it."""
    nb_ = 7
    c = orc.Cache(nb_)
    keys = [42 + k * nb_ for k in range(4)]
    for k in keys[:3]:
        assert c.store(k, row_of(0.5, 3), 0.5, 0.0)
    assert c.load(keys[0], 3)[0]
    assert c.store(keys[3], row_of(0.25, 3), 0.25, 0.0)
    assert [c.load(k, 3)[0] for k in keys] == [False, True, True, True]
    c.close()


def test_host_cache_matches_oracle(orc, pkg):
    """EvalCacheB200 (host/eval_cache.h) replays the same operations through `nsb_host_unit --cache-trace`."""
    subprocess.check_call(["make", "-C", HOST, "-s", "all"])
    exe = os.path.join(HOST, "nsb_host_unit")
    probe = subprocess.run([exe, "--cache-trace", "1"], input="", capture_output=True, text=True, timeout=60)
    num_bundles = int(probe.stdout.split()[1])
    ops = random_ops(num_bundles, 5000, 11)
    lines = []
    for is_store, key, cnt, win, draw in zip(*ops):
        lines.append(f"s {int(key)} {int(cnt)} {float(win)!r} {float(draw)!r}" if is_store else f"l {int(key)} {int(cnt)}")
    out = subprocess.run([exe, "--cache-trace", "1"], input="\n".join(lines) + "\n", capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0
    got = out.stdout.strip().splitlines()[1:]
    c = orc.Cache(num_bundles)
    want = run_ops(c, ops)
    c.close()
    assert len(got) == len(want)
    for g, w, is_store, cnt in zip(got, want, ops[0], ops[2]):
        f = g.split()
        assert int(f[0]) == int(w[0]), (g, w)
        if not is_store and w[0]:
            assert np.float32(f[1]) == np.float32(w[1]) and np.float32(f[3]) == np.float32(w[2])


# ---- device-resident cache ------------------------------------------------------------------------------------------
def _batches(num_bundles, n_batches, batch, seed):
    """Batches whose positions fall into distinct bundles (no intra-batch lock collisions), drawn from a
    key pool that collides heavily ACROSS batches."""
    rng = np.random.default_rng(seed)
    bundles = np.arange(5, 5 + 4 * batch)
    for _ in range(n_batches):
        b = rng.choice(bundles, size=batch, replace=False)
        keys = (b + rng.integers(0, 6, size=batch) * num_bundles).astype(np.uint64)
        cnt = rng.choice(np.array([1, 2, 3, 3, 7, 80, 164, 165, 200], dtype=np.uint32), size=batch)
        yield keys, cnt, rng.random(batch).astype(np.float32), rng.random(batch).astype(np.float32), rng.random() < 0.5


@pytest.mark.gpu
def test_device_cache_matches_oracle(nb, orc):
    desc = nb.net_desc(128, 1)
    with nb.Context(desc, batch_max=64, seed=1) as ctx:
        ctx.cache_create(1)
        num_bundles = ctx.cache_num_bundles()
        assert num_bundles > 400
        c = orc.Cache(num_bundles)
        hits = 0
        for keys, cnt, win, draw, is_store in _batches(num_bundles, 120, 48, 5):
            n = len(keys)
            off = np.zeros(n + 1, dtype=np.uint32)
            off[1:] = np.cumsum(cnt)
            rows = np.concatenate([row_of(w, k) for w, k in zip(win, cnt)])
            d_keys, d_off = nb.DeviceBuffer.from_host(keys), nb.DeviceBuffer.from_host(off)
            d_win, d_draw = nb.DeviceBuffer.from_host(win), nb.DeviceBuffer.from_host(draw)
            if is_store:
                d_rows, d_stored = nb.DeviceBuffer.from_host(rows), nb.DeviceBuffer(n)
                ctx.cache_store_device(0, d_keys.ptr, n, d_off.ptr, d_rows.ptr, d_win.ptr, d_draw.ptr, None, d_stored.ptr)
                ctx.await_(0)
                got = d_stored.to_host((n,), np.uint8)
                want = [c.store(int(k), row_of(w, m), float(w), float(d)) for k, m, w, d in zip(keys, cnt, win, draw)]
                assert got.tolist() == [int(x) for x in want]
                bufs = [d_rows, d_stored]
            else:
                d_rows = nb.DeviceBuffer.from_host(np.full(len(rows), -1.0, dtype=np.float32))
                d_hit, d_miss, d_cnt = nb.DeviceBuffer(n), nb.DeviceBuffer(4 * n), nb.DeviceBuffer(4)
                ctx.cache_probe_device(0, d_keys.ptr, n, d_off.ptr, d_rows.ptr, d_win.ptr, d_draw.ptr, d_hit.ptr, d_miss.ptr,
                                       d_cnt.ptr)
                ctx.await_(0)
                hit = d_hit.to_host((n,), np.uint8)
                out_rows = d_rows.to_host((len(rows),), np.float32)
                out_win, out_draw = d_win.to_host((n,), np.float32), d_draw.to_host((n,), np.float32)
                n_miss = int(d_cnt.to_host((1,), np.int32)[0])
                miss = np.sort(d_miss.to_host((n,), np.int32)[:n_miss])
                want = [c.load(int(k), int(m)) for k, m in zip(keys, cnt)]
                assert hit.tolist() == [int(w[0]) for w in want]
                assert miss.tolist() == [i for i, w in enumerate(want) if not w[0]]
                for i, w in enumerate(want):
                    if w[0]:
                        hits += 1
                        assert np.array_equal(out_rows[off[i]:off[i + 1]], w[1])
                        assert out_win[i] == np.float32(w[2]) and out_draw[i] == np.float32(w[3])
                    else:   # a miss leaves the caller's buffers alone
                        assert np.all(out_rows[off[i]:off[i + 1]] == -1.0) and out_win[i] == win[i]
                bufs = [d_rows, d_hit, d_miss, d_cnt]
            for b in [d_keys, d_off, d_win, d_draw] + bufs:
                b.free()
        c.close()
        assert hits > 100


@pytest.mark.gpu
def test_device_cache_same_bundle_in_one_batch(nb):
    """Positions of ONE batch that share a bundle race for its try-lock (evalcache.cc:58-62): losers are
    dropped, never corrupted - afterwards every key that reports 'stored' loads back its own row."""
    desc = nb.net_desc(128, 1)
    with nb.Context(desc, batch_max=64, seed=1) as ctx:
        ctx.cache_create(1)
        nbund = ctx.cache_num_bundles()
        n = 64
        keys = (9 + np.arange(n, dtype=np.uint64) * nbund).astype(np.uint64)      # all in bundle 9
        cnt = np.full(n, 5, dtype=np.uint32)
        off = np.zeros(n + 1, dtype=np.uint32)
        off[1:] = np.cumsum(cnt)
        win = (np.arange(n) / 100.0).astype(np.float32)
        rows = np.concatenate([row_of(w, 5) for w in win])
        d = [nb.DeviceBuffer.from_host(x) for x in (keys, off, rows, win, win)]
        d_stored = nb.DeviceBuffer(n)
        ctx.cache_store_device(0, d[0].ptr, n, d[1].ptr, d[2].ptr, d[3].ptr, d[4].ptr, None, d_stored.ptr)
        ctx.await_(0)
        stored = d_stored.to_host((n,), np.uint8)
        assert 1 <= stored.sum() <= n
        # probe one key at a time (no contention): at most 3 entries survive, each intact
        alive = 0
        for i in range(n):
            dk, do = nb.DeviceBuffer.from_host(keys[i:i + 1]), nb.DeviceBuffer.from_host(np.array([0, 5], dtype=np.uint32))
            dr, dw, dd = nb.DeviceBuffer(20), nb.DeviceBuffer(4), nb.DeviceBuffer(4)
            dh, dm, dc = nb.DeviceBuffer(1), nb.DeviceBuffer(4), nb.DeviceBuffer(4)
            ctx.cache_probe_device(0, dk.ptr, 1, do.ptr, dr.ptr, dw.ptr, dd.ptr, dh.ptr, dm.ptr, dc.ptr)
            ctx.await_(0)
            if dh.to_host((1,), np.uint8)[0]:
                alive += 1
                assert stored[i] == 1
                assert np.array_equal(dr.to_host((5,), np.float32), row_of(win[i], 5))
                assert dw.to_host((1,), np.float32)[0] == win[i]
            for b in (dk, do, dr, dw, dd, dh, dm, dc):
                b.free()
        assert 1 <= alive <= 3
        for b in d + [d_stored]:
            b.free()


@pytest.mark.gpu
@pytest.mark.parametrize("channels,direct", [(128, False), (256, False), (128, True), (256, True)])
def test_eval_cached_serves_hits_and_evaluates_misses(nb, orc, synth, channels, direct):
    """nsb_eval_cached_decode_async: first call evaluates everything and fills the cache, the second call
    is served from it bit for bit, a mixed batch evaluates only the new positions (trunk launch on the
    probe's miss list), and a changed move count is a miss (searchworker.cc:546).  `direct`: every buffer
    of the cached calls is mapped page-locked memory and the packed positions go in, so the probe and the
    trunk work on the caller's buffers themselves (NSB_IO_DIRECT, stage 1 in the prologue, misses only)."""
    desc = nb.net_desc(channels, 2)
    blob = nb.random_blob(desc, 5)
    n = 37
    pos = synth.random_positions(2 * n, seed=31)
    fb = orc.pack(pos).reshape(2 * n, 86)
    off_all, idx_all = synth.random_legal_moves(2 * n, seed=3, edge_rows=False)
    hashes = (np.arange(2 * n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(12345))

    def csr(sel):
        cnt = np.diff(off_all)[sel]
        off = np.zeros(len(sel) + 1, dtype=np.uint32)
        off[1:] = np.cumsum(cnt)
        idx = np.concatenate([idx_all[off_all[i]:off_all[i + 1]] for i in sel]).astype(np.uint16)
        return off, idx

    def run(ctx, sel, cached=True):
        off, idx = csr(sel)
        m = len(sel)
        legal = np.zeros(int(off[-1]), dtype=np.float32)
        win, draw = np.zeros(m, dtype=np.float32), np.zeros(m, dtype=np.float32)
        flag, hit = np.zeros(m, dtype=np.uint8), np.zeros(m, dtype=np.uint8)
        f = np.ascontiguousarray(fb[sel].reshape(-1))
        if cached and direct:
            host = [np.ascontiguousarray(pos[sel]), np.ascontiguousarray(hashes[sel]), off, idx, legal, win, draw, flag, hit]
            pins = []
            for h in host:
                p = nb.PinnedArray(h.shape, h.dtype)
                p.array[...] = h
                pins.append(p)
            arrs = [p.array for p in pins]
            ctx.eval_positions_cached_decode_async(0, arrs[0], m, arrs[1], arrs[2], arrs[3], nb.DECODE_PROBS, *arrs[4:])
            ctx.await_(0)
            for h, p in zip(host[4:], pins[4:]):
                h[...] = p.array
            for p in pins:
                p.free()
            return legal, win, draw, hit, off
        if cached:
            ctx.eval_cached_decode_async(0, f, m, np.ascontiguousarray(hashes[sel]), off, idx, nb.DECODE_PROBS, legal, win,
                                         draw, flag, hit)
        else:
            ctx.eval_decode_async(0, f, m, off, idx, nb.DECODE_PROBS, legal, win, draw, flag)
        ctx.await_(0)
        return legal, win, draw, hit, off

    with nb.Context(desc, batch_max=2 * n, blob=blob) as ctx:
        ctx.cache_create(4)
        assert ctx.io_mode() == "direct"                                   # one-slot ctx
        first = list(range(n))
        base = run(ctx, first, cached=False)
        l0 = ctx.launch_count()
        a = run(ctx, first)
        assert a[3].sum() == 0 and ctx.launch_count() - l0 == 2          # probe + trunk
        assert np.array_equal(a[0].view(np.uint32), base[0].view(np.uint32)) and np.array_equal(a[1], base[1])
        b = run(ctx, first)
        cacheable = np.diff(a[4]) <= 164
        assert np.array_equal(b[3].astype(bool), cacheable)              # rows > 164 moves are never stored
        assert np.array_equal(b[0].view(np.uint32), base[0].view(np.uint32))
        assert np.array_equal(b[1], base[1]) and np.array_equal(b[2], base[2])
        mixed = list(range(n // 2, n + n // 2))                          # half known, half new, shuffled
        np.random.default_rng(0).shuffle(mixed)
        ref = run(ctx, mixed, cached=False)
        c = run(ctx, mixed)
        known = np.array([i < n for i in mixed]) & (np.diff(c[4]) <= 164)
        assert np.array_equal(c[3].astype(bool), known)
        assert np.array_equal(c[0].view(np.uint32), ref[0].view(np.uint32)) and np.array_equal(c[1], ref[1])
        # a different move count for a cached hash is a miss and gets evaluated
        off2, idx2 = csr(first)
        cnt = np.diff(off2).copy()
        victim = int(np.argmax((cnt > 2) & (cnt <= 164)))
        cnt[victim] -= 1
        off3 = np.zeros(n + 1, dtype=np.uint32)
        off3[1:] = np.cumsum(cnt)
        idx3 = np.concatenate([idx2[off2[i]:off2[i] + cnt[i]] for i in range(n)]).astype(np.uint16)
        legal = np.zeros(int(off3[-1]), dtype=np.float32)
        win, draw = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        hit = np.zeros(n, dtype=np.uint8)
        ctx.eval_cached_decode_async(0, np.ascontiguousarray(fb[first].reshape(-1)), n, np.ascontiguousarray(hashes[first]),
                                     off3, idx3, nb.DECODE_PROBS, legal, win, draw, None, hit)
        ctx.await_(0)
        assert hit[victim] == 0 and abs(legal[off3[victim]:off3[victim + 1]].sum() - 1.0) < 1e-5


@pytest.mark.gpu
def test_two_executors_share_one_cache(nb, orc, synth):
    """nsb_cache_attach: the reference's evaluation workers (two executors per GPU by default, context.h:75) feed ONE
    EvalCache (manager.cc:202-206).  What executor A evaluated is a hit for executor B, bit for bit, and both can
    run at once."""
    desc = nb.net_desc(128, 1)
    blob = nb.random_blob(desc, 5)
    n = 48
    pos = synth.random_positions(n, seed=3)
    off, idx = synth.random_legal_moves(n, seed=3, edge_rows=False)
    total = int(off[-1])
    hashes = np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(99)

    def call(ctx, start=True):
        legal = np.zeros(total, dtype=np.float32)
        win, draw = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        hit = np.zeros(n, dtype=np.uint8)
        ctx.eval_positions_cached_decode_async(0, pos, n, hashes, off, idx, nb.DECODE_PROBS, legal, win, draw, None, hit)
        return legal, win, draw, hit

    with nb.Context(desc, batch_max=n, blob=blob) as a, nb.Context(desc, batch_max=n, blob=blob) as b:
        a.cache_create(4)
        b.cache_attach(a)
        assert b.cache_num_bundles() == a.cache_num_bundles()
        ra = call(a)
        a.await_(0)
        rb = call(b)
        b.await_(0)
        cacheable = np.diff(off) <= 164
        assert ra[3].sum() == 0 and np.array_equal(rb[3].astype(bool), cacheable)
        assert np.array_equal(rb[0].view(np.uint32), ra[0].view(np.uint32)) and np.array_equal(rb[1], ra[1])
        a.cache_clear()
        ra, rb = call(a), call(b)        # both in flight on the same table
        a.await_(0)
        b.await_(0)
        assert np.array_equal(rb[0].view(np.uint32), ra[0].view(np.uint32))
        with pytest.raises(Exception):
            a.cache_attach(a)
