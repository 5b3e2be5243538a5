#!/usr/bin/env python
"""Roofline sweep (BASELINE.json configs 3-5 + the HBM-bound stages), device-timed with CUDA events.

  trunk : nets 10x128 / 20x256 / 40x256 over batch sizes 64..4096 -> evals/s, TFLOP/s and the
          fraction of the measured bf16 peaks (burst and sustained, MEASURED_PEAKS.json), one launch
          at a time on one stream ("latency") and 4 streams round-robin ("throughput").
  hbm   : standalone extract (NCHW fp32 / NHWC fp32), pack and decode kernels over batch sizes
          256..16384 -> GB/s of ALGORITHMIC bytes against the measured HBM peak.
Writes one JSON object per line to stdout; `--md FILE` also writes a markdown table.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

pkg = graft.load_package()
nb, synth = pkg.binding, pkg.synth


def time_launches(ctx, launch, reps, slots=1, warm=3):
    for i in range(warm):
        launch(i, i % slots)
    for s in range(slots):
        ctx.await_(s)
    e0 = nb.Event()
    e1 = [nb.Event() for _ in range(slots)]
    e0.record(ctx, 0)
    for i in range(reps):
        launch(warm + i, i % slots)
    for s in range(slots):
        e1[s].record(ctx, s)
    for s in range(slots):
        e1[s].sync()
    for s in range(slots):
        ctx.await_(s)
    return max(e0.elapsed_ms(e) for e in e1) / reps


def trunk_sweep(nets, batches, out):
    tf_burst, tf_sust, _, _ = bench.peaks()
    for C, blocks in nets:
        desc = nb.net_desc(C, blocks)
        blob = nb.random_blob(desc, 1234)
        bmax = max(batches)
        ctx = nb.Context(desc, batch_max=bmax, slots=4, blob=blob)
        pos = synth.random_positions(2048, seed=20240203)
        d_pos = nb.DeviceBuffer.from_host(pos)
        d_fbu = nb.DeviceBuffer(len(pos) * 86 * 16)
        ctx.pack_positions_device(0, d_pos.ptr, len(pos), d_fbu.ptr)
        ctx.await_(0)
        fbu = d_fbu.to_host((len(pos), 86), nb.FEATURE_BITBOARD)
        rng = np.random.default_rng(5)
        for B in batches:
            fb_bytes = B * 86 * 16
            pool = max(4, min(64, (bench.L2_BYTES // fb_bytes) + 2))
            host_pool = fbu[rng.integers(0, len(pos), size=pool * B)].reshape(pool, B * 86)
            d_pool = nb.DeviceBuffer.from_host(host_pool)
            off, idx = synth.random_legal_moves(B, seed=20240203, edge_rows=False)
            d_off, d_idx = nb.DeviceBuffer.from_host(off), nb.DeviceBuffer.from_host(idx)
            outs = [(nb.DeviceBuffer(int(off[-1]) * 4), nb.DeviceBuffer(B * 4), nb.DeviceBuffer(B * 4), nb.DeviceBuffer(B))
                    for _ in range(4)]

            def launch(i, slot):
                o = outs[slot]
                ctx.eval_decode_device(slot, d_pool.ptr + (i % pool) * fb_bytes, B, d_off.ptr, d_idx.ptr,
                                       nb.DECODE_PROBS, None, o[0].ptr, o[1].ptr, o[2].ptr, o[3].ptr)

            flops = bench.trunk_flops_per_sample(C, blocks) * B
            est_ms = flops / 1.0e15 * 1e3 + 0.02
            reps = int(max(8, min(400, 400.0 / est_ms)))
            ms1 = time_launches(ctx, launch, reps, slots=1)
            ms4 = time_launches(ctx, launch, reps, slots=4)
            row = {"kind": "trunk", "net": f"{blocks}x{C}", "batch": B, "reps": reps,
                   "latency_ms": round(ms1, 5), "evals_per_s_1stream": round(B / ms1 * 1e3, 1),
                   "tflops_1stream": round(flops / ms1 / 1e9, 1),
                   "ms_per_batch_4streams": round(ms4, 5), "evals_per_s_4streams": round(B / ms4 * 1e3, 1),
                   "tflops_4streams": round(flops / ms4 / 1e9, 1),
                   "frac_burst_peak": round(flops / ms4 / 1e9 / tf_burst, 4),
                   "frac_sustained_peak": round(flops / ms4 / 1e9 / tf_sust, 4)}
            out(row)
            for b in [d_pool, d_off, d_idx] + [x for o in outs for x in o]:
                b.free()
        ctx.close()


def hbm_sweep(batches, out):
    _, _, hbm, _ = bench.peaks()
    desc = nb.net_desc(128, 1)
    ctx = nb.Context(desc, batch_max=8, slots=1, seed=1)
    for B in batches:
        pos = synth.random_positions(min(B, 4096), seed=3)
        pos = np.resize(pos, B)
        # rotate over > 2 x L2 of inputs/outputs so that no launch finds its lines in the 126 MB L2
        nbuf = int(min(48, max(3, -(-300_000_000 // (B * 27864)))))
        d_pos = [nb.DeviceBuffer.from_host(pos) for _ in range(nbuf)]
        d_fb = [nb.DeviceBuffer(B * 86 * 16) for _ in range(nbuf)]
        for k in range(nbuf):
            ctx.pack_positions_device(0, d_pos[k].ptr, B, d_fb[k].ptr)
        ctx.await_(0)
        d_planes = [nb.DeviceBuffer(B * 86 * 81 * 4) for _ in range(nbuf)]
        reps = 30 if B >= 4096 else 100
        kernels = []
        kernels.append(("pack_positions_kernel", 108 + 1376,
                        lambda i, s: ctx.pack_positions_device(0, d_pos[i % nbuf].ptr, B, d_fb[i % nbuf].ptr)))
        kernels.append(("extract_nchw_kernel", 1376 + 27864,
                        lambda i, s: ctx.extract_device(0, d_fb[i % nbuf].ptr, B, 86, True, d_planes[i % nbuf].ptr)))
        kernels.append(("extract_nhwc_kernel", 1376 + 27864,
                        lambda i, s: ctx.extract_device(0, d_fb[i % nbuf].ptr, B, 86, False, d_planes[i % nbuf].ptr)))
        off, idx = synth.random_legal_moves(B, seed=9, edge_rows=False)
        n_moves = int(off[-1])
        logits = np.random.default_rng(1).normal(0, 3, size=(B, nb.POLICY_SIZE)).astype(np.float32)
        d_logits = [nb.DeviceBuffer.from_host(logits) for _ in range(nbuf)]
        d_w = nb.DeviceBuffer.from_host(np.full(B, 0.5, np.float32))
        d_off, d_idx = nb.DeviceBuffer.from_host(off), nb.DeviceBuffer.from_host(idx)
        d_legal, d_flag = nb.DeviceBuffer(max(n_moves, 1) * 4), nb.DeviceBuffer(B)
        decode_bytes = (2187 * 4 + 8) + (6 * n_moves + 8 * B) / B     # per position
        kernels.append(("decode_kernel", decode_bytes,
                        lambda i, s: ctx.decode_device(0, d_logits[i % nbuf].ptr, d_w.ptr, d_w.ptr, B, d_off.ptr, d_idx.ptr,
                                                       nb.DECODE_PROBS, d_legal.ptr, d_flag.ptr)))
        for name, bytes_per_pos, fn in kernels:
            ms = time_launches(ctx, fn, reps, slots=1)
            gbs = bytes_per_pos * B / ms / 1e6
            out({"kind": "hbm", "kernel": name, "batch": B, "reps": reps, "us_per_launch": round(ms * 1e3, 2),
                 "algorithmic_bytes": int(bytes_per_pos * B), "gb_per_s": round(gbs, 1), "frac_hbm_peak": round(gbs / hbm, 4),
                 "hbm_peak_gbs": hbm})
        for b in d_pos + d_fb + d_planes + d_logits + [d_w, d_off, d_idx, d_legal, d_flag]:
            b.free()
    ctx.close()


def cache_sweep(out):
    """Device-resident evaluation cache: probe / store kernels (GB/s of entry + row traffic) and the
    cached evaluation path (probe -> trunk on the misses -> fused store) at several hit rates."""
    _, _, hbm, _ = bench.peaks()
    B = 4096
    desc = nb.net_desc(128, 10)
    ctx = nb.Context(desc, batch_max=B, slots=4, seed=1234)
    ctx.cache_create(8192)                       # 8 GiB, like the reference's default (src/context.h:89)
    nbund = ctx.cache_num_bundles()
    rng = np.random.default_rng(3)
    off, idx = synth.random_legal_moves(B, seed=9, edge_rows=False)
    cnt = np.diff(off)
    n_moves = int(off[-1])
    pos = synth.random_positions(2048, seed=20240203)
    d_pos = nb.DeviceBuffer.from_host(pos)
    d_fbu = nb.DeviceBuffer(len(pos) * 86 * 16)
    ctx.pack_positions_device(0, d_pos.ptr, len(pos), d_fbu.ptr)
    ctx.await_(0)
    fbu = d_fbu.to_host((len(pos), 86), nb.FEATURE_BITBOARD)
    d_feat = nb.DeviceBuffer.from_host(fbu[rng.integers(0, len(pos), size=B)].reshape(-1))
    d_off, d_idx = nb.DeviceBuffer.from_host(off), nb.DeviceBuffer.from_host(idx)
    d_rows = nb.DeviceBuffer.from_host(rng.random(n_moves).astype(np.float32))
    d_win = nb.DeviceBuffer.from_host(np.full(B, 0.5, np.float32))
    d_legal, d_flag, d_hit = nb.DeviceBuffer(n_moves * 4), nb.DeviceBuffer(B), nb.DeviceBuffer(B)
    d_miss, d_cnt = nb.DeviceBuffer(4 * B), nb.DeviceBuffer(4)
    nsets = 24                                    # distinct key sets so that no launch re-touches warm lines
    keysets = [(rng.integers(1, 2**62, size=B, dtype=np.int64).astype(np.uint64)) for _ in range(nsets)]
    d_keys = [nb.DeviceBuffer.from_host(k) for k in keysets]
    cacheable = cnt <= 164
    row_bytes = float((cnt[cacheable] * 4).sum())
    # store: hash + 3 headers (24 B) + meta RMW (8 B) read, row read + row/entry written
    store_bytes = B * (8 + 72 + 8) + cacheable.sum() * 24 + 2 * row_bytes
    ms = time_launches(ctx, lambda i, s: ctx.cache_store_device(0, d_keys[i % nsets].ptr, B, d_off.ptr, d_rows.ptr, d_win.ptr,
                                                                d_win.ptr), nsets - 3, slots=1)
    out({"kind": "hbm", "kernel": "cache_store_kernel", "batch": B, "reps": nsets - 3, "us_per_launch": round(ms * 1e3, 2),
         "algorithmic_bytes": int(store_bytes), "gb_per_s": round(store_bytes / ms / 1e6, 1),
         "frac_hbm_peak": round(store_bytes / ms / 1e6 / hbm, 4), "hbm_peak_gbs": hbm})
    for k in range(nsets):                        # make sure every set is resident
        ctx.cache_store_device(0, d_keys[k].ptr, B, d_off.ptr, d_rows.ptr, d_win.ptr, d_win.ptr)
    ctx.await_(0)
    probe_bytes = B * (8 + 72 + 8 + 8) + 2 * row_bytes + cacheable.sum() * 8
    ms = time_launches(ctx, lambda i, s: ctx.cache_probe_device(0, d_keys[i % nsets].ptr, B, d_off.ptr, d_legal.ptr, d_win.ptr,
                                                                d_win.ptr, d_hit.ptr, d_miss.ptr, d_cnt.ptr), nsets - 3, slots=1)
    out({"kind": "hbm", "kernel": "cache_probe_kernel (all hits)", "batch": B, "reps": nsets - 3,
         "us_per_launch": round(ms * 1e3, 2), "algorithmic_bytes": int(probe_bytes),
         "gb_per_s": round(probe_bytes / ms / 1e6, 1), "frac_hbm_peak": round(probe_bytes / ms / 1e6 / hbm, 4),
         "hbm_peak_gbs": hbm})
    # cached evaluation at a given hit rate: a fraction of the batch reuses resident keys
    Bc = 256
    offc, idxc = synth.random_legal_moves(Bc, seed=20240203, edge_rows=False)
    d_offc, d_idxc = nb.DeviceBuffer.from_host(offc), nb.DeviceBuffer.from_host(idxc)
    d_rowsc = nb.DeviceBuffer.from_host(rng.random(int(offc[-1])).astype(np.float32))
    resident = rng.integers(1, 2**62, size=Bc, dtype=np.int64).astype(np.uint64)
    d_res = nb.DeviceBuffer.from_host(resident)
    ctx.cache_store_device(0, d_res.ptr, Bc, d_offc.ptr, d_rowsc.ptr, d_win.ptr, d_win.ptr)
    ctx.await_(0)
    outs = [(nb.DeviceBuffer(int(offc[-1]) * 4), nb.DeviceBuffer(Bc * 4), nb.DeviceBuffer(Bc * 4), nb.DeviceBuffer(Bc),
             nb.DeviceBuffer(Bc)) for _ in range(4)]
    reps = 400
    for rate in (0.0, 0.25, 0.5, 0.9):
        nhit = int(round(rate * Bc))
        keybufs = []
        for _ in range(reps + 8):                # fresh miss keys every launch (they get stored by the launch)
            k = rng.integers(1, 2**62, size=Bc, dtype=np.int64).astype(np.uint64)
            k[:nhit] = resident[:nhit]
            keybufs.append(nb.DeviceBuffer.from_host(k))

        def launch(i, slot):
            o = outs[slot]
            ctx.eval_cached_decode_device(slot, d_feat.ptr, Bc, keybufs[i].ptr, d_offc.ptr, d_idxc.ptr, nb.DECODE_PROBS,
                                          o[0].ptr, o[1].ptr, o[2].ptr, o[3].ptr, o[4].ptr)

        ms4 = time_launches(ctx, launch, reps, slots=4, warm=4)
        hits = int(outs[0][4].to_host((Bc,), np.uint8).sum())
        out({"kind": "cached_eval", "net": "10x128", "batch": Bc, "hit_rate_target": rate, "hits_in_last_batch": hits,
             "ms_per_batch_4streams": round(ms4, 5), "positions_per_s": round(Bc / ms4 * 1e3, 1)})
        for b in keybufs:
            b.free()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="all", choices=["all", "trunk", "hbm", "cache"])
    ap.add_argument("--nets", default="10x128,20x256,40x256")
    ap.add_argument("--batches", default="64,128,256,512,1024,2048,4096")
    ap.add_argument("--hbm-batches", default="256,1024,4096,16384")
    ap.add_argument("--md", default=None)
    args = ap.parse_args()
    rows = []

    def out(r):
        rows.append(r)
        print(json.dumps(r), flush=True)

    if args.what in ("all", "hbm"):
        hbm_sweep([int(x) for x in args.hbm_batches.split(",")], out)
    if args.what in ("all", "cache"):
        cache_sweep(out)
    if args.what in ("all", "trunk"):
        nets = [(int(s.split("x")[1]), int(s.split("x")[0])) for s in args.nets.split(",")]
        trunk_sweep(nets, [int(x) for x in args.batches.split(",")], out)
    if args.md:
        with open(args.md, "w") as f:
            f.write("# Roofline sweep (tools/sweep.py, B200, CUDA events)\n\n")
            f.write("## Trunk kernel (extract + forward + fused decode, one launch per batch)\n\n")
            f.write("| net | batch | latency ms (1 stream) | evals/s (1 stream) | evals/s (4 streams) | TFLOP/s (4 streams) | % burst peak | % sustained peak |\n|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                if r["kind"] == "trunk":
                    f.write(f"| {r['net']} | {r['batch']} | {r['latency_ms']:.4f} | {r['evals_per_s_1stream']:.0f} | "
                            f"{r['evals_per_s_4streams']:.0f} | {r['tflops_4streams']:.0f} | {100 * r['frac_burst_peak']:.1f} | "
                            f"{100 * r['frac_sustained_peak']:.1f} |\n")
            f.write("\n## HBM-bound stages (standalone kernels; algorithmic bytes)\n\n")
            f.write("| kernel | batch | us / launch | MB moved | GB/s | % of measured HBM peak |\n|---|---|---|---|---|---|\n")
            for r in rows:
                if r["kind"] == "hbm":
                    f.write(f"| {r['kernel']} | {r['batch']} | {r['us_per_launch']:.2f} | {r['algorithmic_bytes'] / 1e6:.2f} | "
                            f"{r['gb_per_s']:.0f} | {100 * r['frac_hbm_peak']:.1f} |\n")
            if any(r["kind"] == "cached_eval" for r in rows):
                f.write("\n## Cached evaluation (probe -> trunk on the misses -> fused store), 4 streams\n\n")
                f.write("| net | batch | target hit rate | hits in last batch | ms / batch | positions/s |\n|---|---|---|---|---|---|\n")
                for r in rows:
                    if r["kind"] == "cached_eval":
                        f.write(f"| {r['net']} | {r['batch']} | {r['hit_rate_target']:.2f} | {r['hits_in_last_batch']} | "
                                f"{r['ms_per_batch_4streams']:.4f} | {r['positions_per_s']:.0f} |\n")


if __name__ == "__main__":
    main()
