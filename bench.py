#!/usr/bin/env python
"""bench.py — leaf evals/sec of the B200 leaf-evaluation path (BASELINE.json metric, config 2).

A step = one batch of 256 synthetic positions through the hot path: feature bitboards ->
(in-kernel plane expansion) -> 10 x 128 ResNet on tcgen05 -> policy head -> fused legal-move
gather + softmax + value/draw sigmoids.  One fused kernel launch per step.

  value : device-timed (CUDA events on the launch stream), inputs already resident in HBM,
          rotating over a pool of input batches larger than L2.
  e2e   : the same step through the host-buffer C-ABI call (nsb_eval_decode_async + nsb_await):
          pinned host inputs, H2D + kernel + D2H inside the timed region, several slots in flight.
  --impl reference : the reference's CPU path for its CPU-runnable config (pack + its own Random
          executor compiled from /root/reference into oracle/_ref + decode) on all host cores.
Multi-GPU: one process per GPU (torchrun), independent replicas, weak scaling; NCCL only for the
barrier, the max-over-ranks of the timed region and the final counter reduction.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

METRIC = "nn_leaf_evals_per_sec_batch256"
UNIT = "evals/s"
L2_BYTES = 126 * 1024 * 1024
FLOPS_PER_SAMPLE = {(128, 10): 0.4944e9, (256, 20): 3.8554e9, (256, 40): 7.6774e9}


def workload_name(C, blocks, B):
    return (f"{blocks}-block {C}-ch ResNet random-init, batch {B} leaf evaluation on 1xB200 per replica "
            "(feature planes + forward + policy decode)")


def trunk_flops_per_sample(C, blocks, in_ch=86):
    """2*MACs, unpadded channels: stem + 2*blocks convs + policy/value heads + FCs."""
    conv = lambda ci, co, k: 2.0 * 81 * k * ci * co
    f = conv(in_ch, C, 9) + 2 * blocks * conv(C, C, 9) + conv(C, 27, 1) + conv(C, 1, 1)
    f += 2.0 * 81 * 256 + 2.0 * 256 * 2
    return f


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        load = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def build_workload(nb, synth, orc_unused, B, pool, seed=20240203):
    """Pool of `pool` distinct input batches (feature bitboards via the product's own stage-1
    kernel would need a GPU; here they come from packed synthetic positions uploaded once)."""
    base_positions = synth.random_positions(min(pool * B, 4096), seed=seed)
    off1, idx1 = synth.random_legal_moves(B, seed=seed, edge_rows=False)
    return base_positions, off1, idx1


def run_b200(args):
    pkg = graft.load_package()
    nb, synth, rep = pkg.binding, pkg.synth, pkg.replica
    info = rep.RankInfo.from_env()
    world = info.world
    if world > 1 or args.gpus > 1:
        import torch
        import torch.distributed as dist

        if world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
        torch.cuda.set_device(info.local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", info.local_rank))
    if nb.device_count() < 1:
        raise SystemExit("bench: no CUDA device; the product path has no CPU fallback")
    gpu = info.local_rank
    B, C, blocks = args.batch, args.channels, args.blocks
    desc = nb.net_desc(C, blocks)
    blob = nb.random_blob(desc, 1234)
    slots = args.slots
    ctx = nb.Context(desc, batch_max=B, slots=slots, gpu=gpu, blob=blob)

    # ---- synthetic inputs: pool of batches > L2 so no step re-reads a cached input --------------
    fb_bytes = B * 86 * 16
    pool = max(4, (L2_BYTES + fb_bytes - 1) // fb_bytes + 8) if not args.small_pool else 8
    pos = synth.random_positions(2048, seed=20240203 + info.rank)
    d_pos = nb.DeviceBuffer.from_host(pos)
    d_fb_unique = nb.DeviceBuffer(len(pos) * 86 * 16)
    ctx.pack_positions_device(0, d_pos.ptr, len(pos), d_fb_unique.ptr)   # product stage-1 kernel
    ctx.await_(0)
    fb_unique = d_fb_unique.to_host((len(pos), 86), nb.FEATURE_BITBOARD)
    rng = np.random.default_rng(99 + info.rank)
    off, idx = synth.random_legal_moves(B, seed=20240203, edge_rows=False)
    n_moves = int(off[-1])
    # device pool
    host_pool = fb_unique[rng.integers(0, len(pos), size=pool * B)].reshape(pool, B * 86)
    d_pool = nb.DeviceBuffer.from_host(host_pool)
    d_off = nb.DeviceBuffer.from_host(off)
    d_idx = nb.DeviceBuffer.from_host(idx)
    n_out = 8
    d_legal = [nb.DeviceBuffer(max(n_moves, 1) * 4) for _ in range(n_out)]
    d_win = [nb.DeviceBuffer(B * 4) for _ in range(n_out)]
    d_draw = [nb.DeviceBuffer(B * 4) for _ in range(n_out)]
    d_flag = [nb.DeviceBuffer(B) for _ in range(n_out)]

    def dev_step(i, slot=0):
        j = i % n_out
        # the production path: legal-move rows out, dense logits never leave the SM (d_policy = NULL)
        ctx.eval_decode_device(slot, d_pool.ptr + (i % pool) * fb_bytes, B, d_off.ptr, d_idx.ptr, nb.DECODE_PROBS,
                               None, d_legal[j].ptr, d_win[j].ptr, d_draw[j].ptr, d_flag[j].ptr)

    def await_all():
        for s_ in range(slots):
            ctx.await_(s_)

    K, W = args.steps, max(args.warmup, 3)
    S = slots                       # launches round-robin over S streams (n_out % S == 0: an output
    assert n_out % S == 0           # buffer is only ever reused on the stream that wrote it last)
    for i in range(W):
        dev_step(i, i % S)
    await_all()
    sampler = ClockSampler(gpu)
    sampler.start()
    time.sleep(0.3)
    rep.barrier()
    nb.device_sync()
    l0 = ctx.launch_count()
    e0 = nb.Event()
    e_end = [nb.Event() for _ in range(S)]
    e0.record(ctx, 0)
    for i in range(K):
        dev_step(W + i, i % S)
    for s_ in range(S):
        e_end[s_].record(ctx, s_)
    for s_ in range(S):
        e_end[s_].sync()
    await_all()
    nb.device_sync()
    rep.barrier()
    elapsed_ms = max(e0.elapsed_ms(e) for e in e_end)
    launches = ctx.launch_count() - l0

    # per-launch duration of the trunk kernel: the same launches issued back to back on ONE stream
    # (no overlap between launches), a CUDA event pair around each, harvested in nsb_await
    K2 = min(K, 500)
    ctx.set_timing(True)
    ctx.trunk_time_reset()
    es0, es1 = nb.Event(), nb.Event()
    es0.record(ctx, 0)
    for i in range(K2):
        dev_step(W + K + i, 0)
    es1.record(ctx, 0)
    es1.sync()
    ctx.await_(0)
    seq_elapsed_ms = es0.elapsed_ms(es1)
    trunk_ms_sum, trunk_n = ctx.trunk_time()
    ctx.set_timing(False)
    dev_device = None
    if world > 1:
        import torch
        dev_device = torch.device("cuda", info.local_rank)
    counters, elapsed_max = rep.aggregate({"evals": B * K, "batches": K, "legal_moves": n_moves * K}, elapsed_ms,
                                          device=dev_device)
    value = rep.whole_job_rate(counters["evals"], elapsed_max)

    # ---- e2e: host buffers through the C ABI, `slots` batches in flight ---------------------------------
    h_pool_n = 16
    h_fb = [nb.PinnedArray((B * 86,), nb.FEATURE_BITBOARD) for _ in range(h_pool_n)]
    for k, a in enumerate(h_fb):
        a.array[:] = host_pool[k % pool]
    h_off = nb.PinnedArray((B + 1,), np.uint32); h_off.array[:] = off
    h_idx = nb.PinnedArray((max(n_moves, 1),), np.uint16); h_idx.array[:n_moves] = idx
    h_legal = [nb.PinnedArray((max(n_moves, 1),), np.float32) for _ in range(slots)]
    h_win = [nb.PinnedArray((B,), np.float32) for _ in range(slots)]
    h_draw = [nb.PinnedArray((B,), np.float32) for _ in range(slots)]
    h_flag = [nb.PinnedArray((B,), np.uint8) for _ in range(slots)]
    h_policy = [nb.PinnedArray((B * 2187,), np.float32) for _ in range(slots)]
    sink = 0.0

    def e2e_loop(steps, fused):
        nonlocal sink
        for i in range(steps):
            s = i % slots
            if i >= slots:
                ctx.await_(s)
                sink += float(h_win[s].array[0])            # consume the step's result on the host
            if fused:
                ctx.eval_decode_async(s, h_fb[i % h_pool_n].array, B, h_off.array, h_idx.array, nb.DECODE_PROBS,
                                      h_legal[s].array, h_win[s].array, h_draw[s].array, h_flag[s].array)
            else:
                ctx.eval_async(s, h_fb[i % h_pool_n].array, B, h_policy[s].array, h_win[s].array, h_draw[s].array)
        for s in range(slots):
            ctx.await_(s)
            sink += float(h_win[s].array[0])

    def time_e2e(fused):
        e2e_loop(max(W, slots), fused)
        rep.barrier()
        nb.device_sync()
        t0 = time.perf_counter()
        e2e_loop(K, fused)
        nb.device_sync()
        dt = (time.perf_counter() - t0) * 1e3
        rep.barrier()
        _, dt_max = rep.aggregate({}, dt, device=dev_device)
        return rep.whole_job_rate(B * K * world, dt_max)

    e2e_fused = time_e2e(True)
    e2e_infer = time_e2e(False)

    # packed positions in (108 B instead of 1,376 B per position over PCIe): stage 1 runs in the trunk
    # kernel's prologue, still one launch per batch
    h_pos = [nb.PinnedArray((B,), nb.POSITION) for _ in range(h_pool_n)]
    for k, a_ in enumerate(h_pos):
        a_.array[:] = pos[rng.integers(0, len(pos), size=B)]

    def e2e_positions_loop(steps):
        nonlocal sink
        for i in range(steps):
            s = i % slots
            if i >= slots:
                ctx.await_(s)
                sink += float(h_win[s].array[0])
            ctx.eval_positions_decode_async(s, h_pos[i % h_pool_n].array, B, h_off.array, h_idx.array, nb.DECODE_PROBS,
                                            h_legal[s].array, h_win[s].array, h_draw[s].array, h_flag[s].array)
        for s in range(slots):
            ctx.await_(s)
            sink += float(h_win[s].array[0])

    def timed(loop, steps):
        rep.barrier()
        nb.device_sync()
        t0 = time.perf_counter()
        loop(steps)
        nb.device_sync()
        dt = (time.perf_counter() - t0) * 1e3
        rep.barrier()
        _, dt_max = rep.aggregate({}, dt, device=dev_device)
        return dt_max

    # By now the board sits at its power cap and clocks have dropped since the first leg, so the positions-in leg
    # is timed interleaved with the bitboards-in call in the same window (two halves each): that pair compares.
    e2e_positions_loop(max(W, slots))
    lp0 = ctx.launch_count()
    half = max(K // 2, slots)
    ms_pos = ms_bb = 0.0
    for _ in range(2):
        ms_pos += timed(e2e_positions_loop, half)
        ms_bb += timed(lambda n_: e2e_loop(n_, True), half)
    pos_launches = (ctx.launch_count() - lp0) / 2          # both legs launch one kernel per step
    e2e_positions = rep.whole_job_rate(B * 2 * half * world, ms_pos)
    e2e_bb_same_window = rep.whole_job_rate(B * 2 * half * world, ms_bb)
    K_pos = 2 * half

    # one batch at a time (the reference executor's usage: computeNonBlocking -> await per evaluator thread,
    # src/mcts/evaluationworker.cc:158-180): a one-slot context = classic kernel + direct I/O (the kernel
    # reads / writes the page-locked host buffers itself), against the same context with staged copies
    latency = None
    if not args.no_latency_leg:
        ctx1 = nb.Context(desc, batch_max=B, slots=1, gpu=gpu, blob=blob)
        KL = min(K, 500)

        def one_at_a_time(steps):
            nonlocal sink
            t0 = time.perf_counter()
            for i in range(steps):
                ctx1.eval_decode_async(0, h_fb[i % h_pool_n].array, B, h_off.array, h_idx.array, nb.DECODE_PROBS,
                                       h_legal[0].array, h_win[0].array, h_draw[0].array, h_flag[0].array)
                ctx1.await_(0)
                sink += float(h_win[0].array[0])
            return (time.perf_counter() - t0) * 1e6 / steps

        latency = {"kernel": ctx1.trunk_kernel_name(), "batch": B, "steps": KL,
                   "api": "nsb_eval_decode_async + nsb_await, one batch in flight, host buffers"}
        for mode in ("staged", "direct"):
            ctx1.set_io_mode(mode == "direct")
            one_at_a_time(50)
            latency[f"{mode}_us_per_batch"] = round(one_at_a_time(KL), 2)
        latency["default_io"] = "direct"
        ctx1.close()
    clocks = sampler.stop()
    selfplay = None if args.no_selfplay else selfplay_leg(args, info, rep, dev_device)

    # ---- roofline of the dominant (only) kernel ----------------------------------------------------------
    tf_burst, tf_sust, hbm, which = peaks()
    flops_launch = trunk_flops_per_sample(C, blocks) * B
    avg_launch_ms = trunk_ms_sum / max(trunk_n, 1)
    achieved = flops_launch / (avg_launch_ms * 1e-3) / 1e12 if avg_launch_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "trunk_traffic.json")
    if os.path.exists(tp):
        kname = ctx.trunk_kernel_name()
        tag = ":duo" if "duo" in kname else ""
        traffic = json.load(open(tp)).get(f"{C}x{blocks}@{B}{tag}")
    # `achieved`: the timed region itself.  Its launches overlap (`slots` streams), so a launch's duration
    # there is the steady-state time per launch, elapsed / launches; the same kernel launched alone
    # (one stream, back to back, an event pair per launch) is reported beside it.
    step_ms = elapsed_max / K
    achieved_ss = flops_launch / (step_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": ctx.trunk_kernel_name(), "achieved": round(achieved_ss, 2),
                "peak": tf_burst, "unit": "TFLOP/s", "frac": round(achieved_ss / tf_burst, 4),
                "frac_of_sustained_peak": round(achieved_ss / tf_sust, 4), "peak_source": which,
                "flops_per_launch": flops_launch, "avg_launch_ms": round(step_ms, 5),
                "timing": f"timed region: {K} launches over {slots} streams, CUDA events on the streams, "
                          "duration per launch = elapsed / launches (launches overlap); per GPU",
                "isolated_launch": {"avg_launch_ms": round(avg_launch_ms, 5), "achieved": round(achieved, 2),
                                    "frac": round(achieved / tf_burst, 4),
                                    "frac_of_sustained_peak": round(achieved / tf_sust, 4),
                                    "kernel_share_of_step": round(trunk_ms_sum / seq_elapsed_ms, 4) if seq_elapsed_ms > 0 else None,
                                    "timing": f"{trunk_n} launches back to back on one stream, one CUDA event pair per "
                                              f"launch; that sub-run: {seq_elapsed_ms / max(K2, 1):.5f} ms/step"},
                "traffic": traffic}

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(elapsed_max / K, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(C, blocks, B), "launches_per_step": 1,
                   "batch": B, "channels": C, "blocks": blocks, "legal_moves_per_batch": n_moves,
                   "l2_policy": f"inputs rotate over a pool of {pool} batches ({pool * fb_bytes >> 20} MiB > L2); "
                                "weights stay L2-resident as in steady-state serving",
                   "replicas": world, "streams_in_flight": slots,
                   "value_leg": f"{slots} streams round-robin, device-resident inputs/outputs, events on the streams"},
        "e2e": {"value": round(e2e_fused, 1), "unit": UNIT, "api": "nsb_eval_decode_async + nsb_await (host buffers)",
                "h2d_bytes_per_step": fb_bytes + (B + 1) * 4 + n_moves * 2,
                "d2h_bytes_per_step": n_moves * 4 + B * 4 * 2 + B, "clock": "host, device-synchronised both sides"},
        "e2e_infer_contract": {"value": round(e2e_infer, 1), "unit": UNIT,
                               "api": "nsb_eval_async + nsb_await == Infer::computeNonBlocking/await (dense logits)",
                               "h2d_bytes_per_step": fb_bytes, "d2h_bytes_per_step": B * 2187 * 4 + B * 8},
        "e2e_positions": {"value": round(e2e_positions, 1), "unit": UNIT,
                          "api": "nsb_eval_positions_decode_async + nsb_await (packed positions in, stage 1 in the trunk prologue)",
                          "h2d_bytes_per_step": B * 108 + (B + 1) * 4 + n_moves * 2,
                          "d2h_bytes_per_step": n_moves * 4 + B * 4 * 2 + B, "launches_per_step": pos_launches / K_pos,
                          "bitboards_in_same_window": round(e2e_bb_same_window, 1),
                          "note": "timed in alternating half-windows with the bitboards-in call, late in the run "
                                  "(board at its power cap): compare with bitboards_in_same_window, not with e2e"},
        "latency_one_batch_in_flight": latency,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "counters": counters,
    }
    if selfplay is not None:
        line["selfplay"] = selfplay
    if info.rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(synth, B, seconds=args.cpu_seconds)
    if info.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return sink


def selfplay_leg(args, info, rep, device):
    """BASELINE.json's second metric, "self-play positions/sec" (configs[3]: 20x256 net, 1024 concurrent
    games per GPU): the C++ self-play loop of nshogi-engine_b200/host/selfplay_sim.cc (frame pool, search
    workers, pinned multi-slot evaluation worker) against infer::B200, one process per GPU, counters
    summed over ranks.  Rules are synthetic (libnshogi is not available); the GPU path is the real one."""
    exe = os.path.join(ROOT, "nshogi-engine_b200", "host", "nsb_selfplay_sim")
    if not os.path.exists(exe):
        return {"unavailable": "nsb_selfplay_sim not built"}
    workers = max(2, min(8, (os.cpu_count() or 4) // max(1, info.world) - 1))
    cmd = [exe, "--gpu", str(info.local_rank), "--channels", "256", "--blocks", "20", "--batch-size", "512",
           "--frame-pool-size", "1024", "--num-search-workers", str(workers), "--seconds", str(args.selfplay_seconds),
           "--warmup", "1.5"]
    rep.barrier()
    rec, err = None, None
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=120 + args.selfplay_seconds)
        if out.returncode == 0:
            rec = json.loads(out.stdout.strip().splitlines()[-1])
        else:
            err = out.stderr.strip()[-300:] or f"exit code {out.returncode}"
    except (OSError, subprocess.SubprocessError, ValueError, IndexError) as e:
        err = repr(e)
    # every rank takes part in the reduction, also one whose harness failed (it contributes zeros)
    mine = {"records": rec["records"], "games": rec["games"], "evals": rec["evals"], "batches": rec["batches"],
            "nan_rows": 0 if rec else 1} if rec else {"nan_rows": 1}
    counters, ms_max = rep.aggregate(mine, rec["seconds"] * 1e3 if rec else 0.0, device=device)
    if counters["nan_rows"] or rec is None:      # "nan_rows" doubles as the count of failed ranks in this reduction
        return {"unavailable": f"self-play harness failed on {counters['nan_rows']} rank(s): {err}"}
    return {"metric": "selfplay_positions_per_sec", "value": round(rep.whole_job_rate(counters["records"], ms_max), 1),
            "unit": "positions/s", "leaf_evals_per_sec": round(rep.whole_job_rate(counters["evals"], ms_max), 1),
            "games_per_sec": round(rep.whole_job_rate(counters["games"], ms_max), 2),
            "avg_batch": round(counters["evals"] / max(counters["batches"], 1), 1), "seconds": rec["seconds"],
            "config": {"workload": "self-play data generation, 20x256 ResNet, 1024 concurrent games per GPU",
                       "batch_size": 512, "num_playouts": rec["num_playouts"], "full_search_ratio": rec["full_search_ratio"],
                       "search_workers_per_gpu": workers, "slots": rec["slots"], "rules": rec["rules"]}}


def cpu_baseline(synth, B, seconds=12.0, threads=None):
    """Reference's CPU path for its CPU-runnable config (BASELINE.json configs[0]): stage-1 pack
    (port; libnshogi absent) + Random executor (the reference's own random.cc compiled into
    oracle/_ref when present, else the port) + decode (port), one batch of B per thread-iteration."""
    orc = graft.load_oracle()
    threads = threads or (os.cpu_count() or 1)
    pos = synth.random_positions(1024, seed=20240203)
    off, idx = synth.random_legal_moves(len(pos), seed=20240203, edge_rows=False)
    use_ref = orc.have_ref_random()
    v, sec = orc.cpu_path(pos, off, idx, B, threads, 2, False, use_ref)       # calibrate
    per_thread = max(2, int(seconds / max(sec / 2, 1e-6)))
    v, sec = orc.cpu_path(pos, off, idx, B, threads, per_thread, False, use_ref)
    # context: the same net's fp32 forward on the host cores (oracle port, vectorised C, all threads)
    pkg = graft.load_package()
    desc = pkg.binding.net_desc(128, 10)
    blob = pkg.binding.random_blob(desc, 1234)
    planes = orc.expand(orc.pack(pos[:256]), 256)
    t0 = time.perf_counter()
    orc.forward(desc, blob, planes, False)
    fwd = 256 / (time.perf_counter() - t0)
    return {"value": round(v, 1), "with_oracle_forward_10x128": {"value": round(fwd, 1), "unit": UNIT,
            "sample": "256 positions through the oracle's fp32 CPU forward of the 10x128 net, all threads"}, "unit": UNIT, "cores": threads, "kind": "reference" if use_ref else "port",
            "sample": f"{threads} threads x {per_thread} batches of {B}: pack (port of FeatureStackComptime, "
                      f"builder-defined) + Random executor ({'reference src/infer/random.cc compiled in place' if use_ref else 'port of random.cc'}) "
                      f"+ decode (port of feedworker.cc:100-136); {sec:.1f} s; NN forward NOT included "
                      "(the reference's CPU config replaces it with the Random executor)"}


def run_reference(args):
    """--impl reference: the reference's CPU path on the host cores (see cpu_baseline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = graft.load_package()
    synth = pkg.synth
    orc = graft.load_oracle()
    B = args.batch
    threads = os.cpu_count() or 1
    pos = synth.random_positions(1024, seed=20240203)
    off, idx = synth.random_legal_moves(len(pos), seed=20240203, edge_rows=False)
    use_ref = orc.have_ref_random()
    v, sec = orc.cpu_path(pos, off, idx, B, threads, 2, False, use_ref)
    K, W = args.steps, max(args.warmup, 1)
    budget = 120.0                                   # whole run bounded to ~2 minutes
    per_step = max(1, int((budget / (K + W)) / max(sec / 2, 1e-6)))
    per_step = min(per_step, 64)
    for _ in range(W):
        orc.cpu_path(pos, off, idx, B, threads, per_step, False, use_ref)
    t0 = time.perf_counter()
    total = 0
    for _ in range(K):
        orc.cpu_path(pos, off, idx, B, threads, per_step, False, use_ref)
        total += threads * per_step * B
    dt = time.perf_counter() - t0
    value = total / dt
    sample = (f"each step = {threads} threads x {per_step} batches of {B} through pack (port) + Random executor "
              f"({'reference random.cc compiled in place' if use_ref else 'port'}) + decode (port); no NN forward on CPU")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": round(dt / K * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.channels, args.blocks, B), "batch": B,
                       "channels": args.channels, "blocks": args.blocks,
                       "reference_arm": "the reference has no CPU forward: its CPU-runnable executor (Random, "
                                        "BASELINE.json configs[0]) stands in for the ResNet"},
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": threads,
                             "kind": "reference" if use_ref else "port", "sample": sample},
            "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=10)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-selfplay", action="store_true")
    ap.add_argument("--no-latency-leg", action="store_true")
    ap.add_argument("--selfplay-seconds", type=float, default=6.0)
    ap.add_argument("--small-pool", action="store_true", help="8-batch input pool (profiling runs only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
