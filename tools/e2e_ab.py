#!/usr/bin/env python
"""A/B of the two host-buffer entry points, alternating so that clock / power drift hits both alike:
   bb  : nsb_eval_decode_async            (1,376 B per position over PCIe, bitboards in)
   pos : nsb_eval_positions_decode_async  (108 B per position, stage 1 in the trunk prologue)
Each is run with staged I/O (copy nodes around the kernel) and direct I/O (the kernel reads / writes the mapped
page-locked buffers itself).
usage: e2e_ab.py [channels blocks batch slots steps rounds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nb, synth = pkg.binding, pkg.synth
argv = [int(x) for x in sys.argv[1:]]
C, blocks, B, slots, K, rounds = (argv + [128, 10, 256, 4, 2000, 3][len(argv):])[:6]
desc = nb.net_desc(C, blocks)
ctx = nb.Context(desc, batch_max=B, slots=slots, seed=1234)
pos = synth.random_positions(2048, seed=1)
d_pos = nb.DeviceBuffer.from_host(pos)
d_fb = nb.DeviceBuffer(len(pos) * 86 * 16)
ctx.pack_positions_device(0, d_pos.ptr, len(pos), d_fb.ptr)
ctx.await_(0)
fb_unique = d_fb.to_host((len(pos), 86), nb.FEATURE_BITBOARD)
rng = np.random.default_rng(0)
off, idx = synth.random_legal_moves(B, seed=20240203, edge_rows=False)
n_moves = int(off[-1])
NP = 16
h_fb, h_pos = [], []
for k in range(NP):
    sel = rng.integers(0, len(pos), size=B)
    a = nb.PinnedArray((B * 86,), nb.FEATURE_BITBOARD); a.array[:] = fb_unique[sel].reshape(-1); h_fb.append(a)
    p = nb.PinnedArray((B,), nb.POSITION); p.array[:] = pos[sel]; h_pos.append(p)
h_off = nb.PinnedArray((B + 1,), np.uint32); h_off.array[:] = off
h_idx = nb.PinnedArray((n_moves,), np.uint16); h_idx.array[:] = idx
h_legal = [nb.PinnedArray((n_moves,), np.float32) for _ in range(slots)]
h_win = [nb.PinnedArray((B,), np.float32) for _ in range(slots)]
h_draw = [nb.PinnedArray((B,), np.float32) for _ in range(slots)]
h_flag = [nb.PinnedArray((B,), np.uint8) for _ in range(slots)]
RANKED = os.environ.get("E2E_RANKED") == "1"     # also ask for every row's rank order (Node::sort on the GPU)
h_order = [nb.PinnedArray((n_moves,), np.uint16) for _ in range(slots)]


def loop(steps, positions):
    for i in range(steps):
        s = i % slots
        if i >= slots:
            ctx.await_(s)
        if RANKED:
            ctx.eval_request_async(s, B, h_off.array, h_idx.array, nb.DECODE_PROBS, h_legal[s].array, h_win[s].array,
                                   h_draw[s].array, positions=h_pos[i % NP].array if positions else None,
                                   features=None if positions else h_fb[i % NP].array, order_out=h_order[s].array,
                                   nan_flag=h_flag[s].array)
        elif positions:
            ctx.eval_positions_decode_async(s, h_pos[i % NP].array, B, h_off.array, h_idx.array, nb.DECODE_PROBS,
                                            h_legal[s].array, h_win[s].array, h_draw[s].array, h_flag[s].array)
        else:
            ctx.eval_decode_async(s, h_fb[i % NP].array, B, h_off.array, h_idx.array, nb.DECODE_PROBS,
                                  h_legal[s].array, h_win[s].array, h_draw[s].array, h_flag[s].array)
    for s in range(slots):
        ctx.await_(s)


print(f"{ctx.trunk_kernel_name()}  B={B} slots={slots} ranked={RANKED} fuse_pack={os.environ.get('NSB_FUSE_PACK', '1')}")
for direct in (False, True):
    ctx.set_io_mode(direct)
    loop(100, False)
    loop(100, True)
for r in range(rounds):
    for direct in (False, True):
        ctx.set_io_mode(direct)
        for positions in (False, True):
            nb.device_sync()
            t0 = time.perf_counter()
            loop(K, positions)
            nb.device_sync()
            dt = time.perf_counter() - t0
            print(f"round {r} {ctx.io_mode():6s} {'pos' if positions else 'bb '}: {B * K / dt / 1e6:.3f} M evals/s  "
                  f"({dt / K * 1e6:.1f} us/step)")
