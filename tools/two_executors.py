#!/usr/bin/env python
"""The reference's default search configuration on one GPU (src/context.h:74-79): NumEvaluationThreadsPerGPU = 2
executors, BatchSize = 128, each thread one batch at a time (computeNonBlocking -> await).  Here: E one-slot
contexts (classic kernel, direct host I/O), E host threads, fused decode with rank order; per-executor step time
and total evals/s.  usage: two_executors.py [executors batch channels blocks steps]"""
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
argv = [int(x) for x in sys.argv[1:]]
E, B, C, blocks, K = (argv + [2, 128, 128, 10, 1500][len(argv):])[:5]
desc = nb.net_desc(C, blocks)
blob = nb.random_blob(desc, 1234)
off, idx = synth.random_legal_moves(B, seed=20240203, edge_rows=False)
nm = int(off[-1])
P = nb.PinnedArray


class Worker:
    def __init__(self, k):
        self.ctx = nb.Context(desc, batch_max=B, slots=1, blob=blob)
        self.pos = P((B,), nb.POSITION)
        self.pos.array[:] = synth.random_positions(B, seed=10 + k)
        self.off, self.idx = P((B + 1,), np.uint32), P((nm,), np.uint16)
        self.off.array[:], self.idx.array[:] = off, idx
        self.legal, self.order = P((nm,), np.float32), P((nm,), np.uint16)
        self.win, self.draw, self.flag = P((B,), np.float32), P((B,), np.float32), P((B,), np.uint8)
        self.us = 0.0

    def step(self):
        self.ctx.eval_request_async(0, B, self.off.array, self.idx.array, nb.DECODE_PROBS, self.legal.array, self.win.array,
                                    self.draw.array, positions=self.pos.array, order_out=self.order.array,
                                    nan_flag=self.flag.array)
        self.ctx.await_(0)

    def run(self, steps, barrier):
        for _ in range(50):
            self.step()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        self.us = (time.perf_counter() - t0) / steps * 1e6


for n_exec in sorted({1, E}):
    ws = [Worker(k) for k in range(n_exec)]
    bar = threading.Barrier(n_exec)
    ts = [threading.Thread(target=w.run, args=(K, bar)) for w in ws]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    total = sum(B / (w.us * 1e-6) for w in ws)
    print(f"{n_exec} executor(s) x B={B} ({blocks}x{C}, {ws[0].ctx.trunk_kernel_name()}, io={ws[0].ctx.io_mode()}): "
          f"{', '.join(f'{w.us:.1f}' for w in ws)} us per batch -> {total / 1e6:.3f} M evals/s")
    for w in ws:
        w.ctx.close()
