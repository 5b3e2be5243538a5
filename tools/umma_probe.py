#!/usr/bin/env python
"""Issue-to-retire rate of tcgen05.mma for the operand layouts the trunk kernels use (diagnostics)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
nb = graft.load_package().binding
names = {0: "cta1/none", 1: "cta1/sw128", 2: "pair/none", 3: "pair/sw128", 4: "cta1/A-in-TMEM", 5: "cta1/A-in-TMEM + smem stores"}
layouts = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(names)
for layout in layouts:
    for N in (96, 192, 256):
        if layout in (2, 3) and N == 96:
            N = 128
        for shift in (0, 11):
            try:
                err, cyc = nb.umma_probe(N, 128, shift, layout)
                print(f"layout={names[layout]} N={N} shift={shift}: max_err={err} cycles/MMA={cyc:.1f}", flush=True)
            except Exception as e:
                print("ERR", layout, N, shift, e, flush=True)
                sys.exit(1)
