#!/usr/bin/env python
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
nb = graft.load_package().binding
for layout in (0, 1):
    for N in (96, 192, 256):
        for shift in (0, 1, 4, 8, 11):
            try:
                err, cyc = nb.umma_probe(N, 128, shift, layout)
                print(f"layout={'none' if layout == 0 else 'sw128'} N={N} shift={shift}: max_err={err} cycles/MMA={cyc:.1f}", flush=True)
            except Exception as e:
                print("ERR", layout, N, shift, e, flush=True)
                sys.exit(1)
