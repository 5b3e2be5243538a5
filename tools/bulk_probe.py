#!/usr/bin/env python
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
nb = graft.load_package().binding
for ctas in (1, 128, 148):
    for tile, stages, split in ((16384, 6, 1), (16384, 6, 2), (16384, 6, 4), (16384, 6, 8), (8192, 12, 1), (4096, 24, 1),
                                (16384, 3, 1), (16384, 10, 1), (32768, 6, 1)):
        print(f"ctas={ctas} tile={tile} stages={stages} split={split}: {nb.bulk_rate_probe(ctas, tile, stages, split):.1f} B/clk/SM", flush=True)
