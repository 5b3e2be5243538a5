#!/usr/bin/env python
"""Bulk-copy (UBLKCP) L2 -> shared-memory ring probe: bytes/clk/SM for a ring of `stages` tiles.
With stages = 1 the loop is fully serialised, so tile / rate = the round trip
(empty -> issue -> landed -> consumer -> empty) of one copy."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
nb = graft.load_package().binding
for ctas in (1, 148):
    for tile, stages, split in ((16384, 1, 1), (16384, 2, 1), (16384, 3, 1), (16384, 4, 1), (16384, 6, 1), (8192, 1, 1),
                                (8192, 2, 1), (8192, 4, 1), (8192, 12, 1), (4096, 1, 1), (32768, 1, 1), (32768, 2, 1),
                                (32768, 6, 1), (16384, 6, 2), (16384, 6, 4)):
        r = nb.bulk_rate_probe(ctas, tile, stages, split)
        print(f"ctas={ctas} tile={tile} stages={stages} split={split}: {r:.1f} B/clk/SM  ({tile / r:.0f} cycles per tile)", flush=True)
