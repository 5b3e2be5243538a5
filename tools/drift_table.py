#!/usr/bin/env python
"""Prints the drift table of DESIGN.md §4 (GPU needed): for the nets of BASELINE.json's configs, the executor's
logits / decoded probabilities / win / draw against the oracle at both precisions (tests/helpers.drift_metrics).
tests/test_depth_parity.py asserts the same numbers against the stated tolerances; this tool is for reading them."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
import helpers  # noqa: E402

pkg = graft.load_package()
nb, synth = pkg.binding, pkg.synth
orc = graft.load_oracle()

CASES = [(128, 2, 1234, 32, "classic", 1), (128, 10, 1, 64, "classic", 1), (128, 10, 1, 64, "duo", 2), (256, 10, 1234, 32, None, 1),
         (256, 20, 1234, 32, None, 1), (256, 40, 5, 16, None, 1)]
rows = {}
for C, blocks, seed, n, kernel, slots in CASES:
    if kernel:
        os.environ["NSB_TRUNK128"] = kernel
    else:
        os.environ.pop("NSB_TRUNK128", None)
    desc = nb.net_desc(C, blocks)
    blob = nb.random_blob(desc, seed)
    pos = synth.random_positions(n, seed=2024)
    fb = orc.pack(pos)
    off, idx = synth.random_legal_moves(n, seed=3, edge_rows=False)
    policy = np.zeros((n, nb.POLICY_SIZE), dtype=np.float32)
    win, draw = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    with nb.Context(desc, batch_max=n, slots=slots, blob=blob) as ctx:
        ctx.eval_async(0, fb, n, policy, win, draw)
        ctx.await_(0)
    rows[f"{blocks}x{C}" + (f":{kernel}" if kernel else "")] = helpers.drift_metrics(nb, orc, desc, blob, fb, n, (policy, win, draw), off, idx)
keys = ["layers", "n", "logit_rms_fp32", "logit_vs_bf16", "logit_vs_bf16_mean", "value_vs_bf16", "logit_vs_fp32", "prob_vs_fp32",
        "kl_vs_fp32", "win_vs_fp32", "draw_vs_fp32", "oracle_prob_bf16_vs_fp32", "oracle_value_bf16_vs_fp32", "win_spread_fp32"]
print("| net | " + " | ".join(keys) + " |")
print("|---|" + "---|" * len(keys))
for name, m in rows.items():
    print(f"| {name} | " + " | ".join(f"{m[k]:.3g}" if isinstance(m[k], float) else str(m[k]) for k in keys) + " |")
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
