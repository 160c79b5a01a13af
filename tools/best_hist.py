#!/usr/bin/env python
"""Distribution of the final best[] of a workload and what it implies for the band width
(window words per column) of the symmetric pair pass.  Needs a GPU.

    python tools/best_hist.py c2 1.0
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isocon_b200 import _binding, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
S = workloads.CONFIGS[name](scale=scale)
by_seq = {}
for a, s in S.items():
    by_seq[s] = a
L = sorted(by_seq.items(), key=lambda e: len(e[0]))
ctx = _binding.get_context(0)
ctx.set_reads([s for s, _ in L])
n = len(L)
best, q, t, d = ctx.graph(1, 2 ** 32, np.ones(n, np.uint8), None)
lens = np.array([len(s) for s, _ in L])
print("reads", n, "len min/median/max", lens.min(), int(np.median(lens)), lens.max())
print("best percentiles 1/5/25/50/75/95/99/100:", np.percentile(best, [1, 5, 25, 50, 75, 95, 99, 100]).astype(int).tolist())
for cap in (63, 95, 127, 159, 191, 223, 255):
    print("  best > %d: %.2f %%" % (cap, 100.0 * (best > cap).mean()))
sb = np.sort(best)
# pair threshold with the final best: max(b_q, b_t); W = ceil((k + 1) / 32)
W = (sb + 32) // 32
cnt = np.arange(n)            # a read with rank r in sorted order is the max of r pairs
tot = cnt.sum()
print("mean window words per pair with final thresholds: %.3f" % ((W * cnt).sum() / tot))
for cap in (95, 127, 159):
    Wc = (np.minimum(sb, cap) + 32) // 32
    unres = (best > cap).sum()
    print("  cap %d: %.3f words/pair + %d unresolved rows" % (cap, (Wc * cnt).sum() / tot, unres))
print(ctx.stats())
