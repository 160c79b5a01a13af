#!/usr/bin/env python
"""Wall time of every step of one resident graph build (needs a GPU): where the non-kernel time goes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isocon_b200 import _binding, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mode, ist = 1, None
if name == "c5":
    X, C = workloads.CONFIGS[name](scale=scale)
    L = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
    mode = 2
    ist = np.fromiter((1 if a in C else 0 for _, a in L), dtype=np.uint8, count=len(L))
else:
    S = workloads.CONFIGS[name](scale=scale)
    by_seq = {}
    for a, s in S.items():
        by_seq[s] = a
    L = sorted(by_seq.items(), key=lambda e: len(e[0]))
ctx = _binding.get_context(0)
t0 = time.perf_counter()
ctx.set_reads([s for s, _ in L])
print("set_reads %.2f ms (device %.2f)" % (1e3 * (time.perf_counter() - t0), ctx.last_ms(0)))
isq = np.ones(len(L), np.uint8) if ist is None else (1 - ist).astype(np.uint8)
for rep in range(3):
    row = []
    for label, fn in (("begin", lambda: ctx.graph_begin(mode, 2 ** 32, isq, ist)),
                      ("seed", lambda: ctx.graph_run(_binding.PHASE_SEED)),
                      ("pilot", lambda: ctx.graph_run(_binding.PHASE_PILOT)),
                      ("main", lambda: ctx.graph_run(_binding.PHASE_MAIN)),
                      ("wide", lambda: ctx.graph_run(_binding.PHASE_WIDE)),
                      ("finalize", lambda: ctx.graph_finalize()),
                      ("fetch", lambda: ctx.graph_fetch())):
        t0 = time.perf_counter()
        fn()
        row.append("%s %.2f" % (label, 1e3 * (time.perf_counter() - t0)))
    print("rep %d: %s | pair kernels %.2f ms" % (rep, "  ".join(row), ctx.last_ms(5)))
print(ctx.stats())
