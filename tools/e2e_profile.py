#!/usr/bin/env python
"""Where the end-to-end call compute_nearest_neighbor_graph(S, ...) spends its wall time (needs a GPU)."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isocon_b200 import _binding, workloads  # noqa: E402
from isocon_b200 import nearest_neighbor_graph as nn  # noqa: E402


class P(object):
    nr_cores = 16
    neighbor_search_depth = 2 ** 32
    verbose = False
    develop_logfile = None


name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
S = workloads.CONFIGS[name](scale=scale)
for _ in range(2):
    nn.compute_nearest_neighbor_graph(S, set(), P())
ctx = _binding.get_context()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
nn.compute_nearest_neighbor_graph(S, set(), P())
pr.disable()
print("wall %.1f ms; device ms: set_reads %.2f graph %.2f finalize %.2f pair-kernels %.2f" % (
    1e3 * (time.perf_counter() - t0), ctx.last_ms(0), ctx.last_ms(1), ctx.last_ms(2), ctx.last_ms(5)))
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(18)
print(out.getvalue())
