#!/usr/bin/env python
"""Where the end-to-end call spends its wall time (needs a GPU): cProfile of one call with the reads resident
("warm") and one with the device store emptied first ("cold").   usage: e2e_profile.py [c2|c3|c4|c5] [scale]"""
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
data = workloads.CONFIGS[name](scale=scale)


def call():
    if name == "c5":
        return nn.compute_2set_nearest_neighbor_graph(data[0], data[1], P())
    return nn.compute_nearest_neighbor_graph(data, set(), P())


for _ in range(2):
    call()
ctx = _binding.get_context()
for mode in ("warm", "cold"):
    walls = []
    for rep in range(3):
        if mode == "cold":
            ctx.store_reset()
        t0 = time.perf_counter()
        call()
        walls.append(1e3 * (time.perf_counter() - t0))
    if mode == "cold":
        ctx.store_reset()
    pr = cProfile.Profile()
    pr.enable()
    call()
    pr.disable()
    print("== %s: wall %s ms; device ms of the last call: upload %.2f graph %.2f finalize %.2f pair-kernels %.2f" % (
        mode, " ".join("%.2f" % w for w in walls), ctx.last_ms(0), ctx.last_ms(1), ctx.last_ms(2), ctx.last_ms(5)))
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(16)
    print("\n".join(out.getvalue().splitlines()[4:30]))
