#!/usr/bin/env python
"""Multi-GPU parity (needs >= 2 GPUs): the sharded graph (torchrun, one rank per GPU, NCCL + NVLink peer memory)
must equal the graph the same rank builds alone afterwards.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from isocon_b200 import _binding, workloads  # noqa: E402
from isocon_b200 import nearest_neighbor_graph as nn  # noqa: E402


class P(object):
    nr_cores = 1
    neighbor_search_depth = 2 ** 32
    verbose = False
    develop_logfile = None


def _tie_heavy(n, length):
    import numpy as np
    rng = np.random.default_rng(5)
    base = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=length)].tobytes().decode()
    S = {}
    for i in range(n):
        alt = "ACGT"[("ACGT".index(base[i]) + 1) % 4]
        S["t%d" % i] = base[:i] + alt + base[i + 1:]
    return S


def build(job, data):
    """One graph through the reference-facing call; job = (workload, scale, depth, edge reservation)."""
    name, scale, depth, reserve = job
    P.neighbor_search_depth = depth
    ctx = _binding.get_context()
    ctx.reserve_edges(reserve)
    so = sys.stdout; sys.stdout = open(os.devnull, "w")
    try:
        if name == "c5":
            g = nn.compute_2set_nearest_neighbor_graph(data[0], data[1], P())
        else:
            g = nn.compute_nearest_neighbor_graph(data, set(), P())[0]
    finally:
        sys.stdout = so
        ctx.reserve_edges(0)
        P.neighbor_search_depth = 2 ** 32
    return g, ctx.stats()["main_passes"]


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    # (workload, scale, neighbor_search_depth, edge reservation): default depth = pair-matrix algorithm (+ the
    # one-sided MAIN ladder across ranks for c5); finite depth: 1-set closed form inside the window, 2-set SCAN
    # algorithm (a single pass: the driver's MAIN loop must stop); a tiny edge buffer: overflow on some ranks ->
    # every rank learns it from the gather and the graph is rebuilt
    # c2 0.08 / c3 0.02: >= 512 queries, so the MAIN targets are laid out by similarity cluster (the ranks merge their
    # nearest-pilot-row records first); "foreign": 3 % of the reads carry N / n / * (general-alphabet passes)
    jobs = [("c2", 0.08, 2 ** 32, 0), ("c4", 0.01, 2 ** 32, 0), ("c3", 0.02, 2 ** 32, 0), ("c5", 0.03, 2 ** 32, 0),
            ("c5", 0.004, 2 ** 32, 0), ("c5", 0.01, 3, 0), ("c5", 0.01, 0, 0), ("c2", 0.03, 4, 0), ("c2", 0.03, 2 ** 32, -50),
            ("c5", 0.01, 2, -20), ("ties", 0.0, 2 ** 32, -1000), ("foreign", 0.06, 2 ** 32, 0), ("foreign", 0.02, 5, 0),
            ("c5", 0.06, 2 ** 32, 0)]     # 6000 reads x 300 candidates: min-hash clusters, hints, two-level passes

    def make(name, scale):
        if name == "ties":
            return _tie_heavy(300, 400)
        if name == "foreign":
            import numpy as np
            rng = np.random.default_rng(1)
            out = {}
            for a, s in workloads.config2(scale=scale).items():
                if rng.random() < 0.03:
                    b = bytearray(s.encode())
                    for p in rng.choice(len(b), size=max(1, len(b) // 50), replace=False):
                        b[p] = ord("Nn*"[int(rng.integers(0, 3))])
                    s = b.decode()
                out[a] = s
            return out
        return workloads.CONFIGS[name](scale=scale)

    data = [make(name, scale) for name, scale, _, _ in jobs]
    sharded = [build(job, d) for job, d in zip(jobs, data)]
    dist.barrier()
    dist.destroy_process_group()
    ok = True
    for job, d, (g, passes) in zip(jobs, data, sharded):
        alone, passes_alone = build(job, d)
        same = list(alone) == list(g) and all(list(alone[k].items()) == list(g[k].items()) for k in alone)
        ok = ok and same
        print("rank %d %s scale %g depth %d reserve %d: %d keys, %d edges, MAIN passes sharded %d / alone %d -> %s" % (
            rank, job[0], job[1], min(job[2], 99999), job[3], len(g), sum(len(v) for v in g.values()), passes,
            passes_alone, "same" if same else "DIFFERENT"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
