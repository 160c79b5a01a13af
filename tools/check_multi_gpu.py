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


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    jobs = [("c2", 0.05), ("c4", 0.01), ("c3", 0.02), ("c5", 0.03), ("c5", 0.004)]
    data = []
    for name, scale in jobs:
        data.append(workloads.CONFIGS[name](scale=scale))
    sharded = []
    devnull = open(os.devnull, "w")
    for (name, scale), d in zip(jobs, data):
        so = sys.stdout; sys.stdout = devnull
        try:
            g = nn.compute_2set_nearest_neighbor_graph(d[0], d[1], P()) if name == "c5" else nn.compute_nearest_neighbor_graph(d, set(), P())[0]
        finally:
            sys.stdout = so
        stats = _binding.get_context().stats()
        sharded.append((g, stats["main_passes"]))
    dist.barrier()
    dist.destroy_process_group()
    ok = True
    for (name, scale), d, (g, passes) in zip(jobs, data, sharded):
        so = sys.stdout; sys.stdout = devnull
        try:
            alone = nn.compute_2set_nearest_neighbor_graph(d[0], d[1], P()) if name == "c5" else nn.compute_nearest_neighbor_graph(d, set(), P())[0]
        finally:
            sys.stdout = so
        same = list(alone) == list(g) and all(list(alone[k].items()) == list(g[k].items()) for k in alone)
        ok = ok and same
        print("rank %d %s scale %g: %d keys, %d edges, MAIN passes sharded %d / alone %d -> %s" % (
            rank, name, scale, len(g), sum(len(v) for v in g.values()), passes,
            _binding.get_context().stats()["main_passes"], "same" if same else "DIFFERENT"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
