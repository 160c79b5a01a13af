"""CPU suite, part 4: the multi-GPU driver (isocon_b200/sharding.run_sharded) with world_size 2 on
the gloo backend.  The device library is replaced by a TEST DOUBLE built on the oracle that, like
a real rank, only sees part of every query's row (targets t with t % world == rank); the driver's
MIN-reductions, tie filter and edge gather must reassemble the exact graph."""
import os
import socket

import numpy as np
import pytest

import util
from isocon_b200 import workloads


class OracleShardOps(object):
    """Stand-in for CudaShardOps: same seam, arithmetic from the oracle (tests only)."""

    def __init__(self, L, mode, is_query, is_target):
        self.L, self.mode, self.is_query, self.is_target = L, mode, is_query, is_target

    def begin(self, rank, world):
        import torch
        from oracle import oracle as O
        L, n = self.L, len(self.L)
        seqs = [s for s, _ in L]
        best = np.array([len(s) for s in seqs], dtype=np.int32)          # best_ed = len(seq1)
        a, b = [], []
        for q in range(n):
            if not self.is_query[q]:
                continue
            for t in range(n):
                if t != q and t % world == rank and (self.mode == 1 or self.is_target[t]):
                    a.append(q); b.append(t)
        d = O.ed_pairs(seqs, a, b) if a else np.zeros(0, np.int32)
        self.a, self.b, self.d = np.array(a, np.int32), np.array(b, np.int32), d
        self.best = torch.from_numpy(best)
        self.prev_cap = -1
        self.main_calls = 0

    CAPS = (60, 150, 10 ** 9)   # the MAIN phase climbs a ladder of caps like the device library (one pass per call)

    def run(self, phases):
        from isocon_b200 import _binding
        if phases != _binding.PHASE_MAIN:
            return 0
        self.main_calls += 1
        best = self.best.numpy()                      # reduced over the ranks since the last pass
        rows = [q for q in range(len(self.L)) if self.is_query[q] and (self.prev_cap < 0 or best[q] > self.prev_cap)]
        if self.prev_cap >= self.CAPS[-1] or not rows:
            return 0
        cap = self.CAPS[[c > self.prev_cap for c in self.CAPS].index(True)]
        todo = set(rows)
        for q, dist_ in zip(self.a.tolist(), self.d.tolist()):
            if q in todo and (self.mode == 2 or dist_ > 0) and dist_ <= min(best[q], cap) and dist_ < best[q]:
                best[q] = dist_
        self.prev_cap = cap
        return len(rows)

    def best_tensor(self):
        return self.best

    def finalize(self):
        import torch
        keep = self.d == self.best.numpy()[self.a] if self.a.size else np.zeros(0, bool)
        return (torch.from_numpy(self.a[keep].copy()), torch.from_numpy(self.b[keep].copy()),
                torch.from_numpy(self.d[keep].astype(np.int32)))

    def sync_before_collective(self):
        pass

    def sync_after_collective(self):
        pass


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from isocon_b200 import sharding
    from isocon_b200 import nearest_neighbor_graph as nn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = workloads.config2(scale=0.004)        # 40 reads
        L = sorted(((s, a) for a, s in S.items()), key=lambda e: len(e[0]))
        n = len(L)
        isq = np.ones(n, np.uint8); isq[3] = 0     # one converged read
        timing = {}
        best, q, t, d = sharding.run_sharded(OracleShardOps(L, 1, isq, None), dist, timing=timing)
        q, t, d = nn._order_edges(q, t, d)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), best=best, q=q, t=t, d=d)
        # 2-set
        ist = np.zeros(n, np.uint8); ist[::5] = 1
        ops2 = OracleShardOps(L, 2, 1 - ist, ist)
        ops2.gather_width = 3             # force the overflow round of the edge gather
        best, q, t, d = sharding.run_sharded(ops2, dist)
        q, t, d = nn._order_edges(q, t, d)
        np.savez(os.path.join(out_dir, "s%d.npz" % rank), best=best, q=q, t=t, d=d)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_reassemble_the_exact_graph(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    S = workloads.config2(scale=0.004)
    L = sorted(((s, a) for a, s in S.items()), key=lambda e: len(e[0]))
    n = len(L)
    hc = {L[3][0]}
    want = O.get_nearest_neighbors(L, 0, 0, L, hc, 2 ** 32)
    ist = np.zeros(n, np.uint8); ist[::5] = 1
    want2 = O.get_nearest_neighbors_2set(L, 0, L, {L[i][1] for i in range(n) if ist[i]}, 2 ** 32)
    for rank in range(2):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % rank))
        got = {a: {} for _, a in L}
        for q, t, d in zip(z["q"].tolist(), z["t"].tolist(), z["d"].tolist()):
            got[L[q][1]][L[t][1]] = d
        util.assert_same_graph(got, want, "rank %d 1-set" % rank)
        z = np.load(os.path.join(str(tmp_path), "s%d.npz" % rank))
        got = {L[i][1]: {} for i in range(n) if not ist[i]}
        for q, t, d in zip(z["q"].tolist(), z["t"].tolist(), z["d"].tolist()):
            got[L[q][1]][L[t][1]] = d
        util.assert_same_graph(got, want2, "rank %d 2-set" % rank)


def test_item_shards_cover_everything_once():
    # the C side deals row tiles round-robin: rank r takes tiles r, r + world, r + 2*world, ... < total
    for total in (0, 1, 7, 1000, 12345):
        for world in (1, 2, 3, 8):
            seen = np.zeros(total, np.int32)
            for r in range(world):
                seen[r:total:world] += 1
            assert (seen == 1).all()


def _worker_small(rank, world, port, out_dir):
    import torch.distributed as dist
    from isocon_b200 import sharding
    from isocon_b200 import nearest_neighbor_graph as nn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = workloads.config2(scale=0.0032)       # 32 reads: with three ranks some rank may own no edge at all
        L = sorted(((s, a) for a, s in S.items()), key=lambda e: len(e[0]))[:7]
        ops = OracleShardOps(L, 1, np.ones(len(L), np.uint8), None)
        ops.gather_width = 1                       # one edge per rank in the first round, the rest in the overflow round
        best, q, t, d = sharding.run_sharded(ops, dist)
        q, t, d = nn._order_edges(q, t, d)
        np.savez(os.path.join(out_dir, "w%d.npz" % rank), best=best, q=q, t=t, d=d)
    finally:
        dist.destroy_process_group()


def test_three_ranks_tiny_input_and_overflowing_gather(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    mp.spawn(_worker_small, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    S = workloads.config2(scale=0.0032)
    L = sorted(((s, a) for a, s in S.items()), key=lambda e: len(e[0]))[:7]
    want = O.get_nearest_neighbors(L, 0, 0, L, set(), 2 ** 32)
    for rank in range(3):
        z = np.load(os.path.join(str(tmp_path), "w%d.npz" % rank))
        got = {a: {} for _, a in L}
        for q, t, d in zip(z["q"].tolist(), z["t"].tolist(), z["d"].tolist()):
            got[L[q][1]][L[t][1]] = d
        util.assert_same_graph(got, want, "rank %d of 3" % rank)
