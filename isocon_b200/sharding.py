"""Multi-GPU driver of one graph build: one process per GPU, ``torch.distributed`` for the
plumbing.

The path shards without any data-path exchange: every rank holds the whole packed read set
(<= ~60 MB for the largest config) and computes a contiguous, equally sized slice of the row
tiles (one query x 256 targets each -- equal cost by construction).  With the peers' memory mapped over NVLink
(one box) the library runs the whole graph in one call and the ranks meet at device-side barriers (``run_fused``,
include/isocon_nn.h); otherwise two small reductions glue the ranks together:

1. ``all_reduce(MIN)`` on ``best[n]`` (int32) after each of the SEED, PILOT, MAIN and WIDE phases (and after
   every pass of the MAIN phase's threshold ladder) -- a rank only saw part of each row, so its running best
   is an upper bound;
2. ``all_gather`` of the edges that survive the tie filter ``distance == best[query]``.

Payload is a few bytes per read (<= 1 MB at N = 200k): latency-bound on NVSwitch, so parallel
efficiency is set by tile balance, not by the collectives.

``ShardOps`` is the seam between this driver and the device library so the driver itself is
testable with ``gloo`` on CPUs (tests/ plug in a test double; the product only ever uses
``CudaShardOps``).
"""
import os

import numpy as np

from . import _binding


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # torch missing: single process
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class CudaShardOps(object):
    """The device library behind the driver (the only implementation the product uses)."""

    def __init__(self, ctx, mode, depth, is_query, is_target, algo=_binding.ALGO_AUTO, symmetric=True):
        self.ctx, self.mode, self.depth = ctx, mode, depth
        self.is_query, self.is_target, self.algo, self.symmetric = is_query, is_target, algo, symmetric
        self.n_reads = ctx.n

    def begin(self, rank, world):
        self.ctx.graph_begin(self.mode, self.depth, self.is_query, self.is_target, self.algo, self.symmetric,
                             rank=rank, world=world)

    def run(self, phases):
        self.ctx.graph_run(phases)
        return self.ctx.last_run_rows()

    def connect_peers(self, dist, group=None):
        """Map the other ranks' best[] and counters over NVLink (CUDA IPC): the pair kernels push every
        improvement of best[] to all GPUs of the box while they run, and all ranks pull their row tiles
        from one queue in rank 0's memory.  Handles are exchanged only when the allocation moved (every
        rank runs the same use_list sequence, so all ranks agree on when).  Ranks that cannot map each other
        (more than 8, several nodes, no peer access) agree to do without: tiles are then dealt round-robin and
        only the collectives of run_sharded connect the ranks."""
        import torch
        ctx = self.ctx
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if os.environ.get("ISOCON_NN_P2P", "1") == "0" or ctx.n == 0:
            return
        handle, generation = ctx.ipc_handles()
        key = (generation, world, rank)
        if getattr(ctx, "_peer_key", None) == key:
            return
        dev = "cuda:%d" % ctx.device
        one_box = world <= 8 and int(os.environ.get("LOCAL_WORLD_SIZE", str(world))) == world
        mine = torch.from_numpy(handle).to(dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        failed = 0 if one_box else 1
        if one_box:
            try:
                ctx.set_peers(torch.stack(parts).cpu().numpy(), world, rank)
            except _binding.IsoconNNError:
                failed = 1
        flag = torch.tensor([failed], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)   # doubles as the barrier: every rank has dropped
        if int(flag.item()):                                       # its mapping of outgrown allocations
            ctx.set_peers(None, 1, 0)                              # every rank: static sharding, no peer memory
        ctx.release_retired()
        ctx._peer_key = key

    def best_tensor(self):
        import torch
        if self.ctx.n == 0:
            return torch.zeros(0, dtype=torch.int32, device="cuda:%d" % self.ctx.device)
        return torch.as_tensor(self.ctx.best_dev(), device="cuda:%d" % self.ctx.device)

    def finalize(self):
        """(q, t, d, needed): this rank's edges at the global best; needed > 0 = the candidate-edge buffer
        overflowed on this rank (that many edges) and the graph must be rebuilt with more room."""
        import torch
        dev = "cuda:%d" % self.ctx.device
        z = torch.zeros(0, dtype=torch.int32, device=dev)
        try:
            ne = self.ctx.graph_finalize()
        except _binding.IsoconNNError as e:
            if e.code != _binding.ERR_OVERFLOW:
                raise
            return z, z.clone(), z.clone(), int(self.ctx.stats()["edges_raw"])
        if ne == 0:
            return z, z.clone(), z.clone(), 0
        q, t, d = self.ctx.edges_dev()
        return (torch.as_tensor(q, device=dev), torch.as_tensor(t, device=dev), torch.as_tensor(d, device=dev), 0)

    def reserve_edges(self, capacity):
        self.ctx.reserve_edges(capacity)

    def can_fuse(self):
        return self.ctx.can_fuse()

    def run_fused(self):
        """All phases in one library call (device-side barriers over NVLink peer memory, every rank receives every
        rank's edges): (best, q, t, d) as numpy, or the number of edges to reserve after an overflow -- the same on
        every rank."""
        self.ctx.graph_run(_binding.PHASE_ALL)
        try:
            self.ctx.graph_finalize()
        except _binding.IsoconNNError as e:
            if e.code != _binding.ERR_OVERFLOW:
                raise
            return int(self.ctx.stats()["edges_raw"])
        return self.ctx.graph_fetch()

    def merge_pilot_near(self, dist, group=None):
        """Similarity order of the MAIN phase: every rank recorded the two nearest pilot rows of each read among the
        pairs IT aligned; all ranks need the same records (they build the same target layout and tile table).
        One all-gather, then the two smallest (distance, row) per read."""
        import torch
        view = self.ctx.pilot_near_dev()
        if view is None:
            return False
        dev = "cuda:%d" % self.ctx.device
        mine = torch.as_tensor(view, device=dev)
        world, n = dist.get_world_size(group), mine.numel() // 2
        parts = torch.empty(world * mine.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(parts, mine, group=group)
        v = parts.view(world * 2, n)
        big = torch.iinfo(torch.int64).max
        v = torch.where(v < 0, torch.full_like(v, big), v)          # ~0 (none) is -1 as int64
        two = torch.sort(v, dim=0).values[:2]
        two = torch.where(two == big, torch.full_like(two, -1), two)
        mine.copy_(two.reshape(-1))
        torch.cuda.synchronize(self.ctx.device)
        return True

    def agree(self):
        self.ctx.best_agree()

    def sync_before_collective(self):
        self.ctx.sync()

    def sync_after_collective(self):
        import torch
        torch.cuda.synchronize(self.ctx.device)


class _CollectiveTimer(object):
    """Device time of the collectives (CUDA events on torch's current stream); no-op on CPU."""

    def __init__(self, enabled):
        self.enabled, self.pairs = enabled, []

    def __enter__(self):
        if self.enabled:
            import torch
            self.e0 = torch.cuda.Event(enable_timing=True); self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.enabled:
            self.e1.record()
            self.pairs.append((self.e0, self.e1))
        return False

    def total_ms(self):
        if not self.enabled:
            return 0.0
        import torch
        torch.cuda.synchronize()
        return float(sum(a.elapsed_time(b) for a, b in self.pairs))


def run_sharded(ops, dist, group=None, timing=None):
    """SPMD: every rank calls this; returns (best[n], edge_q, edge_t, edge_d) as numpy on every rank.
    ``timing`` (dict, optional) receives ``collective_ms`` (device time spent in the collectives) and
    ``host_ms`` (wall time per section of this function).  When the candidate-edge buffer of any rank overflowed
    (tie-heavy input) every rank learns it from the edge gather, reserves more and the graph is built again."""
    for _ in range(8):
        out = _run_sharded_once(ops, dist, group, timing)
        if isinstance(out, tuple):
            return out
        ops.reserve_edges(out * 3 // 2 + 4096)
    raise _binding.IsoconNNError(_binding.ERR_OVERFLOW, "candidate edge buffer still too small after regrowing")


def _run_sharded_once(ops, dist, group=None, timing=None):
    import time
    import torch
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    marks = [("start", time.perf_counter())]

    def mark(name):
        if timing is not None:
            marks.append((name, time.perf_counter()))

    ops.begin(rank, world)
    if hasattr(ops, "connect_peers"):
        ops.connect_peers(dist, group)
    if hasattr(ops, "can_fuse") and ops.can_fuse():       # peers mapped: no collective on the data path
        mark("begin")
        out = ops.run_fused()
        mark("fused")
        if timing is not None:
            timing["collective_ms"] = 0.0
            timing["host_ms"] = {b[0]: 1e3 * (b[1] - a[1]) for a, b in zip(marks, marks[1:])}
        return out
    best = ops.best_tensor()
    token = torch.zeros(1, dtype=torch.int32, device=best.device)
    timer = _CollectiveTimer(best.is_cuda)
    mark("begin")

    def phase(which, name):
        rows = ops.run(which)             # rows of the pair matrix the phase covered on ALL ranks together
        mark(name)
        if rows == 0:                     # same number on every rank: nothing ran anywhere, best[] is unchanged
            return 0
        ops.sync_before_collective()
        if best.numel():
            with timer:
                dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
        ops.sync_after_collective()
        # every rank keeps a snapshot of the agreed best[] for the host-side decisions of its next phase, and no rank
        # starts that phase (its kernels lower the peers' live best[] over NVLink) before all have their snapshot
        if hasattr(ops, "agree"):
            ops.agree()
            if which == _binding.PHASE_PILOT and hasattr(ops, "merge_pilot_near") and ops.merge_pilot_near(dist, group):
                pass                       # its all-gather is the barrier
            else:
                with timer:
                    dist.all_reduce(token, group=group)
                ops.sync_after_collective()
        mark(name + "_reduce")
        return rows

    phase(_binding.PHASE_SEED, "seed")    # each rank seeds its share of the queries
    phase(_binding.PHASE_PILOT, "pilot")  # symmetric graph: first rows against everything behind them
    # targets re-binned by class from the global best; each rank aligns its tiles.  One-sided graphs climb a ladder
    # of threshold caps, one pass per call: which rows are still unresolved is decided from the reduced best[]
    passes = 0
    while phase(_binding.PHASE_MAIN, "main%d" % passes if passes else "main") > 0:
        passes += 1
    phase(_binding.PHASE_WIDE, "wide")    # needs the global best to know which rows are unresolved
    fin = ops.finalize()                  # local edges whose distance equals the GLOBAL best
    q, t, d = fin[0], fin[1], fin[2]
    needed = int(fin[3]) if len(fin) > 3 else 0
    ops.sync_before_collective()
    mark("finalize")
    width_override = getattr(ops, "gather_width", None)
    # ONE collective for the edges: every rank contributes a fixed-width record [count, q.., t.., d..] (width from
    # the number of reads -- an NN graph has about one edge per read); a rank with more edges than fit sends the
    # rest in a second, exactly sized round (rare).  Counts travel with the payload: no separate exchange, one D2H.
    n_reads = int(getattr(ops, "n_reads", 0)) or int(best.numel())
    width = width_override or max(4096, 2 * n_reads // world)
    ne = int(q.numel())
    head = min(ne, width)
    H = 2                                 # header words: [edge count, edges needed after an overflow (0 = fine)]
    mine = torch.zeros(H + 3 * width, dtype=torch.int32, device=q.device)
    mine[0] = ne
    mine[1] = min(needed, 2 ** 31 - 1)
    if head:
        mine[H:H + head] = q[:head]
        mine[H + width:H + width + head] = t[:head]
        mine[H + 2 * width:H + 2 * width + head] = d[:head]
    parts = torch.empty(world * (H + 3 * width), dtype=torch.int32, device=q.device)
    with timer:
        dist.all_gather_into_tensor(parts, mine, group=group)
    ops.sync_after_collective()
    mark("gather")
    host = parts.cpu().numpy().reshape(world, H + 3 * width)       # one D2H for counts and edges
    if int(host[:, 1].max()) > 0:         # same decision on every rank
        return int(host[:, 1].max())
    counts = [int(c) for c in host[:, 0]]
    heads = [min(c, width) for c in counts]
    allq = [host[r, H:H + h] for r, h in enumerate(heads)]
    allt = [host[r, H + width:H + width + h] for r, h in enumerate(heads)]
    alld = [host[r, H + 2 * width:H + 2 * width + h] for r, h in enumerate(heads)]
    over = max(counts) - width
    if over > 0:                          # same decision on every rank: the counts are common knowledge now
        rest = torch.zeros((3, over), dtype=torch.int32, device=q.device)
        if ne > width:
            rest[0, :ne - width] = q[width:]; rest[1, :ne - width] = t[width:]; rest[2, :ne - width] = d[width:]
        more = torch.empty(world * 3 * over, dtype=torch.int32, device=q.device)
        with timer:
            dist.all_gather_into_tensor(more, rest.view(-1), group=group)
        ops.sync_after_collective()
        more = more.cpu().numpy().reshape(world, 3, over)
        for r, c in enumerate(counts):
            if c > width:
                allq.append(more[r, 0, :c - width]); allt.append(more[r, 1, :c - width]); alld.append(more[r, 2, :c - width])
    allq, allt, alld = np.concatenate(allq), np.concatenate(allt), np.concatenate(alld)
    best_host = best.cpu().numpy()
    mark("fetch")
    if timing is not None:
        timing["collective_ms"] = timer.total_ms()
        timing["host_ms"] = {b[0]: 1e3 * (b[1] - a[1]) for a, b in zip(marks, marks[1:])}
    return best_host, allq, allt, alld


def device_graph(ctx, mode, depth, is_query, is_target, algo=_binding.ALGO_AUTO, symmetric=True):
    """(best, edge_q, edge_t, edge_d) for the resident reads: single GPU, or all ranks together."""
    dist = _dist()
    if dist is None:
        return ctx.graph(mode, depth, is_query, is_target, algo, symmetric)
    return run_sharded(CudaShardOps(ctx, mode, depth, is_query, is_target, algo, symmetric), dist)
