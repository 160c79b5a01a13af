#!/usr/bin/env python
"""bench.py -- all-vs-all nearest-neighbour-graph throughput on B200 (BASELINE.json metric).

One "step" = one complete 1-set NN-graph build over the workload (default: BASELINE.json
configs[1] = "c2": synthetic 10k reads x 1.5 kb, 20 near-identical gene copies, 5 % error).

  value      GCUPS = cells_full / (device time of exactly K resident steps / K): graph_begin + SEED/PILOT/MAIN/WIDE
             + tie filter + results on the host (+ the NCCL reductions and the edge gather at N > 1), packed reads
             already in HBM, bracketed by
             barrier + synchronize, CUDA events on the library's stream, max over ranks.
             cells_full = sum of len(q)*len(t) over every pair the REFERENCE hands to edlib on this
             input (oracle counter, tests/golden/bench_<workload>.json) -- the conventional,
             implementation-independent GCUPS numerator.
  e2e        the same numerator over the wall time of compute_nearest_neighbor_graph(S, ...) called
             with host dicts: 2-bit packing + H2D + kernels + D2H + dict rebuild inside the timer.
  roofline   dominant kernel (nn_row_kernel, PILOT + MAIN launches): the DP cells the thresholds made necessary
             (counted by the kernel) * 9/32 int-ops/cell / its CUDA-event duration vs the INT32 ALU issue rate
             measured by the library's probe kernel on this very GPU (integer bit-parallel DP: not HBM-, not
             tensor-bound).
             roofline_hbm shows the HBM side for contrast.
  cpu_baseline / --impl reference
             the oracle's C++ port of the reference scan + 64-bit Myers (edlib-compatible
             stand-in; real edlib is not installable here) on the box's host cores, on a bounded
             sample of the same workload's queries.

Multi-GPU (torchrun, one rank per GPU): row tiles are split across ranks, best[] is MIN-reduced
and surviving edges gathered over NCCL; value = cells_full / max-over-ranks step time.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

INT_OPS_PER_CELL_SURVEY = 0.375  # SURVEY.md §8d's a-priori estimate: 12 integer instructions per 32-cell word-column
INT_OPS_PER_CELL = 9.0 / 32.0    # the shipped recurrence (diag_band.cuh): 7 LOP3 + 1 IADD3.X + 1 SHF per 32 cells
# ALU-pipe instructions in the SASS of the unrolled column body (tools/sass_loops.py): 9 per word (7 LOP3 + IADD3 + SHF)
# plus 2.8 per column whatever the width (symbol extraction, bottom-diagonal bit, address): W = 5: 9.56 per
# word-column, W = 1: 11.8
ALU_INSTR_PER_WORD = 9.0
ALU_INSTR_PER_COLUMN = 2.8
WORKLOAD_DESC = {
    "c2": "c2: synthetic 10k reads x 1.5 kb, 20 near-identical gene copies, 5% indel-heavy error, 1-set all-vs-all NN graph",
    "c3": "c3: synthetic 50k Iso-Seq-like reads x 3 kb, 100 paralogs 0.5-2% apart, 2% error, 1-set NN graph",
    "c4": "c4: synthetic 200k ONT-like amplicon reads x 1 kb, 10% error, 1-set NN graph",
    "c5": "c5: 2-set NN graph for stat_filter, 100k reads vs 5k candidates x 2.5 kb, 3% error",
}


class Params(object):
    nr_cores = 16
    neighbor_search_depth = 2 ** 32
    verbose = False
    develop_logfile = None


def fingerprint(seqs):
    h = hashlib.sha256()
    for s in seqs:
        h.update(s.encode()); h.update(b"\n")
    return h.hexdigest()[:16]


def graph_digest(G):
    return hashlib.sha256(json.dumps([[a, list(v.items())] for a, v in G.items()]).encode()).hexdigest()[:16]


class Workload(object):
    """The sorted list a graph build works on, the masks of the call, and the reference-facing call itself."""

    def __init__(self, name, scale):
        from isocon_b200 import workloads
        self.name, self.scale = name, scale
        if name == "c5":
            self.X, self.C = workloads.CONFIGS[name](scale=scale)
            self.mode = 2
            self.lst = sorted([(s, a) for a, s in self.X.items()] + [(s, a) for a, s in self.C.items()],
                              key=lambda e: len(e[0]))
            self.is_target = np.fromiter((1 if a in self.C else 0 for _, a in self.lst), dtype=np.uint8, count=len(self.lst))
            self.is_query = (1 - self.is_target).astype(np.uint8)
        else:
            self.S = workloads.CONFIGS[name](scale=scale)
            self.mode = 1
            by_seq = {}
            for a, s in self.S.items():
                by_seq[s] = a
            self.lst = sorted(by_seq.items(), key=lambda e: len(e[0]))
            self.is_target = None
            self.is_query = np.ones(len(self.lst), np.uint8)
        self.seqs = [s for s, _ in self.lst]
        tag = name if scale == 1.0 else "%s_s%g" % (name, scale)
        gold_path = os.path.join(ROOT, "tests", "golden", "bench_%s.json" % tag)
        self.gold = None
        if os.path.exists(gold_path):
            with open(gold_path) as fh:
                self.gold = json.load(fh)
            if self.gold.get("fingerprint") != fingerprint(self.seqs):
                self.gold = None     # generator drift: the stored counters do not describe this input

    def call(self, nn, params):
        """The call a user of the reference makes (graphs.py:58 / graphs.py:154)."""
        if self.mode == 2:
            return nn.compute_2set_nearest_neighbor_graph(self.X, self.C, params)
        return nn.compute_nearest_neighbor_graph(self.S, set(), params)[0]

    def all_pairs_cells(self):
        lens = np.array([len(s) for s in self.seqs], dtype=np.float64)
        if self.mode == 2:
            return float(lens[self.is_query == 1].sum() * lens[self.is_target == 1].sum())
        return float(lens.sum() ** 2 - (lens ** 2).sum())       # every ordered pair


# ----------------------------------------------------------------------------- CPU side (oracle)

def cpu_sample(wl, n_queries, threads):
    """Oracle port on host cores: `n_queries` evenly spaced queries of the workload, each scanned
    against the whole list exactly as the reference does (no cross-query seeding), one query per
    worker call, `threads` worker threads.  Returns (cells_full, cells_band, calls, wall_s, queries)."""
    from oracle import oracle as O
    n = len(wl.lst)
    cand = np.flatnonzero(wl.is_query)
    n_queries = max(1, min(n_queries, cand.size))
    qs = cand[np.unique(np.linspace(0, cand.size - 1, n_queries).astype(np.int64))]
    cat, off = O.concat(wl.seqs)
    mask = np.zeros(max(n, 1), np.uint8) if wl.mode == 1 else np.ascontiguousarray(wl.is_target)
    L = O.lib()
    fn = L.nn_oracle_1set if wl.mode == 1 else L.nn_oracle_2set
    import concurrent.futures as cf
    tot = np.zeros(4, dtype=np.uint64)

    def one(q):
        best = np.full(n, -1, np.int32)
        st = np.zeros(4, np.uint64)
        cap = 4096
        eq = np.empty(cap, np.int32); et = np.empty(cap, np.int32); ed = np.empty(cap, np.int32)
        fn(cat, off, n, mask, 2 ** 32, int(q), 1, 1, 1, 0, best, eq, et, ed, cap, st)   # releases the GIL
        return st

    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        for st in ex.map(one, qs.tolist()):
            tot += st
    wall = time.perf_counter() - t0
    return int(tot[2]), int(tot[3]), int(tot[0]), wall, int(qs.size)


def minimal_band_cells(wl, G, samples=2000000, seed=11):
    """Algorithmic minimum of the DP work behind a graph: every pair the graph must consider, once
    (1-set: unordered -- d(q,t) = d(t,q); 2-set: read x candidate), inside the Ukkonen strip of the
    SMALLEST threshold that still proves the answer, i.e. the final best distance of the query (of the
    farther-off end for an unordered pair): cells = n * min(m, |n-m| + 2*floor((k-|n-m|)/2) + 1), or 0 when
    the lengths alone exclude the pair (|n-m| > k).  Depends only on the input and on the (verified)
    output graph, not on how any implementation orders its alignments.  Exact below 5M pairs, else a
    seeded Monte-Carlo estimate over `samples` pairs."""
    lens = np.array([len(s) for s in wl.seqs], dtype=np.int64)
    best = lens.copy()                                   # no neighbour within len(q): the bound stays len(q)
    pos = {a: i for i, (_, a) in enumerate(wl.lst)}
    for a, nbrs in G.items():
        if nbrs:
            best[pos[a]] = min(nbrs.values())
    rng = np.random.default_rng(seed)
    n = lens.size
    if wl.mode == 1:
        total = n * (n - 1) // 2
        if total <= 5000000:
            a, b = np.triu_indices(n, 1)
        else:
            a = rng.integers(0, n, size=samples); b = rng.integers(0, n, size=samples)
            keep = a != b
            a, b = a[keep], b[keep]
        k = np.maximum(best[a], best[b])
    else:
        qs = np.flatnonzero(wl.is_query); ts = np.flatnonzero(wl.is_target)
        total = qs.size * ts.size
        if total <= 5000000:
            a = np.repeat(qs, ts.size); b = np.tile(ts, qs.size)
        else:
            a = rng.choice(qs, size=samples); b = rng.choice(ts, size=samples)
        k = best[a]
    m, nn_ = lens[a], lens[b]
    dl = np.abs(nn_ - m)
    w = dl + 2 * ((k - dl) // 2) + 1
    cells = np.where(dl <= k, np.minimum(nn_ * np.minimum(m, w), m * np.minimum(nn_, w)), 0).astype(np.float64)
    return float(cells.mean() * total) if a.size else 0.0


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.workload, args.scale)
    lst = wl.lst
    threads = host_threads()
    total_q = int(wl.is_query.sum())
    # default: every 16th query (c2 on 16 cores: about 5 s per step); --cpu-queries -1 = every query (same_config)
    nq = total_q if args.cpu_queries < 0 else (args.cpu_queries or max(threads * 8, 64, total_q // 16))
    for _ in range(args.warmup):
        cpu_sample(wl, max(threads, 8), threads)
    cells = wall = 0.0
    for _ in range(args.steps):
        cf_, cb_, calls, w, used = cpu_sample(wl, nq, threads)
        cells += cf_; wall += w
    gcups = cells / wall / 1e9
    c1, _, _, w1, _ = cpu_sample(wl, 4, 1)        # the same port on ONE core, for context
    sample = "%d of %d queries (evenly spaced) x the whole list per step, %d threads" % (used, int(wl.is_query.sum()), threads)
    line = {
        "impl": "reference", "metric": "all-vs-all NN-graph GCUPS", "value": gcups, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "scale": args.scale, "reads": len(lst)},
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port", "sample": sample,
                         "sample_fraction": used / float(total_q), "same_config": used == total_q,
                         "single_thread_gcups": c1 / w1 / 1e9,
                         "note": "oracle C++ port of the reference scan + 64-bit Myers (edlib-compatible stand-in; "
                                 "edlib itself is absent); no Python-per-pair overhead, so faster than the real reference"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks

class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc, self.path, self.device = None, None, device

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as fh:
            for line in fh:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return None
        busy = [x for x in sm if x >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu-queries", type=int, default=0,
                    help="queries in the CPU sample (0 = default: 8 x threads in the GPU arm, every 16th query in the "
                         "reference arm; -1 = every query)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("ISOCON_NN_DEVICE", str(local_rank))

    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from isocon_b200 import _binding, sharding
    from isocon_b200 import nearest_neighbor_graph as nn

    wl = Workload(args.workload, args.scale)
    lst, gold, seqs = wl.lst, wl.gold, wl.seqs
    n = len(lst)
    ctx = _binding.get_context(local_rank)
    int32_peak = ctx.int32_peak()
    ctx.use_list(seqs)
    isq, ist = wl.is_query, wl.is_target
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    shard_timing = {}

    def resident_step():
        """One graph build with the packed reads resident in HBM; returns the device milliseconds of
        the MAIN-phase tile kernel of this rank (the dominant kernel, CUDA events around its launch)."""
        flush_buf.zero_()                     # flush L2 between steps (256 MiB > 126 MB L2)
        torch.cuda.synchronize()
        if dist is None:
            ctx.graph_begin(wl.mode, 2 ** 32, isq, ist)
            ctx.graph_run(_binding.PHASE_ALL)
            ctx.graph_finalize()
            ctx.graph_fetch()                 # best[] and the edges on the host, like every rank of the sharded step
        else:
            sharding.run_sharded(sharding.CudaShardOps(ctx, wl.mode, 2 ** 32, isq, ist), dist, timing=shard_timing)
        return ctx.last_ms(5)

    def e2e_step(cold=True):
        """The reference-facing call with host dicts.  cold: the device store is emptied first, so every read of
        the call is gathered, copied host -> device and packed inside the timed region (the first round of a
        pipeline run); otherwise the reads are resident from the previous call (a later correction round whose
        reads did not change: nothing but the masks and the list order travels)."""
        flush_buf.zero_()
        torch.cuda.synchronize()
        if cold:
            ctx.store_reset()
        return wl.call(nn, Params())

    import contextlib
    import io
    quiet = contextlib.redirect_stdout(io.StringIO())

    # ---- warm-up (also the full-size parity gate)
    G = None
    for _ in range(max(args.warmup, 1)):
        resident_step()
        with quiet:
            G = e2e_step()
    parity = None
    if gold is not None and not args.no_parity:
        parity = (graph_digest(G) == gold["digest"])
        if not parity:
            raise SystemExit("PARITY FAILURE: graph digest %s != oracle %s" % (graph_digest(G), gold["digest"]))

    # ---- timed region 1: exactly K resident steps, bracketed by barrier + synchronize; CUDA events on the
    # library's stream (the GPU idles on that stream while the host works, so host gaps are inside the bracket)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ctx.timer_start()
    t0 = time.perf_counter()
    main_ms = [resident_step() for _ in range(args.steps)]
    torch.cuda.synchronize()
    step_ms = ctx.timer_stop() / args.steps
    step_wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    barrier()
    stats = ctx.stats()
    # ---- timed region 2: K end-to-end calls of the reference-facing function with host dicts, every read
    # uploaded inside the timer
    # (the returned graphs are kept until the timer has stopped: dropping a dict of 100 000 dicts costs as much as
    # a tenth of building it and is no part of the call)
    held = []
    up0 = ctx.store_info()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        with quiet:
            held.append(e2e_step(cold=True))
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    up1 = ctx.store_info()
    del held[:]
    barrier()
    # ---- timed region 3: the same call with the reads resident (round k+1 of the pipeline, nothing changed)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        with quiet:
            held.append(e2e_step(cold=False))
    torch.cuda.synchronize()
    e2e_warm_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    del held[:]
    up2 = ctx.store_info()
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    main_kernel_ms = sum(main_ms) / len(main_ms)
    if dist is not None:
        t = torch.tensor([step_ms, e2e_ms, main_kernel_ms, step_wall_ms, e2e_warm_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms, main_kernel_ms, step_wall_ms, e2e_warm_ms = [float(x) for x in t.tolist()]
        keys = ["pairs", "word_columns", "groups", "items", "edges_raw", "launches", "useful_cells", "columns"]
        cnt = torch.tensor([stats[k] for k in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        for k, v in zip(keys, cnt.tolist()):
            stats[k] = int(v)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- numerators
    if gold is not None:
        cells_full, cells_band = gold["work"]["cells_full"], gold["work"]["cells_band"]
        numerator = "oracle run of the whole workload (tests/golden/bench_%s.json, reference chunking nr_cores=16)" % args.workload
    else:
        cells_full = wl.all_pairs_cells()       # every ordered (query, target) pair (lengths within the bound)
        cells_band = None
        numerator = "all (query, target) pairs (no oracle counters stored for this input)"

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = host_threads()
        nq = int(isq.sum()) if args.cpu_queries < 0 else (args.cpu_queries or max(threads * 8, 64))
        cf_, cb_, calls, wall, used = cpu_sample(wl, nq, threads)
        cpu = {"value": cf_ / wall / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
               "sample": "%d of %d queries (evenly spaced) x the whole list, %.1f s wall, %d edit-distance calls" % (
                   used, int(isq.sum()), wall, calls),
               "cells_band_per_cells_full": cb_ / cf_}
        try:    # the same port on ONE core, for context (4 queries, about a second)
            c1, _, _, w1, _ = cpu_sample(wl, 4, 1)
            cpu["single_thread_gcups"] = c1 / w1 / 1e9
        except Exception:
            pass
        if cells_band is None:
            cells_band = cb_ * (float(isq.sum()) / used)     # extrapolated from the sample
            numerator += "; cells_band extrapolated from the CPU sample"

    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_ncu.json")) as fh:
            ncu = json.load(fh).get(args.workload if args.scale == 1.0 else "", {})
    except Exception:
        pass
    bytes_in = sum(len(s) for s in seqs) + 8 * (n + 1)
    n_edges = sum(len(v) for v in G.values())
    # Algorithmic work of the dominant kernel, counted by the kernel itself per aligned pair: (columns until
    # that pair's answer was known) x (rows of that pair's own Ukkonen strip for the threshold in force) --
    # no word padding, no lock-step waiting, no bookkeeping -- at 9 integer instructions per 32 cells.
    t_k = main_kernel_ms * 1e-3
    peak_all = int32_peak * world
    useful = float(stats["useful_cells"])
    alg_ops = useful * INT_OPS_PER_CELL
    achieved = alg_ops / t_k / 1e12
    executed = 32 * (stats["word_columns"] * ALU_INSTR_PER_WORD + stats["columns"] * ALU_INSTR_PER_COLUMN) / t_k / 1e12
    kernel_name = "nn_row_kernel (the PILOT + MAIN [+ WIDE] launches of one step, summed)"
    if wl.mode == 2 and stats["clusters"] > 0 and stats["word_columns"] > 0:
        # two-level one-sided pass: the alignments are spread over nn_row_swapped_kernel
        # (hinted SEED launch), nn_tile_kernel (sample, repeats, level 2; no useful-cell counter) and nn_row_kernel
        # (level 1 of the later ladder passes): report the executed word-columns at the recurrence's 9 instructions
        kernel_name = ("pair kernels of the two-level pass: nn_row_swapped_kernel (hinted SEED launch), nn_tile_kernel "
                       "(sample, repeats, level 2), nn_row_kernel (level 1 by alignment); level 1 of the first pass is the "
                       "q-gram filter kernel, no alignment")
        achieved = executed
    roofline = {"bound": "int32", "kernel": kernel_name,
                "achieved": achieved, "peak": peak_all / 1e12, "unit": "Tint-op/s", "frac": achieved / (peak_all / 1e12),
                "peak_source": "measured on this GPU by isocon_nn_int32_peak (LOP3/IADD3 probe kernel) x %d GPU(s)" % world,
                "algorithmic_ops": alg_ops, "useful_cells": useful, "int_ops_per_cell": INT_OPS_PER_CELL,
                "kernel_ms": main_kernel_ms,
                "numerator_source": "counted by the kernel itself (stats.useful_cells): the share of the executed work that the "
                                    "thresholds in force made necessary -- NOT an implementation-independent bound; the only "
                                    "independent numerators printed here (cells_final_thresholds, survey_convention) are not "
                                    "lower bounds of the work of an exact algorithm either (a better alignment order needs "
                                    "less) and exceed the peak",
                "independent_bound": None,
                "executed_lane_word_columns": stats["word_columns"] * 32,
                "mean_window_words": stats["word_columns"] / max(1, stats["columns"]),
                "executed_alu_ops_frac_of_peak": executed / (peak_all / 1e12),
                "ncu_alu_pipe_pct_of_peak": ncu.get("alu_pipe_pct"), "ncu_source": ncu.get("source"),
                "traffic": ncu.get("dram_bytes_per_launch"),
                "cells_final_thresholds": minimal_band_cells(wl, G),
                "note": "frac = necessary integer ops / measured INT32 issue peak; necessary = for every aligned pair and "
                        "every column until ITS answer was known, the rows of its own Ukkonen strip that the window "
                        "in force still held (the window shrinks as cells die: D + distance to the final diagonal > k), "
                        "9 instr per 32 cells. The gap to executed_alu_ops_frac_of_peak (9 ALU instr per word-column + "
                        "2.8 per column, cross-checked by ncu sm__inst_executed_pipe_alu) is 32-bit word granularity, "
                        "lanes waiting for the slowest pair of their warp and per-column bookkeeping. "
                        "cells_final_thresholds = every pair once over its full length inside the strip of the FINAL "
                        "best distance (what a scheduler that knew the answer would still touch if no pair could stop "
                        "early and no cell could be dropped)."}
    if cells_band:
        # SURVEY.md §8d's a-priori convention (the reference's own thresholds, both directions, 12 instr per
        # word-column): kept for continuity; it is not a lower bound of the work and can exceed the peak
        roofline["survey_convention"] = {"cells_band": cells_band, "int_ops_per_cell": INT_OPS_PER_CELL_SURVEY,
                                         "frac": cells_band * INT_OPS_PER_CELL_SURVEY / t_k / peak_all,
                                         "note": "not a bound of the work (exceeds 1): kept for continuity only"}
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = stats["pairs"] * (sum(len(s) for s in seqs) / max(n, 1)) / 4.0     # packed target bytes streamed
    roofline_hbm = {"bound": "hbm", "achieved": alg_bytes / (main_kernel_ms * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": alg_bytes / (main_kernel_ms * 1e-3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback of B200_PROFILING.md",
                    "traffic": ncu.get("dram_bytes_per_launch"), "note": "packed reads (%.1f MB) are L2-resident; this path is not memory-bound" % (
                        sum(len(s) for s in seqs) / 4e6)}

    line = {
        "metric": "all-vs-all NN-graph GCUPS", "value": cells_full / (step_ms * 1e-3) / 1e9, "unit": "GCUPS",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "scale": args.scale, "reads": n,
                   "l2": "flushed between steps (256 MiB memset); the 2-bit read set itself is L2-sized by design",
                   "numerator": numerator, "parallelism": "row tiles split over %d GPU(s)" % world},
        "wall_ms_per_step": step_wall_ms, "main_kernel_ms": main_kernel_ms,
        "e2e": {"value": cells_full / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int((up1["uploaded_bytes"] - up0["uploaded_bytes"]) // args.steps + 8 * (n + 1) + 14 * n),
                "d2h_bytes_per_step": int(4 * n + 12 * n_edges),
                "reads_uploaded_per_step": int((up1["uploaded_reads"] - up0["uploaded_reads"]) // args.steps),
                "note": "compute_[2set_]nearest_neighbor_graph(host dicts): device store emptied before every call, so all "
                        "reads are gathered, copied host->device and packed inside the timer (round 1 of a pipeline run)"},
        "e2e_resident": {"value": cells_full / (e2e_warm_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_warm_ms,
                         "h2d_bytes_per_step": int((up2["uploaded_bytes"] - up1["uploaded_bytes"]) // args.steps + 14 * n),
                         "d2h_bytes_per_step": int(4 * n + 12 * n_edges),
                         "reads_uploaded_per_step": int((up2["uploaded_reads"] - up1["uploaded_reads"]) // args.steps),
                         "note": "the same call with the reads resident from the previous call (a later correction round "
                                 "uploads only the reads that changed; here none)"},
        "gpu_launches": (3 * int(stats["launches"]) + 1) * args.steps,   # resident + 2 e2e graph builds per step, + 1 pack kernel per cold e2e step
        "parity": {"full_size_graph_digest_equals_oracle": parity, "edges": n_edges},
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "clocks": clocks,
        "device_stats": stats,
    }
    if shard_timing:
        line["sharding_rank0_last_step"] = {"collective_device_ms": shard_timing.get("collective_ms"),
                                            "host_ms": shard_timing.get("host_ms")}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
