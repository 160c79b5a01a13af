#!/usr/bin/env python
"""bench.py -- all-vs-all nearest-neighbour-graph throughput on B200 (BASELINE.json metric).

One "step" = one complete 1-set NN-graph build over the workload (default: BASELINE.json
configs[1] = "c2": synthetic 10k reads x 1.5 kb, 20 near-identical gene copies, 5 % error).

  value      GCUPS = cells_full / (device time of exactly K resident steps / K): graph_begin + SEED/MAIN/WIDE
             + tie filter (+ the NCCL reductions at N > 1), packed reads already in HBM, bracketed by
             barrier + synchronize, CUDA events on the library's stream, max over ranks.
             cells_full = sum of len(q)*len(t) over every pair the REFERENCE hands to edlib on this
             input (oracle counter, tests/golden/bench_<workload>.json) -- the conventional,
             implementation-independent GCUPS numerator.
  e2e        the same numerator over the wall time of compute_nearest_neighbor_graph(S, ...) called
             with host dicts: 2-bit packing + H2D + kernels + D2H + dict rebuild inside the timer.
  roofline   dominant kernel (MAIN-phase tile kernel): cells_band * 0.375 int-ops/cell / its
             CUDA-event duration vs the INT32 ALU issue rate measured by the library's probe kernel
             on this very GPU (this path is integer bit-parallel DP: not HBM-, not tensor-bound).
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
ALU_INSTR_PER_WORD_COLUMN = 9.58  # measured in the SASS of the unrolled body at W = 5 (766 ALU-pipe instr / 80 word-columns)
WORKLOAD_DESC = {
    "c2": "c2: synthetic 10k reads x 1.5 kb, 20 near-identical gene copies, 5% indel-heavy error, 1-set all-vs-all NN graph",
    "c3": "c3: synthetic 50k Iso-Seq-like reads x 3 kb, 100 paralogs 0.5-2% apart, 2% error, 1-set NN graph",
    "c4": "c4: synthetic 200k ONT-like amplicon reads x 1 kb, 10% error, 1-set NN graph",
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


def load_workload(name, scale):
    from isocon_b200 import workloads
    S = workloads.CONFIGS[name](scale=scale)
    by_seq = {}
    for a, s in S.items():
        by_seq[s] = a
    lst = sorted(by_seq.items(), key=lambda e: len(e[0]))
    tag = name if scale == 1.0 else "%s_s%g" % (name, scale)
    gold_path = os.path.join(ROOT, "tests", "golden", "bench_%s.json" % tag)
    gold = None
    if os.path.exists(gold_path):
        with open(gold_path) as fh:
            gold = json.load(fh)
        if gold.get("fingerprint") != fingerprint([s for s, _ in lst]):
            gold = None     # generator drift: the stored counters do not describe this input
    return S, lst, gold


# ----------------------------------------------------------------------------- CPU side (oracle)

def cpu_sample(lst, n_queries, threads):
    """Oracle port on host cores: `n_queries` evenly spaced queries of the workload, each scanned
    against the whole list exactly as the reference does (no cross-query seeding), one query per
    worker call, `threads` worker threads.  Returns (cells_full, cells_band, calls, wall_s)."""
    from oracle import oracle as O
    n = len(lst)
    n_queries = max(1, min(n_queries, n))
    qs = np.unique(np.linspace(0, n - 1, n_queries).astype(np.int64))
    cat, off = O.concat([s for s, _ in lst])
    conv = np.zeros(max(n, 1), np.uint8)
    L = O.lib()
    import concurrent.futures as cf
    tot = np.zeros(4, dtype=np.uint64)

    def one(q):
        best = np.full(n, -1, np.int32)
        st = np.zeros(4, np.uint64)
        cap = 4096
        eq = np.empty(cap, np.int32); et = np.empty(cap, np.int32); ed = np.empty(cap, np.int32)
        L.nn_oracle_1set(cat, off, n, conv, 2 ** 32, int(q), 1, 1, 1, 0, best, eq, et, ed, cap, st)   # releases the GIL
        return st

    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        for st in ex.map(one, qs.tolist()):
            tot += st
    wall = time.perf_counter() - t0
    return int(tot[2]), int(tot[3]), int(tot[0]), wall, int(qs.size)


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, lst, gold = load_workload(args.workload, args.scale)
    threads = host_threads()
    nq = args.cpu_queries or max(threads * 8, 64)
    for _ in range(args.warmup):
        cpu_sample(lst, max(threads, 8), threads)
    cells = wall = 0.0
    for _ in range(args.steps):
        cf_, cb_, calls, w, used = cpu_sample(lst, nq, threads)
        cells += cf_; wall += w
    gcups = cells / wall / 1e9
    sample = "%d of %d queries (evenly spaced) x the whole list per step, %d threads" % (used, len(lst), threads)
    line = {
        "impl": "reference", "metric": "all-vs-all NN-graph GCUPS", "value": gcups, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "scale": args.scale, "reads": len(lst)},
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port", "sample": sample,
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
    ap.add_argument("--cpu-queries", type=int, default=0, help="queries in the CPU-baseline sample (0 = 2 x threads)")
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

    S, lst, gold = load_workload(args.workload, args.scale)
    n = len(lst)
    seqs = [s for s, _ in lst]
    ctx = _binding.get_context(local_rank)
    int32_peak = ctx.int32_peak()
    ctx.set_reads(seqs)
    isq = np.ones(n, np.uint8)
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
            ctx.graph_begin(1, 2 ** 32, isq, None)
            ctx.graph_run(_binding.PHASE_ALL)
            ctx.graph_finalize()
        else:
            sharding.run_sharded(sharding.CudaShardOps(ctx, 1, 2 ** 32, isq, None), dist, timing=shard_timing)
        return ctx.last_ms(5)

    def e2e_step():
        flush_buf.zero_()
        torch.cuda.synchronize()
        return nn.compute_nearest_neighbor_graph(S, set(), Params())[0]

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
    # ---- timed region 2: K end-to-end calls of the reference-facing function with host dicts
    t0 = time.perf_counter()
    for _ in range(args.steps):
        with quiet:
            e2e_step()
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    main_kernel_ms = sum(main_ms) / len(main_ms)
    if dist is not None:
        t = torch.tensor([step_ms, e2e_ms, main_kernel_ms, step_wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms, main_kernel_ms, step_wall_ms = [float(x) for x in t.tolist()]
        keys = ["pairs", "word_columns", "groups", "items", "edges_raw", "launches"]
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
        lens = np.array([len(s) for s in seqs], dtype=np.float64)
        cells_full = float(lens.sum() ** 2 - (lens ** 2).sum())       # every ordered pair (lengths within the bound)
        cells_band = None
        numerator = "all ordered pairs (no oracle counters stored for this input)"

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = host_threads()
        nq = args.cpu_queries or max(threads * 8, 64)
        cf_, cb_, calls, wall, used = cpu_sample(lst, nq, threads)
        cpu = {"value": cf_ / wall / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
               "sample": "%d of %d queries (evenly spaced) x the whole list, %.1f s wall, %d edit-distance calls" % (
                   used, n, wall, calls),
               "cells_band_per_cells_full": cb_ / cf_}
        if cells_band is None:
            cells_band = cb_ * (n / used)     # extrapolated from the sample
            numerator += "; cells_band extrapolated from the CPU sample"

    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_ncu.json")) as fh:
            ncu = json.load(fh).get(args.workload if args.scale == 1.0 else "", {})
    except Exception:
        pass
    bytes_in = sum(len(s) for s in seqs) + 8 * (n + 1)
    n_edges = sum(len(v) for v in G.values())
    roofline = None
    if cells_band:
        # algorithmic work: every UNORDERED pair once (the 1-set graph is symmetric: d(q,t) = d(t,q)), inside
        # the Ukkonen strip of the reference's own threshold, at the instruction count of the recurrence
        t_k = main_kernel_ms * 1e-3
        alg_ops = 0.5 * cells_band * INT_OPS_PER_CELL
        achieved = alg_ops / t_k / 1e12
        executed = stats["word_columns"] * 32 * ALU_INSTR_PER_WORD_COLUMN / t_k / 1e12
        roofline = {"bound": "int32", "kernel": "nn_row_kernel (the PILOT + MAIN [+ WIDE] launches of one step, summed)",
                    "achieved": achieved, "peak": int32_peak / 1e12, "unit": "Tint-op/s", "frac": achieved / (int32_peak / 1e12),
                    "peak_source": "measured on this GPU by isocon_nn_int32_peak (LOP3/IADD3 probe kernel)",
                    "algorithmic_ops": alg_ops, "cells_band": cells_band, "cells_band_unordered": 0.5 * cells_band,
                    "int_ops_per_cell": INT_OPS_PER_CELL, "kernel_ms": main_kernel_ms,
                    "executed_lane_word_columns": stats["word_columns"] * 32,
                    "executed_alu_ops_frac_of_peak": executed / (int32_peak / 1e12),
                    "frac_survey_convention": cells_band * INT_OPS_PER_CELL_SURVEY / t_k / int32_peak,
                    "ncu_alu_pipe_pct_of_peak": ncu.get("alu_pipe_pct"), "ncu_source": ncu.get("source"),
                    "traffic": ncu.get("dram_bytes_per_launch"),
                    "note": "frac = algorithmically necessary integer ops / measured INT32 issue peak; the gap to "
                            "executed_alu_ops_frac_of_peak (cross-checked by ncu sm__inst_executed_pipe_alu) is 32-bit word "
                            "granularity of the band, lanes waiting for the slowest pair of their warp, and per-column "
                            "bookkeeping. frac_survey_convention (both directions, 12 instr/word-column) exceeds 1 and is "
                            "kept only for continuity with SURVEY.md §8d"}
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
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "scale": args.scale, "reads": n,
                   "l2": "flushed between steps (256 MiB memset); the 2-bit read set itself is L2-sized by design",
                   "numerator": numerator, "parallelism": "row tiles split over %d GPU(s)" % world},
        "wall_ms_per_step": step_wall_ms, "main_kernel_ms": main_kernel_ms,
        "e2e": {"value": cells_full / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(bytes_in + 2 * n), "d2h_bytes_per_step": int(4 * n + 12 * n_edges)},
        "gpu_launches": (2 * int(stats["launches"]) + 1) * args.steps,   # resident + e2e graph builds, + 1 pack kernel per e2e step
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
