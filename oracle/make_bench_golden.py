"""oracle/make_bench_golden.py -- TEST INFRASTRUCTURE.  Full-size oracle run of a synthetic
bench workload (default: BASELINE.json configs[1], "c2") -> tests/golden/bench_<name>.json:
the implementation-independent work counters of SURVEY.md §8(d) (calls, cells_full,
cells_band at the reference's default nr_cores = 16 chunking) and the digest of the exact
graph, keyed by a fingerprint of the generated reads.  bench.py uses it as the GCUPS
numerator and as the full-size bit-exact parity check.  Usage:
    python oracle/make_bench_golden.py c2 [scale] [threads]
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from isocon_b200 import workloads          # noqa: E402
from oracle import oracle as O             # noqa: E402


class P(object):
    nr_cores = 16                           # IsoCon:197 default
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


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    t0 = time.time()
    if name == "c5":
        X, C = workloads.config5(scale=scale)
        lst = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
        G = O.get_exact_nearest_neighbor_graph_2set(lst, set(C), P(), _threads=threads)
        n_entries = len(lst)
        fp = fingerprint([s for s, _ in lst])
    else:
        S = workloads.CONFIGS[name](scale=scale)
        by_seq = {}
        for a, s in S.items():
            by_seq[s] = a
        lst = sorted(by_seq.items(), key=lambda e: len(e[0]))
        G = O.get_exact_nearest_neighbor_graph(lst, set(), P(), _threads=threads)
        n_entries = len(lst)
        fp = fingerprint([s for s, _ in lst])
    wall = time.time() - t0
    out = dict(generator="oracle/make_bench_golden.py", workload=name, scale=scale, entries=n_entries,
               fingerprint=fp, nr_cores=P.nr_cores, work=dict(O.LAST_STATS), digest=graph_digest(G),
               edges=sum(len(v) for v in G.values()), sum_ed=sum(d for v in G.values() for d in v.values()),
               oracle_threads=threads, oracle_wall_s=round(wall, 1))
    tag = name if scale == 1.0 else "%s_s%g" % (name, scale)
    path = os.path.join(ROOT, "tests", "golden", "bench_%s.json" % tag)
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
