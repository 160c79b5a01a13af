#!/usr/bin/env python
"""Workload of the compute-sanitizer runs (tools/r02c.sh): every kernel of the library on small inputs, through the
reference-facing calls, checked against the committed known answers.  Run as
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize.py [--quick | --c5]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import util  # noqa: E402
from isocon_b200 import _binding, workloads  # noqa: E402
from isocon_b200 import nearest_neighbor_graph as nn  # noqa: E402
from isocon_b200 import edlib_alignment_module as em  # noqa: E402


def two_level():
    """The 2-set graph over clustered candidates (min-hash sketch, hints, sampled cap, swapped SEED launch, q-gram
    filter, level 1 by alignment, level 2) on c5 at scale 0.06, against the oracle."""
    from oracle import oracle as O
    X, C = workloads.config5(scale=0.06)
    G = nn.compute_2set_nearest_neighbor_graph(X, C, util.Params())
    st = _binding.get_context().stats()
    util.assert_same_graph(G, O.compute_2set_nearest_neighbor_graph(X, C, util.Params(nr_cores=4)), "c5 at 0.06")
    assert st["clusters"] > 0 and st["main_passes"] >= 1
    print("sanitize workload ok: two-level 2-set graph, stats %s" % st)


def main():
    if "--c5" in sys.argv:
        return two_level()
    quick = "--quick" in sys.argv
    ctx = _binding.get_context()
    n_cases = 0
    cases = util.known_answers() + util.known_answers(foreign=True)
    for case in cases[::6] if quick else cases:          # K cases, random cases, depth <= 0, foreign symbols, ties
        if _is_foreign(case):
            ctx.store_reset(); ctx._slot_of = None
        util.assert_same_graph(util.run_case(nn, case), case["graph"], case["name"])
        n_cases += 1
    # the row kernel (PILOT + binned MAIN), the one-sided ladder, the scan emulation, explicit pairs, regrowth
    S = util.load_reads(200)
    exp = util.c1_expected()["200"]["cases"]
    Sp, hc = workloads.round1_call(S)
    G, _ = nn.compute_nearest_neighbor_graph(Sp, hc, util.Params())
    util.assert_same_graph(G, exp["1set_round1"]["graph"], "n_200 1-set")
    X, C = util.two_set_split(S)
    util.assert_same_graph(nn.compute_2set_nearest_neighbor_graph(X, C, util.Params()), exp["2set_every17"]["graph"], "2-set")
    util.assert_same_graph(nn.compute_2set_nearest_neighbor_graph(X, C, util.Params(neighbor_search_depth=3)),
                           exp["2set_every17_depth3"]["graph"], "2-set depth 3")
    ctx.reserve_edges(-8)
    G, _ = nn.compute_nearest_neighbor_graph(Sp, hc, util.Params())
    ctx.reserve_edges(0)
    util.assert_same_graph(G, exp["1set_round1"]["graph"], "n_200 1-set after regrowth")
    seqs = list(Sp.values())[:40]
    got = em.edlib_align_sequences({seqs[i]: [seqs[(i + 1) % 40], seqs[(i + 7) % 40]] for i in range(40)})
    assert len(got) == 40
    if not quick:
        T = workloads.config2(scale=0.03)
        G, _ = nn.compute_nearest_neighbor_graph(T, set(), util.Params())
        assert sum(len(v) for v in G.values()) > 0
    print("sanitize workload ok: %d known-answer cases + n_200 graphs, stats %s" % (n_cases, ctx.stats()))


def _is_foreign(case):
    return util._is_foreign_case(case)


if __name__ == "__main__":
    main()
