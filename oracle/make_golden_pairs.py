"""oracle/make_golden_pairs.py -- fixture for the explicit-pair path (authoring container only).

Runs the UNMODIFIED reference ``/root/reference/modules/edlib_alignment_module.py`` (imported with
the edlib stand-in of oracle/edlib_shim.cpp, like oracle/reference_driver.py does for the graph
module) on pair lists shaped like the ones IsoCon builds right after a graph build:

  * ``edlib_align_sequences``: every read of n_200 against its nearest neighbours (the partition
    alignments of isocon_get_candidates.py:38), single- and 3-core;
  * ``edlib_align_sequences_keeping_accession``: every candidate against the reads assigned to it
    (isocon_statistical_test.py:289), from the 2-set golden graph.

Output: tests/golden/pairs_n200.json (accessions + distances; sequences come from c1_n200.npz).
Every distance is cross-checked against the oracle's plain DP.  Usage: python oracle/make_golden_pairs.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O                # noqa: E402
from oracle import reference_driver as rd     # noqa: E402
from isocon_b200 import workloads             # noqa: E402
import util                                   # noqa: E402


def main():
    rd.load()                                  # puts the edlib stand-in and /root/reference on sys.path
    from modules import edlib_alignment_module as ref
    assert ref.__file__.startswith(rd.REFERENCE_ROOT)
    S = util.load_reads(200)
    Sp, hc = workloads.round1_call(S)
    exp = util.c1_expected()["200"]["cases"]
    acc_of = {s: a for a, s in Sp.items()}
    # 1-set shape: sequence -> list of neighbour sequences (with a repeated partner, like a multigraph edge)
    matches = {}
    for q, nbrs in exp["1set_round1"]["graph"]:
        if nbrs:
            matches[Sp[q]] = [Sp[t] for t, _ in nbrs] + [Sp[nbrs[0][0]]]
    with rd.quiet():
        got1 = ref.edlib_align_sequences(matches, nr_cores=1)
        got3 = ref.edlib_align_sequences(matches, nr_cores=3)
    assert got1 == got3
    for s1 in got1:
        for s2, ed in got1[s1].items():
            assert ed == O.ed_plain(s1.encode(), s2.encode())
    seq_pairs = [[acc_of[s1], [[acc_of[s2], ed] for s2, ed in v.items()]] for s1, v in got1.items()]
    # 2-set shape: candidate accession -> read accession -> (candidate seq, read seq)
    X, C = util.two_set_split(S)
    by_cand = {}
    for r, nbrs in exp["2set_every17"]["graph"]:
        for c, _ in nbrs[:1]:
            by_cand.setdefault(c, {})[r] = (C[c], X[r])
    with rd.quiet():
        gotk = ref.edlib_align_sequences_keeping_accession(by_cand, nr_cores=1)
        gotk3 = ref.edlib_align_sequences_keeping_accession(by_cand, nr_cores=2)
    assert gotk == gotk3
    acc_pairs = [[c, [[r, v[2]] for r, v in rows.items()]] for c, rows in gotk.items()]
    out = {"source": "unmodified reference modules/edlib_alignment_module.py + edlib stand-in",
           "edlib_align_sequences": seq_pairs, "edlib_align_sequences_keeping_accession": acc_pairs}
    path = os.path.join(ROOT, "tests", "golden", "pairs_n200.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, sum(len(v) for _, v in seq_pairs), "+", sum(len(v) for _, v in acc_pairs), "pairs")


if __name__ == "__main__":
    main()
