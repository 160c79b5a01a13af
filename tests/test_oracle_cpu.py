"""CPU suite, part 1: the oracle against the golden vectors (tests/golden/, produced from the
unmodified reference driver by oracle/make_golden.py) and against itself (three independent
edit-distance implementations)."""
import os

import numpy as np
import pytest

import util
from isocon_b200 import workloads
from oracle import oracle as O
from oracle import reference_driver as rd


@pytest.mark.parametrize("case", util.known_answers() + util.known_answers(foreign=True), ids=lambda c: c["name"])
def test_known_answers(case):
    util.assert_same_graph(util.run_case(O, case), case["graph"], case["name"])


def test_correction_round_sequence():
    """Three consecutive rounds (oracle/make_golden_r2.py): the call graphs.py:37-58 makes and a 2-set call."""
    for k, r in enumerate(util.correction_rounds()):
        Sp, hc = workloads.round1_call(r["S"])
        G, _ = O.compute_nearest_neighbor_graph(Sp, hc, util.Params())
        util.assert_same_graph(G, r["graph_1set"], "round %d 1-set" % k)
        util.assert_same_graph(O.compute_2set_nearest_neighbor_graph(r["S"], r["C"], util.Params()), r["graph_2set"],
                               "round %d 2-set" % k)


@pytest.mark.parametrize("n", [200, 500])
def test_fasta_fixture_graphs(n):
    S = util.load_reads(n)
    exp = util.c1_expected()[str(n)]["cases"]
    Sp, hc = workloads.round1_call(S)
    G, iso = O.compute_nearest_neighbor_graph(Sp, hc, util.Params())
    util.assert_same_graph(G, exp["1set_round1"]["graph"], "1set_round1")
    assert O.LAST_STATS == exp["1set_round1"]["work"]       # same calls / cells as the reference run
    G = O.compute_nearest_neighbor_graph(Sp, hc, util.Params(nr_cores=3))[0]
    util.assert_same_graph(G, exp["1set_round1_cores3"]["graph"], "cores3")
    for depth in (1, 5, 20):
        G = O.compute_nearest_neighbor_graph(Sp, hc, util.Params(neighbor_search_depth=depth))[0]
        util.assert_same_graph(G, exp["1set_depth%d" % depth]["graph"], "depth%d" % depth)
    X, C = util.two_set_split(S)
    util.assert_same_graph(O.compute_2set_nearest_neighbor_graph(X, C, util.Params()), exp["2set_every17"]["graph"])
    for depth in (1, 3, 10):
        G = O.compute_2set_nearest_neighbor_graph(X, C, util.Params(neighbor_search_depth=depth))
        util.assert_same_graph(G, exp["2set_every17_depth%d" % depth]["graph"], "2set depth%d" % depth)


def test_three_edit_distances_agree():
    rng = np.random.default_rng(11)
    for _ in range(300):
        L = int(rng.integers(1, 300))
        tpl = rng.integers(0, 4, size=L, dtype=np.uint8)
        err = float(rng.choice([0.0, 0.03, 0.12, 0.4]))
        a = workloads._mutate(rng, tpl, err / 3, err / 3, err / 3)
        b = workloads._mutate(rng, tpl, err / 3, err / 3, err / 3)
        if a.size == 0 or b.size == 0:
            continue
        x, y = workloads._to_str(a), workloads._to_str(b)
        d = O.ed_plain(x, y)
        assert O.ed_myers64(x, y, -1) == d
        for k in {0, max(d - 1, 0), d, d + 1, d + 50, max(len(x), len(y))}:
            want = d if d <= k else -1
            assert O.ed_myers64(x, y, k) == want
            assert O.ed_banded_dp(x, y, k) == want


def test_exact_characters_like_edlib():
    # characters are compared exactly (the oracle works on raw bytes)
    assert O.ed_myers64("ACGTNCGT", "ACGTnCGT", 8) == 1
    assert O.ed_plain("ACGTNCGT", "ACGTACGT") == 1
    assert O.ed_myers64("", "ACG", 5) == 3 and O.ed_myers64("ACG", "", 2) == -1


def test_allpairs_fixture_matches_plain_dp_sample():
    z = np.load(os.path.join(util.GOLD, "c1_n200_allpairs.npz"))
    S = util.load_reads(200)
    seqs = [S[a] for a in z["acc"].tolist()]
    rng = np.random.default_rng(3)
    for p in rng.choice(z["a"].size, size=60, replace=False):
        assert O.ed_plain(seqs[z["a"][p]], seqs[z["b"][p]]) == z["ed"][p]


@pytest.mark.skipif(not rd.available(), reason="reference tree / edlib stand-in only exist in the authoring container")
def test_oracle_equals_unmodified_reference_driver():
    ref, _ = rd.load()
    S = util.load_reads(200)
    Sp, hc = workloads.round1_call(S)
    for kw in (dict(), dict(nr_cores=2), dict(neighbor_search_depth=3)):
        with rd.quiet():
            G, _ = ref.compute_nearest_neighbor_graph(Sp, hc, rd.Params(**kw))
        util.assert_same_graph(O.compute_nearest_neighbor_graph(Sp, hc, util.Params(**kw))[0], G)
    X, C = util.two_set_split(S)
    for kw in (dict(), dict(neighbor_search_depth=2)):
        with rd.quiet():
            G = ref.compute_2set_nearest_neighbor_graph(X, C, rd.Params(**kw))
        util.assert_same_graph(O.compute_2set_nearest_neighbor_graph(X, C, util.Params(**kw)), G)
