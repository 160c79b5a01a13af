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


@pytest.mark.skipif(not rd.available(), reason="needs /root/reference (authoring container)")
def test_consumers_one_level_up_see_the_same_partitions(monkeypatch):
    """the only callers of the path are graphs.py:58 / :154, reached from
    partitions.partition_strings / partition_strings_2set.  Run the UNMODIFIED reference consumers once on the
    reference's own nearest_neighbor_graph module and once after isocon_b200.install() shadowed it: same G_star edges,
    same partition, same centres.  No GPU in this container, so the replacement's device seam (_build_graph) is stood
    in by the oracle here -- what is tested is the wiring (install() reaches graphs.py, the consumers accept the
    replacement's dicts); the device path itself is compared dict by dict in the GPU suite."""
    import sys
    import types
    import networkx
    import isocon_b200
    from isocon_b200 import nearest_neighbor_graph as nn
    ref_nn, _ = rd.load()                               # puts /root/reference and the edlib stand-in on sys.path
    if not hasattr(networkx.Graph, "node"):             # the reference needs networkx <= 2.3 (G.node[...])
        monkeypatch.setattr(networkx.Graph, "node", property(lambda self: self._node), raising=False)
    for name in ("parasail", "pysam"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from modules import graphs, partitions

    def oracle_build(seqs, accs, lens, mode, is_query, is_target, depth, lo, hi):
        L = list(zip(seqs, accs))
        if mode == 1:
            hc = {s for s, q in zip(seqs, is_query) if not q}
            return O.get_nearest_neighbors(L[lo:hi], 0, lo, L, hc, depth)
        return O.get_nearest_neighbors_2set(L[lo:hi], lo, L, {a for a, t in zip(accs, is_target) if t}, depth)

    monkeypatch.setattr(nn, "_build_graph", oracle_build)
    S = util.load_reads(200)
    P = rd.Params()
    X, C = util.two_set_split(S)

    def run():
        with rd.quiet():
            G_star, graph_partition, M, converged = partitions.partition_strings(S, P)
            G2, part2 = partitions.partition_strings_2set(X, C, "x.fa", "c.fa", P)
        return (sorted(G_star.edges(data="edit_distance")), {k: sorted(v) for k, v in graph_partition.items()}, sorted(M),
                converged, sorted(G2.edges(data=True), key=lambda e: e[:2]), {k: sorted(v) for k, v in part2.items()})

    assert graphs.nearest_neighbor_graph is ref_nn
    want = run()
    try:
        assert isocon_b200.install() is nn
        assert graphs.nearest_neighbor_graph is nn
        got = run()
    finally:                                            # put the reference module back for the other tests
        sys.modules["modules.nearest_neighbor_graph"] = ref_nn
        graphs.nearest_neighbor_graph = ref_nn
        import modules
        modules.nearest_neighbor_graph = ref_nn
        ref_pairs = sys.modules.get("modules.edlib_alignment_module")
        if ref_pairs is not None and ref_pairs.__name__.startswith("isocon_b200"):
            del sys.modules["modules.edlib_alignment_module"]
    assert got == want
    assert len(want[0]) > 200 and len(want[1]) > 10 and len(want[5]) > 3
