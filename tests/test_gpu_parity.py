"""GPU suite: the CUDA path, called through the reference-facing module (and so through the
C ABI), against the golden vectors of the unmodified reference driver and against the oracle.
Bit-exact: identical dicts, identical key order, identical insertion order."""
import os

import numpy as np
import pytest

import util
from isocon_b200 import _binding, workloads
from isocon_b200 import nearest_neighbor_graph as nn
from oracle import oracle as O


_C5_SMALL = {}


def _c5_small():
    """c5 at scale 0.06 with the oracle's graph (computed once per session: 20 s)."""
    if not _C5_SMALL:
        X, C = workloads.config5(scale=0.06)              # 6000 reads x 300 candidates of 30 families
        _C5_SMALL["v"] = (X, C, O.compute_2set_nearest_neighbor_graph(X, C, util.Params(nr_cores=4)))
    return _C5_SMALL["v"]


pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _binding.NNContext(0)


def _sorted_list_1set(S):
    by_seq = {}
    for acc, seq in S.items():
        by_seq[seq] = acc
    return sorted(by_seq.items(), key=lambda e: len(e[0]))


def _graph_via_ctx(ctx, L, mode, depth, is_query, is_target, algo, symmetric):
    ctx.set_reads([s for s, _ in L])
    best, eq, et, ed = ctx.graph(mode, depth, is_query, is_target, algo, symmetric)
    eq, et, ed = nn._order_edges(eq, et, ed)
    out = {}
    for i in range(len(L)):
        if mode == 1 or not is_target[i]:
            out[L[i][1]] = {}
    for q, t, d in zip(eq.tolist(), et.tolist(), ed.tolist()):
        out[L[q][1]][L[t][1]] = d
    return out


@pytest.mark.parametrize("case", util.known_answers(), ids=lambda c: c["name"])
def test_known_answers(case):
    util.assert_same_graph(util.run_case(nn, case), case["graph"], case["name"])


@pytest.mark.parametrize("n", [200, 500, 1000, 2000])
def test_fasta_fixtures_all_golden_cases(n):
    S = util.load_reads(n)
    exp = util.c1_expected()[str(n)]["cases"]
    Sp, hc = workloads.round1_call(S)
    X, C = util.two_set_split(S)
    for name, e in exp.items():
        kw = {}
        if "cores3" in name:
            kw["nr_cores"] = 3
        if "depth" in name:
            kw["neighbor_search_depth"] = int(name.split("depth")[1])
        P = util.Params(**kw)
        if name.startswith("1set"):
            G, iso = nn.compute_nearest_neighbor_graph(Sp, set() if name == "1set_noconv" else hc, P)
            assert iso == set()
        else:
            G = nn.compute_2set_nearest_neighbor_graph(X, C, P)
        util.assert_same_graph(G, e["graph"], "n_%d %s" % (n, name))


@pytest.mark.parametrize("algo,symmetric", [(_binding.ALGO_TILE, True), (_binding.ALGO_TILE, False),
                                            (_binding.ALGO_SCAN, False)])
def test_every_algorithm_variant_gives_the_same_graph(ctx, algo, symmetric):
    S = util.load_reads(500)
    exp = util.c1_expected()["500"]["cases"]
    Sp, hc = workloads.round1_call(S)
    L = _sorted_list_1set(Sp)
    isq = np.array([0 if s in hc else 1 for s, _ in L], dtype=np.uint8)
    G = _graph_via_ctx(ctx, L, 1, 2 ** 32, isq, None, algo, symmetric)
    util.assert_same_graph(G, exp["1set_round1"]["graph"], "1-set algo %d sym %d" % (algo, symmetric))
    G = _graph_via_ctx(ctx, L, 1, 5, isq, None, algo, symmetric)
    util.assert_same_graph(G, exp["1set_depth5"]["graph"], "1-set depth 5 algo %d sym %d" % (algo, symmetric))
    X, C = util.two_set_split(S)
    L2 = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
    ist = np.array([1 if a in C else 0 for _, a in L2], dtype=np.uint8)
    G = _graph_via_ctx(ctx, L2, 2, 2 ** 32, 1 - ist, ist, algo, False)
    util.assert_same_graph(G, exp["2set_every17"]["graph"], "2-set algo %d" % algo)


def test_small_register_band_limit_forces_the_wide_phase(ctx, monkeypatch):
    # with a tiny MAIN threshold most queries stay unresolved and go through the WIDE phase
    monkeypatch.setenv("ISOCON_NN_KCAP", "20")
    c2 = _binding.NNContext(0)
    S = util.load_reads(200)
    exp = util.c1_expected()["200"]["cases"]
    Sp, hc = workloads.round1_call(S)
    L = _sorted_list_1set(Sp)
    isq = np.array([0 if s in hc else 1 for s, _ in L], dtype=np.uint8)
    G = _graph_via_ctx(c2, L, 1, 2 ** 32, isq, None, _binding.ALGO_TILE, True)
    util.assert_same_graph(G, exp["1set_round1"]["graph"], "kcap 20")
    assert c2.stats()["wide_pairs"] > 0
    c2.close()


@pytest.mark.parametrize("row_kernel", ["0", "1"])
def test_main_phase_kernels_agree(monkeypatch, row_kernel):
    # MAIN runs on the diagonal-band row kernel (one query per block) when the mask table fits in
    # shared memory, else on the block-band tile kernel; both must give the golden graphs
    monkeypatch.setenv("ISOCON_NN_ROW_KERNEL", row_kernel)
    c = _binding.NNContext(0)
    for n in (200, 1000):
        S = util.load_reads(n)
        exp = util.c1_expected()[str(n)]["cases"]
        Sp, hc = workloads.round1_call(S)
        L = _sorted_list_1set(Sp)
        isq = np.array([0 if s in hc else 1 for s, _ in L], dtype=np.uint8)
        for sym in (True, False):
            G = _graph_via_ctx(c, L, 1, 2 ** 32, isq, None, _binding.ALGO_TILE, sym)
            util.assert_same_graph(G, exp["1set_round1"]["graph"], "row_kernel %s sym %d" % (row_kernel, sym))
        X, C = util.two_set_split(S)
        L2 = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
        ist = np.array([1 if a in C else 0 for _, a in L2], dtype=np.uint8)
        G = _graph_via_ctx(c, L2, 2, 2 ** 32, 1 - ist, ist, _binding.ALGO_TILE, False)
        util.assert_same_graph(G, exp["2set_every17"]["graph"], "row_kernel %s 2-set" % row_kernel)
    c.close()


@pytest.mark.parametrize("kcap", ["40", "100", "448"])
def test_row_kernel_with_other_band_limits(monkeypatch, kcap):
    # the mask-table padding follows the MAIN threshold; small limits push rows into the WIDE phase
    monkeypatch.setenv("ISOCON_NN_KCAP", kcap)
    c = _binding.NNContext(0)
    S = workloads.config2(scale=0.03)
    L = _sorted_list_1set(S)
    P = util.Params()
    want, _ = O.compute_nearest_neighbor_graph(S, set(), P)
    G = _graph_via_ctx(c, L, 1, 2 ** 32, np.ones(len(L), np.uint8), None, _binding.ALGO_TILE, True)
    util.assert_same_graph(G, want, "kcap %s" % kcap)
    c.close()


@pytest.mark.parametrize("first", ["3", "17", "40"])
def test_one_sided_threshold_ladder_with_many_passes(monkeypatch, first):
    # one-sided graphs (2-set, symmetric off) redo the rows that are unresolved at a cap with the next cap
    # (first, 2 first + 1, ...): forced small first caps give up to seven passes; the graphs must not change
    monkeypatch.setenv("ISOCON_NN_LADDER_FIRST", first)
    c = _binding.NNContext(0)
    passes = []
    for n in (200, 1000):
        S = util.load_reads(n)
        exp = util.c1_expected()[str(n)]["cases"]
        X, C = util.two_set_split(S)
        L2 = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
        ist = np.array([1 if a in C else 0 for _, a in L2], dtype=np.uint8)
        G = _graph_via_ctx(c, L2, 2, 2 ** 32, 1 - ist, ist, _binding.ALGO_TILE, False)
        util.assert_same_graph(G, exp["2set_every17"]["graph"], "ladder first %s 2-set n_%d" % (first, n))
        passes.append(c.stats()["main_passes"])
        Sp, hc = workloads.round1_call(S)
        L = _sorted_list_1set(Sp)
        isq = np.array([0 if s in hc else 1 for s, _ in L], dtype=np.uint8)
        G = _graph_via_ctx(c, L, 1, 2 ** 32, isq, None, _binding.ALGO_TILE, False)     # one-sided 1-set
        util.assert_same_graph(G, exp["1set_round1"]["graph"], "ladder first %s one-sided 1-set n_%d" % (first, n))
        passes.append(c.stats()["main_passes"])
    assert max(passes) >= 3, passes
    X, C = workloads.config5(scale=0.004)
    P = util.Params(nr_cores=4)
    L2 = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
    ist = np.array([1 if a in C else 0 for _, a in L2], dtype=np.uint8)
    G = _graph_via_ctx(c, L2, 2, 2 ** 32, 1 - ist, ist, _binding.ALGO_TILE, False)
    util.assert_same_graph(G, O.compute_2set_nearest_neighbor_graph(X, C, P), "c5 ladder first %s" % first)
    assert c.stats()["main_passes"] >= 2
    c.close()


def test_ed_pairs_against_all_pairs_fixture(ctx):
    z = np.load(os.path.join(util.GOLD, "c1_n200_allpairs.npz"))
    S = util.load_reads(200)
    seqs = [S[a] for a in z["acc"].tolist()]
    order = sorted(range(len(seqs)), key=lambda i: len(seqs[i]))
    rank = np.empty(len(seqs), dtype=np.int32)
    rank[order] = np.arange(len(seqs), dtype=np.int32)
    ctx.set_reads([seqs[i] for i in order])
    a, b, want = rank[z["a"]], rank[z["b"]], z["ed"]
    got = ctx.ed_pairs(a, b, None)                       # unbounded
    assert np.array_equal(got, want)
    got = ctx.ed_pairs(b, a, None)                       # symmetric
    assert np.array_equal(got, want)
    rng = np.random.default_rng(0)
    k = (want + rng.integers(-3, 4, size=want.size)).clip(0).astype(np.int32)
    got = ctx.ed_pairs(a, b, k)
    assert np.array_equal(got, np.where(want <= k, want, -1))


def test_edlib_ed_helper():
    assert nn.edlib_ed("ACGTACGT", "ACGTACGA", k=3) == 1
    assert nn.edlib_ed("ACGTACGT", "TTTTTTTT", k=2) == -1
    assert nn.edlib_ed("ACGTACGTACGT", "ACGT", k=-1) == 8


@pytest.mark.parametrize("name,scale", [("c2", 0.03), ("c3", 0.004), ("c4", 0.002)])
def test_synthetic_1set_configs_vs_oracle(name, scale):
    S = workloads.CONFIGS[name](scale=scale)
    P = util.Params(nr_cores=4)
    want, _ = O.compute_nearest_neighbor_graph(S, set(), P)
    got, iso = nn.compute_nearest_neighbor_graph(S, set(), P)
    assert iso == set()
    util.assert_same_graph(got, want, name)


def test_synthetic_2set_config_vs_oracle():
    X, C = workloads.config5(scale=0.004)
    P = util.Params(nr_cores=4)
    util.assert_same_graph(nn.compute_2set_nearest_neighbor_graph(X, C, P),
                           O.compute_2set_nearest_neighbor_graph(X, C, P), "c5")
    P = util.Params(nr_cores=4, neighbor_search_depth=4)
    util.assert_same_graph(nn.compute_2set_nearest_neighbor_graph(X, C, P),
                           O.compute_2set_nearest_neighbor_graph(X, C, P), "c5 depth 4")


def test_later_round_shape_many_converged_reads():
    # correction rounds: many duplicates (-> has_converged) and tiny distances
    rng = np.random.default_rng(21)
    tpl = [rng.integers(0, 4, size=900, dtype=np.uint8) for _ in range(6)]
    S = {}
    for i in range(400):
        t = tpl[int(rng.integers(0, 6))]
        r = workloads._mutate(rng, t, 0.0005, 0.0005, 0.0005)
        S["r%d" % i] = workloads._to_str(r)
    Sp, hc = workloads.round1_call(S)
    assert hc
    P = util.Params()
    want, _ = O.compute_nearest_neighbor_graph(Sp, hc, P)
    got, _ = nn.compute_nearest_neighbor_graph(Sp, hc, P)
    util.assert_same_graph(got, want, "later round")


def test_query_subrange_like_a_pool_worker():
    S = util.load_reads(200)
    Sp, hc = workloads.round1_call(S)
    L = _sorted_list_1set(Sp)
    want = O.get_nearest_neighbors(L[40:75], 40, 40, L, hc, 2 ** 32)
    got = nn.get_nearest_neighbors(L[40:75], 40, 40, L, hc, 2 ** 32)
    util.assert_same_graph(got, want, "sub-range")
    X, C = util.two_set_split(S)
    L2 = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
    want = O.get_nearest_neighbors_2set(L2[10:90], 10, L2, set(C), 2 ** 32)
    got = nn.get_nearest_neighbors_2set(L2[10:90], 10, L2, set(C), 2 ** 32)
    util.assert_same_graph(got, want, "2-set sub-range")


@pytest.mark.parametrize("case", util.known_answers(foreign=True), ids=lambda c: c["name"])
def test_symbols_beyond_acgt(case):
    """edlib compares raw characters (K9: A, N, n are three symbols).  Reads over other four-symbol alphabets are
    packed with that alphabet; reads with further symbols go through the general-alphabet path."""
    _binding.get_context().store_reset()          # the alphabet is chosen per store from the first reads
    _binding.get_context()._slot_of = None
    util.assert_same_graph(util.run_case(nn, case), case["graph"], case["name"])


def _with_foreign(S, rate, seed):
    rng = np.random.default_rng(seed)
    out = {}
    for a, s in S.items():
        if rng.random() < rate:
            b = bytearray(s.encode())
            for p in rng.choice(len(b), size=max(1, len(b) // 50), replace=False):
                b[p] = ord("Nn*"[int(rng.integers(0, 3))])
            s = b.decode()
        out[a] = s
    return out


@pytest.mark.parametrize("algo", ["0", "2"])
def test_foreign_reads_inside_ordinary_inputs(monkeypatch, algo):
    """3 % of the reads carry N / n / * (2 % of their positions): 1-set (symmetric, finite depth), 2-set (ladder,
    finite depth = scan emulation) and explicit pairs, against the oracle; the other reads stay on the 2-bit path."""
    monkeypatch.setenv("ISOCON_NN_ALGO", algo)
    S = _with_foreign(workloads.config2(scale=0.04), 0.03, 1)
    ctx = _binding.get_context()
    ctx.store_reset()
    for depth in (2 ** 32, 6):
        P = util.Params(nr_cores=4, neighbor_search_depth=depth)
        want, _ = O.compute_nearest_neighbor_graph(S, set(), P)
        got, _ = nn.compute_nearest_neighbor_graph(S, set(), P)
        util.assert_same_graph(got, want, "1-set with foreign reads, depth %d" % depth)
    assert ctx.store_info()["foreign_reads"] > 0
    X, C = workloads.config5(scale=0.004)
    X = _with_foreign(X, 0.03, 2); C = _with_foreign(C, 0.1, 3)
    for depth in (2 ** 32, 4):
        P = util.Params(nr_cores=4, neighbor_search_depth=depth)
        want = O.compute_2set_nearest_neighbor_graph(X, C, P)
        got = nn.compute_2set_nearest_neighbor_graph(X, C, P)
        util.assert_same_graph(got, want, "2-set with foreign reads, depth %d" % depth)
    from isocon_b200 import edlib_alignment_module as em
    seqs = list(S.values())[:60]
    matches = {seqs[i]: [seqs[(i * 7 + j) % 60] for j in range(1, 4)] for i in range(60)}
    got = em.edlib_align_sequences(matches)
    for s1, nb in matches.items():
        for s2 in nb:
            assert got[s1][s2] == O.ed_plain(s1, s2)
    assert nn.edlib_ed("ACGTNCGT", "ACGTnCGT", k=3) == 1 and nn.edlib_ed("ACNT", "ACGTTTTT", k=2) == -1


def test_errors_are_loud():
    with pytest.raises(ValueError):
        nn.compute_nearest_neighbor_graph({"a": "ACGT", "b": "ACG\u00e9"}, set(), util.Params())    # not ASCII
    with pytest.raises(_binding.IsoconNNError):
        nn.get_nearest_neighbors([("ACGTA", "a"), ("ACG", "b")], 0, 0, [("ACGTA", "a"), ("ACG", "b")], set(), 2 ** 32)
    with pytest.raises(ZeroDivisionError):      # reference behaviour with verbose and no edges (:291)
        nn.compute_nearest_neighbor_graph({"a": "ACGT"}, set(), util.Params(verbose=True))


def test_empty_and_degenerate_inputs():
    assert nn.compute_nearest_neighbor_graph({}, set(), util.Params()) == ({}, set())
    assert nn.compute_2set_nearest_neighbor_graph({"r": "ACGT"}, {}, util.Params()) == {"r": {}}
    assert nn.compute_2set_nearest_neighbor_graph({}, {"c": "ACGT"}, util.Params()) == {}
    want = O.get_nearest_neighbors([("", "e"), ("A", "a"), ("AC", "b")], 0, 0, [("", "e"), ("A", "a"), ("AC", "b")], set(), 2 ** 32)
    got = nn.get_nearest_neighbors([("", "e"), ("A", "a"), ("AC", "b")], 0, 0, [("", "e"), ("A", "a"), ("AC", "b")], set(), 2 ** 32)
    util.assert_same_graph(got, want, "empty read")


def test_install_shadows_the_reference_module():
    import sys
    import types
    import isocon_b200
    pkg = types.ModuleType("fake_isocon_modules")
    pkg.__path__ = []
    sys.modules["fake_isocon_modules"] = pkg
    rep = isocon_b200.install("fake_isocon_modules")
    from fake_isocon_modules import nearest_neighbor_graph as shadow
    assert shadow is rep is nn


def test_full_size_properties_c2_slice():
    # size-independent properties on a larger instance: symmetry of reported distances and
    # agreement of the two algorithms' best[] (SCAN emulates the scan, TILE the closed form)
    S = workloads.config2(scale=0.2)
    L = _sorted_list_1set(S)
    c = _binding.NNContext(0)
    c.set_reads([s for s, _ in L])
    isq = np.ones(len(L), np.uint8)
    b1, q1, t1, d1 = c.graph(1, 2 ** 32, isq, None, _binding.ALGO_TILE, True)
    b2, q2, t2, d2 = c.graph(1, 2 ** 32, isq, None, _binding.ALGO_TILE, False)
    assert np.array_equal(b1, b2)
    e1 = set(zip(q1.tolist(), t1.tolist(), d1.tolist()))
    e2 = set(zip(q2.tolist(), t2.tolist(), d2.tolist()))
    assert e1 == e2
    got = c.ed_pairs(t1[:2000], q1[:2000], None)       # d(q,t) == d(t,q), recomputed unbounded
    assert np.array_equal(got, d1[:2000])
    assert np.all(b1[q1] == d1)
    c.close()


@pytest.mark.parametrize("cluster", ["1", "0"])
def test_similarity_order_of_the_targets_never_changes_the_graph(monkeypatch, cluster):
    """From 512 queries on (and read lengths within the MAIN threshold) the targets of the symmetric MAIN pass are
    laid out by similarity cluster instead of by length, and unordered pairs are split between rows by layout rank:
    a scheduling decision only."""
    monkeypatch.setenv("ISOCON_NN_CLUSTER", cluster)
    c = _binding.NNContext(0)
    for name, scale in (("c2", 0.06), ("c3", 0.012)):
        S = workloads.CONFIGS[name](scale=scale)
        L = _sorted_list_1set(S)
        seqs = [s for s, _ in L]
        P = util.Params(nr_cores=4)
        want, _ = O.compute_nearest_neighbor_graph(S, set(), P)
        G = _graph_via_ctx(c, L, 1, 2 ** 32, np.ones(len(L), np.uint8), None, _binding.ALGO_TILE, True)
        util.assert_same_graph(G, want, "%s cluster %s" % (name, cluster))
        assert (c.stats()["clusters"] > 0) == (cluster == "1")
        # some reads converged (not queries, still targets): every row takes its whole window
        hc = set(seqs[5::7])
        want = O.get_nearest_neighbors(L, 0, 0, L, hc, 2 ** 32)
        isq = np.array([0 if s in hc else 1 for s in seqs], dtype=np.uint8)
        G = _graph_via_ctx(c, L, 1, 2 ** 32, isq, None, _binding.ALGO_TILE, True)
        util.assert_same_graph(G, want, "%s cluster %s, converged reads" % (name, cluster))
    # one-sided pass (2-set): the candidates are ordered by min-hash clusters instead
    X, C, want = _c5_small()                              # 6000 reads x 300 candidates of 30 families
    L2 = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
    ist = np.array([1 if a in C else 0 for _, a in L2], dtype=np.uint8)
    G = _graph_via_ctx(c, L2, 2, 2 ** 32, 1 - ist, ist, _binding.ALGO_TILE, False)
    util.assert_same_graph(G, want, "c5 cluster %s" % cluster)
    assert (c.stats()["clusters"] > 0) == (cluster == "1")
    if cluster == "1":
        assert c.stats()["clusters"] <= 60                # 30 families (a few may split)
    c.close()


@pytest.mark.parametrize("two_level,surv_cap,seed_sample,qgram", [("1", "0", "1", "1"), ("0", "0", "1", "1"), ("1", "100", "1", "1"),
                                                                  ("1", "0", "0", "2"), ("1", "0", "1", "0")])
def test_two_level_one_sided_pass(monkeypatch, two_level, surv_cap, seed_sample, qgram):
    """One-sided passes over clustered targets go through the cluster representatives first (triangle inequality:
    d(q, rep) > k + radius dismisses the whole cluster), then meet the members of the surviving clusters.  Exact:
    the graphs equal the oracle's with it, without it, and through the fall-back after a survivor-buffer overflow."""
    monkeypatch.setenv("ISOCON_NN_TWO_LEVEL", two_level)
    monkeypatch.setenv("ISOCON_NN_SURV_CAP", surv_cap)
    monkeypatch.setenv("ISOCON_NN_SEED_SAMPLE", seed_sample)
    monkeypatch.setenv("ISOCON_NN_QGRAM", qgram)          # level 1: 1 = filter on the first pass, alignment later; 2 / 0 = always / never
    c = _binding.NNContext(0)
    X, C, want = _c5_small()
    P = util.Params(nr_cores=4)
    L2 = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
    ist = np.array([1 if a in C else 0 for _, a in L2], dtype=np.uint8)
    G = _graph_via_ctx(c, L2, 2, 2 ** 32, 1 - ist, ist, _binding.ALGO_TILE, False)
    util.assert_same_graph(G, want, "2-set two_level %s" % two_level)
    assert 0 < c.stats()["clusters"] <= 60
    # forced small ladder caps: several two-level passes
    monkeypatch.setenv("ISOCON_NN_LADDER_FIRST", "17")
    c2 = _binding.NNContext(0)
    G = _graph_via_ctx(c2, L2, 2, 2 ** 32, 1 - ist, ist, _binding.ALGO_TILE, False)
    util.assert_same_graph(G, want, "2-set two_level %s, ladder from 17" % two_level)
    assert c2.stats()["main_passes"] >= 3
    c2.close()
    monkeypatch.delenv("ISOCON_NN_LADDER_FIRST")
    # reads much nearer to their candidates (0.6 % errors: distances ~15) and much farther (6 %: ~150): the sample of
    # hinted rows picks another first cap for the rest (32; none) -- same graphs
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    cands = [np.array([code[ch] for ch in s], dtype=np.uint8) for s in C.values()]
    for err in (0.006, 0.06) if (two_level, surv_cap) == ("1", "0") else ():
        rng = np.random.default_rng(int(err * 1000))
        Xe = workloads._reads_from(rng, cands, rng.integers(0, len(cands), size=3000), err * 0.4, err * 0.4, err * 0.2)
        Le = sorted([(s, a) for a, s in Xe.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
        iste = np.array([1 if a in C else 0 for _, a in Le], dtype=np.uint8)
        want_e = O.compute_2set_nearest_neighbor_graph(Xe, C, P)
        G = _graph_via_ctx(c, Le, 2, 2 ** 32, 1 - iste, iste, _binding.ALGO_TILE, False)
        util.assert_same_graph(G, want_e, "2-set, reads with %g errors" % err)
    # one-sided 1-set over clean, clustered sequences: queries are targets too, some are their cluster's representative
    _, C6 = workloads.config5(scale=0.12)                 # 600 candidates of 60 families
    S = {"s%d" % i: s for i, s in enumerate(C6.values())}
    L = _sorted_list_1set(S)
    want, _ = O.compute_nearest_neighbor_graph(S, set(), P)
    G = _graph_via_ctx(c, L, 1, 2 ** 32, np.ones(len(L), np.uint8), None, _binding.ALGO_TILE, False)
    util.assert_same_graph(G, want, "one-sided 1-set two_level %s" % two_level)
    c.close()


# ----------------------------------------------------------------------------- residency across rounds (§8f-4)

def test_correction_rounds_upload_only_what_changed():
    """Three consecutive rounds through the reference-facing calls (fixtures of the unmodified reference driver,
    oracle/make_golden_r2.py): round k+1 is round k with ~5 % of the reads changed and some newly converged.  The
    graphs must be the reference's, and the device must receive exactly the sequences it has not seen before
    (isocon_nn_store_info counters): nothing that was resident is uploaded again."""
    ctx = _binding.get_context()
    ctx.store_reset()
    seen = set()
    rounds = util.correction_rounds()
    uploads = []
    for k, r in enumerate(rounds):
        Sp, hc = workloads.round1_call(r["S"])
        for kind in ("1set", "2set", "pairs"):
            before = ctx.store_info()
            if kind == "1set":
                lst = list(dict(zip(Sp.values(), Sp.keys())).keys())
                G, _ = nn.compute_nearest_neighbor_graph(Sp, hc, util.Params())
                util.assert_same_graph(G, r["graph_1set"], "round %d 1-set" % k)
            elif kind == "2set":
                lst = list(r["S"].values()) + list(r["C"].values())
                G = nn.compute_2set_nearest_neighbor_graph(r["S"], r["C"], util.Params())
                util.assert_same_graph(G, r["graph_2set"], "round %d 2-set" % k)
            else:
                # the pair list IsoCon asks for right after the graph (isocon_get_candidates.py:38): resident reads
                from isocon_b200 import edlib_alignment_module as em
                acc_of = {a: s for a, s in Sp.items()}
                matches = {acc_of[a]: [acc_of[b] for b in nb] for a, nb in list(G1.items())[:40] if nb}
                lst = []
                got = em.edlib_align_sequences(matches)
                for s1, nb in matches.items():
                    for s2 in nb:
                        assert got[s1][s2] == O.ed_plain(s1, s2)
            if kind == "1set":
                G1 = G
            after = ctx.store_info()
            expected = sum(1 for s in lst if s not in seen)
            seen.update(lst)
            assert after["uploaded_reads"] - before["uploaded_reads"] == expected, (k, kind)
            assert after["resets"] == before["resets"]
            uploads.append((k, kind, expected, len(lst)))
    # round 0 uploads every read once; later rounds only the corrected ones (a few per cent)
    assert uploads[0][2] == uploads[0][3]
    for k, kind, up, n in uploads[3:]:
        assert up <= max(1, n // 10), uploads


def test_unchanged_list_uploads_nothing_and_store_resets_when_mostly_dead():
    ctx = _binding.get_context()
    ctx.store_reset()
    S = workloads.config2(scale=0.01)
    P = util.Params()
    want, _ = O.compute_nearest_neighbor_graph(S, set(), P)
    a0 = ctx.store_info()                                              # counters are cumulative per context
    G, _ = nn.compute_nearest_neighbor_graph(S, set(), P)
    a = ctx.store_info()
    G2, _ = nn.compute_nearest_neighbor_graph(dict(S), set(), P)       # same content, new dict
    b = ctx.store_info()
    util.assert_same_graph(G, want); util.assert_same_graph(G2, want)
    assert a["uploaded_reads"] - a0["uploaded_reads"] == len(S) == a["slots"]
    assert b["uploaded_reads"] == a["uploaded_reads"] and b["lists"] == a["lists"]
    # a stream of unrelated small inputs fills the store with dead sequences: it starts over instead of growing
    for i in range(60):
        T = workloads.config2(scale=0.01, seed=100 + i)
        nn.compute_nearest_neighbor_graph(T, set(), P)
    c = ctx.store_info()
    assert c["resets"] > b["resets"] and c["slots"] <= ctx.STORE_SLACK * len(T) + 4096 + len(T)
    G3, _ = nn.compute_nearest_neighbor_graph(S, set(), P)
    util.assert_same_graph(G3, want)


def _tie_heavy(n, length, seed=5):
    rng = np.random.default_rng(seed)
    base = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=length)].tobytes().decode()
    S = {}
    for i in range(n):
        alt = "ACGT"[("ACGT".index(base[i]) + 1) % 4]
        S["t%d" % i] = base[:i] + alt + base[i + 1:]
    return S


def test_edge_buffer_regrows_on_tie_heavy_input():
    """Late correction rounds: every read at distance 2 from every other, no centre -> n (n - 1) edges, more than the
    default candidate-edge buffer max(2^20, 64 n) holds.  The binding reserves more and rebuilds."""
    ctx = _binding.get_context()
    ctx.reserve_edges(0)
    n = 1500
    S = _tie_heavy(n, 1600)
    G, _ = nn.compute_nearest_neighbor_graph(S, set(), util.Params())
    assert ctx.stats()["edges_raw"] >= n * (n - 1) > (1 << 20)
    want, _ = O.compute_nearest_neighbor_graph(S, set(), util.Params(nr_cores=4))
    util.assert_same_graph(G, want, "tie-heavy")
    ctx.reserve_edges(0)
    # a tiny reservation forces several regrowth rounds on an ordinary input
    ctx.reserve_edges(-8)
    Sp, hc = workloads.round1_call(util.load_reads(200))
    G, _ = nn.compute_nearest_neighbor_graph(Sp, hc, util.Params())
    util.assert_same_graph(G, util.c1_expected()["200"]["cases"]["1set_round1"]["graph"], "reserve 8")
    X, C = util.two_set_split(util.load_reads(200))
    ctx.reserve_edges(-8)
    G = nn.compute_2set_nearest_neighbor_graph(X, C, util.Params(neighbor_search_depth=3))     # SCAN algorithm
    util.assert_same_graph(G, util.c1_expected()["200"]["cases"]["2set_every17_depth3"]["graph"], "reserve 8 scan")
    ctx.reserve_edges(0)


# ----------------------------------------------------------------------------- explicit pair lists (§8f-1)

def test_pair_module_matches_the_reference_fixture():
    import json
    from isocon_b200 import edlib_alignment_module as pm
    with open(os.path.join(util.GOLD, "pairs_n200.json")) as fh:
        fx = json.load(fh)
    S = util.load_reads(200)
    Sp, _ = workloads.round1_call(S)
    matches = {Sp[q]: [Sp[t] for t, _ in rows] + [Sp[rows[0][0]]] for q, rows in fx["edlib_align_sequences"]}
    got = pm.edlib_align_sequences(matches, nr_cores=16)
    want = {Sp[q]: {Sp[t]: ed for t, ed in rows} for q, rows in fx["edlib_align_sequences"]}
    assert got == want and list(got) == list(want)
    assert all(list(got[k]) == list(want[k]) for k in want)
    X, C = util.two_set_split(S)
    by_cand = {c: {r: (C[c], X[r]) for r, _ in rows} for c, rows in fx["edlib_align_sequences_keeping_accession"]}
    got = pm.edlib_align_sequences_keeping_accession(by_cand, nr_cores=4)
    want = {c: {r: (C[c], X[r], ed) for r, ed in rows} for c, rows in fx["edlib_align_sequences_keeping_accession"]}
    assert got == want and list(got) == list(want)
    assert pm.edlib_align_sequences({}) == {} and pm.edlib_align_sequences_keeping_accession({}) == {}
    assert pm.edlib_alignment("ACGT", "AGGT", 0, 0) == ("ACGT", "AGGT", 1)
    assert pm.edlib_alignment("ACGT", "AGGTT", 0, 0, x_acc="a", y_acc="b") == ("a", "b", ("ACGT", "AGGTT", 2))
    assert pm.edlib_align_sequences({"ACGT": ["ACNT", "acgt"]}) == {"ACGT": {"ACNT": 1, "acgt": 4}}   # raw symbols, like edlib
    with pytest.raises(NotImplementedError):
        pm.edlib_traceback("ACGT", "ACGT")


# ----------------------------------------------------------------------------- larger instances: properties

def _random_pair_lower_bound(c, n, best, rng, pairs=4000):
    """No sampled pair may be closer than the best of either end (nothing closer was missed)."""
    a = rng.integers(0, n, size=pairs).astype(np.int32)
    b = rng.integers(0, n, size=pairs).astype(np.int32)
    keep = a != b
    a, b = a[keep], b[keep]
    d = c.ed_pairs(a, b, None)
    assert np.all(d >= 1)
    assert np.all(d >= best[a]) and np.all(d >= best[b])


@pytest.mark.parametrize("name,scale", [("c3", 0.1), ("c4", 0.05)])
def test_size_independent_properties_1set(name, scale):
    # c3 / c4 at a size the oracle cannot finish in seconds: (1) the symmetric pair pass with class bins,
    # the one-sided pass and the pass without bins agree edge for edge; (2) every reported distance is
    # re-derived unbounded, in the other argument order; (3) best[q] is the distance of q's edges and a
    # lower bound of sampled random pairs; (4) a second build of the same graph is identical (idempotence,
    # reads resident)
    S = workloads.CONFIGS[name](scale=scale)
    L = _sorted_list_1set(S)
    n = len(L)
    c = _binding.NNContext(0)
    c.set_reads([s for s, _ in L])
    isq = np.ones(n, np.uint8)
    b1, q1, t1, d1 = c.graph(1, 2 ** 32, isq, None, _binding.ALGO_TILE, True)
    assert c.stats()["pilot_rows"] > 0
    e1 = set(zip(q1.tolist(), t1.tolist(), d1.tolist()))
    b2, q2, t2, d2 = c.graph(1, 2 ** 32, isq, None, _binding.ALGO_TILE, False)
    assert np.array_equal(b1, b2) and e1 == set(zip(q2.tolist(), t2.tolist(), d2.tolist()))
    b3, q3, t3, d3 = c.graph(1, 2 ** 32, isq, None, _binding.ALGO_TILE, True)
    assert np.array_equal(b1, b3) and e1 == set(zip(q3.tolist(), t3.tolist(), d3.tolist()))
    qs, ts, ds = nn._order_edges(q1, t1, d1)
    sel = np.random.default_rng(3).choice(qs.size, size=min(3000, qs.size), replace=False)
    assert np.array_equal(c.ed_pairs(ts[sel], qs[sel], None), ds[sel])
    assert np.all(b1[qs] == ds)
    has_edge = np.zeros(n, bool); has_edge[qs] = True
    lens = np.array([len(s) for s, _ in L])
    assert np.all(b1[~has_edge] == lens[~has_edge])          # no neighbour within len(q): best stays len(q)
    _random_pair_lower_bound(c, n, b1, np.random.default_rng(4))
    c.close()


def test_size_independent_properties_2set():
    X, C = workloads.config5(scale=0.2)
    merged = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
    n = len(merged)
    ist = np.array([1 if a in C else 0 for _, a in merged], dtype=np.uint8)
    c = _binding.NNContext(0)
    c.set_reads([s for s, _ in merged])
    b1, q1, t1, d1 = c.graph(2, 2 ** 32, 1 - ist, ist, _binding.ALGO_TILE, False)
    b2, q2, t2, d2 = c.graph(2, 2 ** 32, 1 - ist, ist, _binding.ALGO_SCAN, False)     # the scan emulation
    assert np.array_equal(b1[ist == 0], b2[ist == 0])
    assert set(zip(q1.tolist(), t1.tolist(), d1.tolist())) == set(zip(q2.tolist(), t2.tolist(), d2.tolist()))
    assert np.all(ist[t1] == 1) and np.all(ist[q1] == 0)
    sel = np.random.default_rng(5).choice(q1.size, size=min(3000, q1.size), replace=False)
    assert np.array_equal(c.ed_pairs(t1[sel], q1[sel], None), d1[sel])
    rng = np.random.default_rng(6)
    reads = np.flatnonzero(ist == 0); cands = np.flatnonzero(ist == 1)
    a = rng.choice(reads, size=3000).astype(np.int32); b = rng.choice(cands, size=3000).astype(np.int32)
    assert np.all(c.ed_pairs(a, b, None) >= b1[a])
    c.close()


@pytest.mark.gpu
def test_two_gpus_sharded_graph_equals_the_graph_built_alone():
    """torchrun with one rank per GPU (NCCL, NVLink peer memory, box-wide tile queue, the MAIN ladder across ranks):
    tools/check_multi_gpu.py builds c2/c3/c4/c5 slices sharded, then alone on every rank, and compares."""
    import subprocess
    import sys
    from isocon_b200 import _binding
    if _binding.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(util.ROOT, "tools", "check_multi_gpu.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode()
    assert r.returncode == 0 and "DIFFERENT" not in out and out.count("-> same") >= 10, out[-3000:]
