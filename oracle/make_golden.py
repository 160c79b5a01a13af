"""oracle/make_golden.py -- generates tests/golden/ (authoring container only).

Runs the UNMODIFIED reference driver (``/root/reference/modules/nearest_neighbor_graph.py``,
imported through ``oracle/reference_driver.py`` with the edlib stand-in) on

  * the known-answer corner cases K1-K11 of SURVEY.md Appendix B,
  * the four shipped FASTA fixtures ``test/data/simulated_pacbio_reads_n_{200,500,1000,2000}.fa``
    (round-1 1-set call as graphs.py:37-58 makes it; a 2-set call; finite-depth,
    multi-core and has_converged variants),
  * seeded random small read sets,

and writes inputs + outputs as fixtures.  Inputs of the FASTA fixtures are stored 2-bit
packed (``*.npz``) because ``/root/reference`` does not travel to the GPU box.

Every reference output is also compared here with the C++ oracle (``oracle/oracle.py``)
and, on small inputs, with the plain-DP arithmetic and the closed form of Appendix A.3;
the script aborts on any mismatch.  Usage:  python oracle/make_golden.py
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O                # noqa: E402
from oracle import reference_driver as rd     # noqa: E402
from isocon_b200 import workloads             # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CODE = {c: i for i, c in enumerate("ACGT")}


def digest(G):
    """SURVEY.md Appendix B digest."""
    return hashlib.sha256(json.dumps([[a, list(v.items())] for a, v in G.items()]).encode()).hexdigest()[:16]


def as_lists(G):
    return [[a, [[b, d] for b, d in v.items()]] for a, v in G.items()]


def same(G1, G2):
    return as_lists(G1) == as_lists(G2)


def pack_reads(S, path):
    accs = list(S.keys())
    lens = np.array([len(S[a]) for a in accs], dtype=np.int32)
    cat = np.frombuffer("".join(S[a] for a in accs).encode(), dtype=np.uint8)
    lut = np.full(256, 255, np.uint8)
    for c, i in CODE.items():
        lut[ord(c)] = i
    codes = lut[cat]
    assert codes.max() < 4, "fixture is not pure ACGT"
    pad = (-codes.size) % 4
    codes = np.concatenate([codes, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    packed = (codes[:, 0] | (codes[:, 1] << 2) | (codes[:, 2] << 4) | (codes[:, 3] << 6)).astype(np.uint8)
    np.savez_compressed(path, acc=np.array(accs), lens=lens, packed=packed)


def ref_1set(ref, S, has_converged, **kw):
    with rd.quiet():
        G, iso = ref.compute_nearest_neighbor_graph(S, has_converged, rd.Params(**kw))
    assert iso == set()
    return G


def ref_2set(ref, X, C, **kw):
    with rd.quiet():
        return ref.compute_2set_nearest_neighbor_graph(X, C, rd.Params(**kw))


def check_1set(ref, S, hc, small=False, **kw):
    G = ref_1set(ref, S, hc, **kw)
    G2, _ = O.compute_nearest_neighbor_graph(S, hc, rd.Params(**kw))
    assert same(G, G2), "C++ oracle != reference driver (1-set) %r" % (kw,)
    if small:
        lst = sorted({s: a for a, s in S.items()}.items(), key=lambda x: len(x[0]))
        G3 = O.closed_form_1set(lst, hc, kw.get("neighbor_search_depth", 2 ** 32))
        assert same(G, G3), "closed form != reference driver (1-set)"
    return G


def check_2set(ref, X, C, small=False, **kw):
    G = ref_2set(ref, X, C, **kw)
    G2 = O.compute_2set_nearest_neighbor_graph(X, C, rd.Params(**kw))
    assert same(G, G2), "C++ oracle != reference driver (2-set) %r" % (kw,)
    if small and kw.get("neighbor_search_depth", 2 ** 32) >= 2 ** 32:
        lst = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda x: len(x[0]))
        G3 = O.closed_form_2set(lst, set(C))
        assert same(G, G3), "closed form != reference driver (2-set)"
    return G


def two_set_split(S, every=17):
    """Deterministic 2-set instance from a read set: every 17th read doubles as candidate cand_i."""
    accs = list(S.keys())
    C = {"cand_%d" % i: S[a] for i, a in enumerate(accs[::every])}
    return dict(S), C


def known_answers(ref):
    K = []
    def add(name, kind, expect=None, **inp):
        kw = {k: inp.pop(k) for k in ("neighbor_search_depth", "nr_cores") if k in inp}
        if kind == "1set":
            hc = set(inp.get("has_converged", []))
            G = check_1set(ref, inp["S"], hc, small=True, **kw)
        else:
            G = check_2set(ref, inp["X"], inp["C"], small=True, **kw)
        if expect is not None:
            assert G == expect, (name, G, expect)
        K.append(dict(name=name, kind=kind, params=kw, graph=as_lists(G), **inp))
    add("K1", "1set", {'a': {'b': 4}, 'b': {'a': 4}}, S={"a": "AAAA", "b": "CCCC"})
    add("K2", "1set", {'a': {}, 'b': {'a': 6}}, S={"a": "AAAA", "b": "CCCCCC"})
    K3 = dict(a="ACGTACGT", b="ACGTACGA", c="ACGTACGC", d="ACGTACG", e="ACGTACGTT")
    add("K3", "1set", {'d': {'a': 1, 'b': 1, 'c': 1}, 'a': {'d': 1, 'b': 1, 'c': 1, 'e': 1},
                       'b': {'a': 1, 'c': 1, 'd': 1}, 'c': {'b': 1, 'a': 1, 'd': 1}, 'e': {'a': 1}}, S=K3)
    add("K4", "1set", {'a': {}, 'b': {'a': 1}, 'c': {'b': 3}},
        S=dict(a="ACGTACGT", b="ACGTACGA", c="TTTTACGA"), has_converged=["ACGTACGT"])
    K5 = dict(X=dict(r1="ACGTACGT", r2="ACGTACGA"), C=dict(c1="ACGTACGT", c2="ACGTACGG", c3="ACGTACG"))
    add("K5", "2set", {'r1': {'c1': 0}, 'r2': {'c1': 1, 'c3': 1, 'c2': 1}}, **K5)
    add("K6", "2set", {'r1': {}}, X=dict(r1="AAAA"), C=dict(c1="CCCCCCCCCCCC"))
    add("K7", "1set", {'d': {'a': 1}, 'a': {'d': 1, 'b': 1}, 'b': {'a': 1, 'c': 1}, 'c': {'b': 1}, 'e': {'c': 2}},
        S=K3, neighbor_search_depth=1)
    add("K8", "2set", {'r1': {'c3': 1}, 'r2': {'c1': 1}}, neighbor_search_depth=1, **K5)
    add("K10", "1set", {'a': {}}, S={"a": "ACGT"})
    add("K12_multicore", "1set", None, S=K3, nr_cores=3)
    add("K13_2set_multicore", "2set", None, nr_cores=2, **K5)
    return K


def random_cases(ref, n_cases=40, seed=1234):
    rng = np.random.default_rng(seed)
    out = []
    for c in range(n_cases):
        L = int(rng.integers(5, 120))
        n_tpl = int(rng.integers(1, 4))
        tpls = [rng.integers(0, 4, size=max(1, L + int(rng.integers(-4, 5))), dtype=np.uint8) for _ in range(n_tpl)]
        n = int(rng.integers(2, 40))
        err = float(rng.choice([0.01, 0.05, 0.15, 0.3]))
        seqs = []
        for _ in range(n):
            t = tpls[int(rng.integers(0, n_tpl))]
            r = workloads._mutate(rng, t, err / 3, err / 3, err / 3)
            if r.size == 0:
                r = np.array([0], dtype=np.uint8)
            seqs.append(workloads._to_str(r))
        S = {"s%d" % i: s for i, s in enumerate(seqs)}            # may contain duplicates
        Sp, hc = workloads.round1_call(S)
        depth = int(rng.choice([2 ** 32, 2 ** 32, 1, 2, 3, 7]))
        cores = int(rng.choice([1, 1, 2, 3]))
        kw = dict(neighbor_search_depth=depth, nr_cores=cores)
        if len(Sp) >= 1:
            G = check_1set(ref, Sp, hc, small=(cores == 1), **kw)
            out.append(dict(name="rand1_%d" % c, kind="1set", params=kw, S=Sp, has_converged=sorted(hc),
                            graph=as_lists(G)))
        ncand = int(rng.integers(1, max(2, n // 3)))
        X = dict(S)
        C = {"c%d" % i: seqs[int(rng.integers(0, n))] if rng.random() < 0.5 else
             workloads._to_str(tpls[int(rng.integers(0, n_tpl))]) for i in range(ncand)}
        G = check_2set(ref, X, C, small=(cores == 1), **kw)
        out.append(dict(name="rand2_%d" % c, kind="2set", params=kw, X=X, C=C, graph=as_lists(G)))
    return out


def fasta_fixtures(ref):
    meta = {}
    expected_digest = {200: "b8d19ad1bc0ebd54", 500: "864f522702934c55",
                       1000: "806f1412d7f8f891", 2000: "f2f580c6efbfde0e"}
    for n in (200, 500, 1000, 2000):
        S = workloads.read_fasta(os.path.join(rd.REFERENCE_ROOT, "test", "data",
                                              "simulated_pacbio_reads_n_%d.fa" % n))
        pack_reads(S, os.path.join(GOLD, "c1_n%d.npz" % n))
        Sp, hc = workloads.round1_call(S)
        entry = {"reads": len(S), "unique": len(Sp), "has_converged": len(hc), "cases": {}}

        def put(name, G, stats=None):
            e = dict(digest=digest(G), edges=sum(len(v) for v in G.values()),
                     sum_ed=sum(d for v in G.values() for d in v.values()), graph=as_lists(G))
            if stats:
                e["work"] = stats
            entry["cases"][name] = e

        _, edlib = rd.load()
        edlib._reset()
        G = check_1set(ref, Sp, hc)
        work = edlib._counters()
        assert digest(G) == expected_digest[n], (n, digest(G))
        assert work == O.LAST_STATS, (work, O.LAST_STATS)        # C++ oracle counts the same work
        put("1set_round1", G, work)
        print("n_%d: 1-set digest %s  work %s" % (n, digest(G), work), flush=True)
        if n <= 1000:
            put("1set_round1_cores3", check_1set(ref, Sp, hc, nr_cores=3))
            for depth in (1, 5, 20):
                put("1set_depth%d" % depth, check_1set(ref, Sp, hc, neighbor_search_depth=depth))
            put("1set_noconv", check_1set(ref, Sp, set()))
        X, C = two_set_split(S)
        put("2set_every17", check_2set(ref, X, C))
        if n <= 1000:
            put("2set_every17_cores3", check_2set(ref, X, C, nr_cores=3))
            for depth in (1, 3, 10):
                put("2set_every17_depth%d" % depth, check_2set(ref, X, C, neighbor_search_depth=depth))
        if n == 200:
            # the arithmetic itself: all pairs of the unique reads, plain DP vs Myers vs banded DP
            seqs = list(Sp.values())
            a, b = np.triu_indices(len(seqs), 1)
            d_my = O.ed_pairs(seqs, a, b)
            rng = np.random.default_rng(7)
            pick = rng.choice(len(a), size=1500, replace=False)
            for p in pick:
                x, y = seqs[a[p]], seqs[b[p]]
                d = O.ed_plain(x, y)
                assert d == d_my[p]
                for k in (d - 1, d, d + 1, d + 40, max(len(x), len(y))):
                    if k < 0:
                        continue
                    want = d if d <= k else -1
                    assert O.ed_myers64(x, y, k) == want and O.ed_banded_dp(x, y, k) == want, (len(x), len(y), d, k)
            np.savez_compressed(os.path.join(GOLD, "c1_n200_allpairs.npz"), a=a.astype(np.int32),
                                b=b.astype(np.int32), ed=d_my.astype(np.int32), acc=np.array(list(Sp.keys())))
            print("n_200: %d pairwise distances stored, 1500 verified against plain DP" % len(a), flush=True)
        meta[str(n)] = entry
    return meta


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref, _ = rd.load()
    K = known_answers(ref)
    R = random_cases(ref)
    with open(os.path.join(GOLD, "known_answers.json"), "w") as fh:
        json.dump(dict(generator="oracle/make_golden.py", source="unmodified reference driver + edlib stand-in",
                       cases=K + R), fh, indent=0)
    print("known answers: %d cases" % (len(K) + len(R)), flush=True)
    F = fasta_fixtures(ref)
    with gzip.open(os.path.join(GOLD, "c1_expected.json.gz"), "wt", compresslevel=9) as fh:
        json.dump(dict(generator="oracle/make_golden.py", source="unmodified reference driver + edlib stand-in",
                       fixtures=F), fh)
    print("done")


if __name__ == "__main__":
    main()
