"""oracle/make_golden_r2.py -- TEST INFRASTRUCTURE, authoring container only (second batch of fixtures).

Like ``oracle/make_golden.py`` this runs the UNMODIFIED reference driver
(``/root/reference/modules/nearest_neighbor_graph.py`` through ``oracle/reference_driver.py`` with the edlib
stand-in) and stores inputs + outputs, here for the cases the first batch left out:

  * ``neighbor_search_depth <= 0`` (the scan tests ``j >= depth`` / ``processed >= depth`` AFTER offset 1,
    nearest_neighbor_graph.py:190, :416),
  * alphabets beyond upper-case ACGT -- edlib compares raw characters, so ``A``, ``N`` and ``n`` are three symbols
    (SURVEY.md Appendix A.4, K9): N-containing reads, lower-case (soft-masked) reads, mixed case, RNA,
  * a correction-round SEQUENCE: three consecutive graph builds where round k+1 is round k with ~5 % of the reads
    changed and some reads newly converged (duplicates), the call pattern of isocon_get_candidates.py:141-214
    with graphs.py:37-58 in front -- the input of the residency / delta-upload test,
  * tie-heavy inputs (every read at the same distance from every other, no centre).

Every reference output is compared with the C++ oracle here; the script aborts on any mismatch.
Writes tests/golden/known_answers_r2.json and tests/golden/rounds_r2.json.gz.   Usage: python oracle/make_golden_r2.py
"""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_driver as rd     # noqa: E402
from oracle.make_golden import GOLD, as_lists, check_1set, check_2set  # noqa: E402
from isocon_b200 import workloads             # noqa: E402


def _rand_reads(rng, n, L, err, alphabet="ACGT", n_tpl=2):
    abc = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    tpls = [rng.integers(0, 4, size=max(2, L + int(rng.integers(-3, 4))), dtype=np.uint8) for _ in range(n_tpl)]
    out = []
    for _ in range(n):
        r = workloads._mutate(rng, tpls[int(rng.integers(0, n_tpl))], err / 3, err / 3, err / 3)
        if r.size == 0:
            r = np.array([0], dtype=np.uint8)
        out.append(abc[r].tobytes().decode())
    return out


def _sprinkle(rng, s, symbols, rate):
    """Replace a fraction of the positions of s by symbols outside its alphabet."""
    b = bytearray(s.encode())
    for i in range(len(b)):
        if rng.random() < rate:
            b[i] = ord(symbols[int(rng.integers(0, len(symbols)))])
    return b.decode()


def depth_cases(ref):
    K = []

    def add(name, kind, **inp):
        kw = {k: inp.pop(k) for k in ("neighbor_search_depth", "nr_cores") if k in inp}
        if kind == "1set":
            G = check_1set(ref, inp["S"], set(inp.get("has_converged", [])), **kw)
        else:
            G = check_2set(ref, inp["X"], inp["C"], **kw)
        K.append(dict(name=name, kind=kind, params=kw, graph=as_lists(G), **inp))

    K3 = dict(a="ACGTACGT", b="ACGTACGA", c="ACGTACGC", d="ACGTACG", e="ACGTACGTT")
    K5 = dict(X=dict(r1="ACGTACGT", r2="ACGTACGA"), C=dict(c1="ACGTACGT", c2="ACGTACGG", c3="ACGTACG"))
    for depth in (0, -1, -7):
        add("K3_depth%d" % depth, "1set", S=K3, neighbor_search_depth=depth)
        add("K5_depth%d" % depth, "2set", neighbor_search_depth=depth, **K5)
    # 2-set, depth 0: offset 1 holds no candidate for some reads (the scan still stops there)
    add("K14_2set_depth0_gap", "2set", neighbor_search_depth=0,
        X=dict(r1="ACGTACGT", r2="ACGTACGA", r3="ACGTAC"), C=dict(c1="ACGTACGTAA", c2="ACGT"))
    rng = np.random.default_rng(2024)
    for c in range(12):
        seqs = _rand_reads(rng, int(rng.integers(3, 30)), int(rng.integers(6, 90)), float(rng.choice([0.02, 0.1, 0.25])))
        S = {"s%d" % i: s for i, s in enumerate(seqs)}
        Sp, hc = workloads.round1_call(S)
        depth = int(rng.choice([0, -1, -100]))
        add("depth_rand1_%d" % c, "1set", S=Sp, has_converged=sorted(hc), neighbor_search_depth=depth)
        C = {"c%d" % i: seqs[int(rng.integers(0, len(seqs)))] for i in range(int(rng.integers(1, 6)))}
        add("depth_rand2_%d" % c, "2set", X=dict(S), C=C, neighbor_search_depth=depth)
    return K


def alphabet_cases(ref):
    K = []

    def add(name, kind, expect=None, **inp):
        kw = {k: inp.pop(k) for k in ("neighbor_search_depth", "nr_cores") if k in inp}
        if kind == "1set":
            G = check_1set(ref, inp["S"], set(inp.get("has_converged", [])), **kw)
        else:
            G = check_2set(ref, inp["X"], inp["C"], **kw)
        if expect is not None:
            assert G == expect, (name, G, expect)
        K.append(dict(name=name, kind=kind, params=kw, graph=as_lists(G), **inp))

    # every pair at distance 1 (exact-character comparison)
    add("K9", "1set", {'a': {'b': 1, 'c': 1}, 'b': {'a': 1, 'c': 1}, 'c': {'b': 1, 'a': 1}},
        S=dict(a="ACGTNCGT", b="ACGTnCGT", c="ACGTACGT"))
    add("K9_2set", "2set", None, X=dict(r1="ACGTNCGT", r2="ACGTACGT", r3="acgtacgt"),
        C=dict(c1="ACGTnCGT", c2="ACGTNCGT", c3="ACGTACGT", c4="ACGTACGt"))
    add("K15_all_foreign", "1set", None, S=dict(a="NNNN", b="NNNNN", c="nnnn", d="RYKM", e="NNRN"))
    add("K16_depth1_foreign", "1set", None, neighbor_search_depth=1,
        S=dict(a="ACGTNCGT", b="ACGTnCGT", c="ACGTACGT", d="ACGTCGT", e="ACNTNCGTA"))
    rng = np.random.default_rng(77)
    c = 0
    for alphabet, foreign, rate in (("acgt", "", 0.0), ("ACGU", "", 0.0), ("ACGT", "N", 0.02), ("ACGT", "Nn", 0.05),
                                    ("ACGT", "acgt", 0.3), ("ACGT", "RYKMSWN", 0.01), ("acgt", "ACGTN", 0.05),
                                    ("ACGT", "N", 0.5), ("TGCA", "*-", 0.03)):
        for rep in range(3):
            n = int(rng.integers(4, 36))
            seqs = _rand_reads(rng, n, int(rng.integers(8, 150)), float(rng.choice([0.03, 0.12])), alphabet)
            if foreign:
                # only some reads carry foreign symbols: the others stay on the 2-bit path
                seqs = [_sprinkle(rng, s, foreign, rate) if rng.random() < 0.6 else s for s in seqs]
            S = {"s%d" % i: s for i, s in enumerate(seqs)}
            Sp, hc = workloads.round1_call(S)
            depth = int(rng.choice([2 ** 32, 2 ** 32, 2 ** 32, 2, 5]))
            add("alpha1_%d" % c, "1set", S=Sp, has_converged=sorted(hc), neighbor_search_depth=depth)
            C = {"c%d" % i: seqs[int(rng.integers(0, n))] if rng.random() < 0.5 else
                 _sprinkle(rng, seqs[int(rng.integers(0, n))], foreign or alphabet, 0.05)
                 for i in range(int(rng.integers(1, max(2, n // 3))))}
            add("alpha2_%d" % c, "2set", X=dict(S), C=C, neighbor_search_depth=depth)
            c += 1
    return K


def tie_cases(ref):
    """Every read one substitution away from a centre that is NOT in the set: all pairs at distance 2."""
    K = []
    base = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(5).integers(0, 4, size=300)].tobytes().decode()
    S = {}
    for i in range(120):
        alt = "ACGT"[("ACGT".index(base[i]) + 1) % 4]
        S["t%d" % i] = base[:i] + alt + base[i + 1:]
    G = check_1set(ref, S, set())
    assert all(len(v) == len(S) - 1 for v in G.values())
    K.append(dict(name="ties_120", kind="1set", params={}, graph=as_lists(G), S=S))
    return K


def correction_rounds(ref):
    """Three rounds: S_k (all reads, duplicates allowed) -> the call graphs.py:37-58 makes + a 2-set call."""
    rng = np.random.default_rng(99)
    root = rng.integers(0, 4, size=420, dtype=np.uint8)
    copies = [workloads._diverge(rng, root, 0.01, 1) for _ in range(4)]
    picks = rng.integers(0, 4, size=260)
    S = workloads._reads_from(rng, copies, picks, 0.02, 0.012, 0.008)
    C = {"cand_%d" % i: workloads._to_str(c) for i, c in enumerate(copies)}
    rounds = []
    accs = list(S)
    for k in range(3):
        if k > 0:
            # "correction": ~5 % of the reads change (a few of them INTO the sequence of another read, so they
            # become converged duplicates, graphs.py:47-48); candidates: one is dropped per round
            S = dict(S)
            chosen = rng.choice(len(accs), size=max(1, len(accs) // 20), replace=False)
            for j, i in enumerate(chosen):
                a = accs[int(i)]
                if j % 4 == 0:
                    S[a] = S[accs[int(rng.integers(0, len(accs)))]]
                else:
                    r = np.frombuffer(S[a].encode(), dtype=np.uint8).copy()
                    pos = rng.choice(r.size, size=3, replace=False)
                    r[pos] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=3)]
                    S[a] = r.tobytes().decode()
            C = dict(list(C.items())[:-1])
        Sp, hc = workloads.round1_call(S)
        G1 = check_1set(ref, Sp, hc)
        G2 = check_2set(ref, S, C)
        rounds.append(dict(S=S, C=C, graph_1set=as_lists(G1), graph_2set=as_lists(G2),
                           unique=len(Sp), converged=len(hc)))
        print("round %d: %d reads, %d unique, %d converged, %d / %d edges" % (
            k, len(S), len(Sp), len(hc), sum(len(v) for v in G1.values()), sum(len(v) for v in G2.values())), flush=True)
    return rounds


def main():
    ref, _ = rd.load()
    K = depth_cases(ref) + alphabet_cases(ref) + tie_cases(ref)
    with open(os.path.join(GOLD, "known_answers_r2.json"), "w") as fh:
        json.dump(dict(generator="oracle/make_golden_r2.py", source="unmodified reference driver + edlib stand-in",
                       cases=K), fh, indent=0)
    print("known answers r2: %d cases" % len(K), flush=True)
    R = correction_rounds(ref)
    with gzip.open(os.path.join(GOLD, "rounds_r2.json.gz"), "wt", compresslevel=9) as fh:
        json.dump(dict(generator="oracle/make_golden_r2.py", source="unmodified reference driver + edlib stand-in",
                       rounds=R), fh)
    print("done")


if __name__ == "__main__":
    main()
