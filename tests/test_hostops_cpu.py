"""CPU suite, part 5: the host-side helper library (isocon_b200/csrc/hostops.cpp, ctypes.PyDLL) against plain
Python statements of the same loops, and the vectorised list building of the reference-facing functions against
the reference's own statements (nearest_neighbor_graph.py:243-246, :202-208)."""
import ctypes

import numpy as np
import pytest

import util
from isocon_b200 import _hostops as H
from isocon_b200 import nearest_neighbor_graph as nn


def test_library_exports():
    L = H.load_library()
    for name in H.EXPORTS:
        assert hasattr(L, name), name


def test_lengths_lookup_register_gather():
    rng = np.random.default_rng(1)
    seqs = ["".join("ACGT"[c] for c in rng.integers(0, 4, size=int(rng.integers(0, 70)))) for _ in range(300)]
    seqs += seqs[:20]                                        # repeated strings
    assert H.lengths(seqs).tolist() == [len(s) for s in seqs]
    assert H.lengths([]).size == 0
    store = {}
    slots, missing = H.lookup(store, seqs)
    assert missing == len(seqs) and (slots == -1).all()
    sel = np.arange(0, len(seqs), 3, dtype=np.int32)
    H.register(store, seqs, sel, 1000)
    slots, missing = H.lookup(store, seqs)
    for i, s in enumerate(seqs):
        assert slots[i] == store.get(s, -1)
    assert missing == sum(1 for s in seqs if s not in store)
    buf = np.zeros(sum(len(seqs[i]) for i in sel) + 8, np.uint8)
    total, off = H.gather(seqs, sel, buf.ctypes.data, buf.size)
    assert bytes(buf[:total]).decode() == "".join(seqs[i] for i in sel)
    assert off.tolist() == np.concatenate([[0], np.cumsum([len(seqs[i]) for i in sel])]).tolist()
    with pytest.raises(BufferError):
        H.gather(seqs, sel, buf.ctypes.data, 3)
    with pytest.raises(ValueError):
        H.gather(["ACé"], [0], buf.ctypes.data, buf.size)
    with pytest.raises(TypeError):
        H.lengths(["AC", 5])


def _python_build(accs, lo, hi, skip, eq, et, ed):
    out = {}
    for i in range(lo, hi):
        if skip is None or not skip[i]:
            out[accs[i]] = {}
    q, t, d = nn._order_edges(np.asarray(eq, np.int32), np.asarray(et, np.int32), np.asarray(ed, np.int32))
    for a, b, c in zip(q.tolist(), t.tolist(), d.tolist()):
        out[accs[a]][accs[b]] = c
    return out


def test_build_graph_is_the_scan_order():
    rng = np.random.default_rng(3)
    n = 3000
    accs = ["r%d" % i for i in range(n)]
    skip = (rng.random(n) < 0.2).astype(np.uint8)
    eq = rng.integers(0, n, size=20000).astype(np.int32)
    et = rng.integers(0, n, size=20000).astype(np.int32)
    keep = (eq != et) & (skip[eq] == 0)
    eq, et = eq[keep], et[keep]
    ed = ((eq.astype(np.int64) * 31 + et) % 300).astype(np.int32)          # a function of the pair
    eq = np.concatenate([eq, eq[:500]]); et = np.concatenate([et, et[:500]]); ed = np.concatenate([ed, ed[:500]])
    for sk in (skip, None):
        if sk is None:
            got = H.build_graph(accs, 0, n, None, eq, et, ed)
            want = _python_build(accs, 0, n, None, eq, et, ed)
        else:
            got = H.build_graph(accs, 0, n, sk, eq, et, ed)
            want = _python_build(accs, 0, n, sk, eq, et, ed)
        util.assert_same_graph(got, want)
        assert all(type(v) is int for nb in got.values() for v in nb.values())
    # a sub-range of keys (a Pool worker's chunk, nearest_neighbor_graph.py:33-54)
    lo, hi = 100, 900
    m = (eq >= lo) & (eq < hi)
    util.assert_same_graph(H.build_graph(accs, lo, hi, None, eq[m], et[m], ed[m]),
                           _python_build(accs, lo, hi, None, eq[m], et[m], ed[m]))
    with pytest.raises(ValueError):
        H.build_graph(accs, lo, hi, None, eq, et, ed)                  # an edge of a query outside the range
    # two entries under one accession: one key, one dict (the later assignment wins it), edges of both land there
    dup = list(accs)
    dup[7] = dup[5]
    m = eq < 50
    util.assert_same_graph(H.build_graph(dup, 0, n, None, eq[m], et[m], ed[m]), _python_build(dup, 0, n, None, eq[m], et[m], ed[m]))
    with pytest.raises(ValueError):                                     # an edge of an entry that is no key
        H.build_graph(accs, 0, n, np.ones(n, np.uint8), eq[:1], et[:1], ed[:1])
    z = np.zeros(0, np.int32)
    assert H.build_graph(accs, 0, 3, None, z, z, z) == {"r0": {}, "r1": {}, "r2": {}}
    assert H.build_graph([], 0, 0, None, z, z, z) == {}


def test_sorted_list_building_matches_the_reference_statements():
    rng = np.random.default_rng(5)
    for _ in range(20):
        n = int(rng.integers(0, 60))
        seqs = ["".join("ACGT"[c] for c in rng.integers(0, 4, size=int(rng.integers(1, 12)))) for _ in range(n)]
        S = {"a%d" % i: s for i, s in enumerate(seqs)}                 # duplicates: the LAST accession wins (:243)
        ref_map = {seq: acc for (acc, seq) in S.items()}
        ref_list = sorted(list(ref_map.items()), key=lambda x: len(x[0]))      # :243-246
        by_seq = dict(zip(S.values(), S.keys()))
        got_s, got_a, lens = nn._sorted_by_length(list(by_seq.keys()), list(by_seq.values()))
        assert list(zip(got_s, got_a)) == ref_list and lens.tolist() == [len(s) for s in got_s]
        X = dict(list(S.items())[: n // 2]); C = {"c" + a: s for a, s in list(S.items())[n // 2:]}
        ref2 = sorted([(seq, acc) for (acc, seq) in X.items()] + [(seq, acc) for (acc, seq) in C.items()],
                      key=lambda x: len(x[0]))                                 # :202-208
        s2 = list(X.values()); s2.extend(C.values())
        a2 = list(X.keys()); a2.extend(C.keys())
        got_s, got_a, _ = nn._sorted_by_length(s2, a2)
        assert list(zip(got_s, got_a)) == ref2


def test_permute_and_contains():
    rng = np.random.default_rng(9)
    items = ["s%d" % i for i in range(500)]
    order = rng.permutation(500)
    assert H.permute(items, order) == [items[i] for i in order]
    assert H.permute(tuple(items), order[:7]) == [items[i] for i in order[:7]]
    assert H.permute([], np.zeros(0, np.int64)) == []
    with pytest.raises(IndexError):
        H.permute(items, [0, 500])
    keys = ["s%d" % i for i in rng.integers(0, 800, size=300)]
    for container in (set(items[::3]), {k: 1 for k in items[::5]}, frozenset(items[:10]), items[:50]):
        assert H.contains(container, keys).tolist() == [1 if k in container else 0 for k in keys]
    assert H.contains(set(), []).size == 0
