"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``oracle/liboracle.so`` (built by ``oracle/Makefile``): the CPU
restatement of IsoCon's nearest-neighbour-graph hot path.  Function names and argument
meaning mirror ``/root/reference/modules/nearest_neighbor_graph.py`` so the parity tests
read like calls into the reference:

* ``edlib_ed``                                 <- nearest_neighbor_graph.py:104-107
* ``get_nearest_neighbors``                    <- :110-198
* ``get_nearest_neighbors_2set``               <- :341-424
* ``get_exact_nearest_neighbor_graph``         <- :19-82   (``nr_cores`` chunking emulated)
* ``get_exact_nearest_neighbor_graph_2set``    <- :300-334
* ``compute_nearest_neighbor_graph``           <- :237-296
* ``compute_2set_nearest_neighbor_graph``      <- :201-234

Parity status: pinned against the unmodified reference driver by ``oracle/make_golden.py``
(fixtures in ``tests/golden/``); the reference itself ships no golden vectors for this path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle` (or __graft_entry__.build())")
        L = ctypes.CDLL(path)
        u8p = ctypes.c_char_p
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
        i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
        u8a = np.ctypeslib.ndpointer(np.uint8, flags="C")
        u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
        L.oracle_ed_plain.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int]
        L.oracle_ed_banded_dp.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int]
        L.oracle_ed_myers64.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int]
        L.oracle_ed_pairs.argtypes = [u8a, i64p, i32p, i32p, ctypes.c_void_p, ctypes.c_int64, i32p]
        L.oracle_ed_pairs.restype = None
        for f in (L.nn_oracle_1set, L.nn_oracle_2set):
            f.argtypes = [u8a, i64p, ctypes.c_int, u8a, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                          ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p, i32p, i32p, ctypes.c_int64, u64p]
            f.restype = ctypes.c_int64
        _LIB = L
    return _LIB


def _b(s):
    return s if isinstance(s, bytes) else s.encode("latin-1")


def ed_plain(x, y):
    x, y = _b(x), _b(y)
    return lib().oracle_ed_plain(x, len(x), y, len(y))


def ed_banded_dp(x, y, k):
    x, y = _b(x), _b(y)
    return lib().oracle_ed_banded_dp(x, len(x), y, len(y), k)


def ed_myers64(x, y, k=-1):
    x, y = _b(x), _b(y)
    return lib().oracle_ed_myers64(x, len(x), y, len(y), k)


def edlib_ed(x, y, mode="NW", task="distance", k=1):
    """nearest_neighbor_graph.py:104-107."""
    assert mode == "NW" and task == "distance"
    return ed_myers64(x, y, k)


def concat(seqs):
    """list of str -> (uint8 concatenation, int64 offsets[n+1])."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    cat = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8)
    if cat.size == 0:
        cat = np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(cat), off


def ed_pairs(seqs, a, b, k=None):
    cat, off = concat(seqs)
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    out = np.empty(len(a), dtype=np.int32)
    kk = None
    if k is not None:
        kk = np.ascontiguousarray(k, dtype=np.int32)
    lib().oracle_ed_pairs(cat, off, a, b, kk.ctypes.data if kk is not None else None, len(a), out)
    return out


def _run(fn, sorted_list, mask, depth, q_start, q_count, cores, threads, use_plain):
    n = len(sorted_list)
    cat, off = concat([s for s, _ in sorted_list])
    best = np.full(max(n, 1), -1, dtype=np.int32)
    stats = np.zeros(4, dtype=np.uint64)
    depth = int(min(depth, 2 ** 62))
    cap = max(4 * n, 1024)
    while True:
        eq = np.empty(cap, np.int32); et = np.empty(cap, np.int32); ed = np.empty(cap, np.int32)
        ne = fn(cat, off, n, mask, depth, q_start, q_count, cores, threads, int(use_plain),
                best, eq, et, ed, cap, stats)
        if ne <= cap:
            break
        cap = int(ne)
    return best, eq[:ne], et[:ne], ed[:ne], dict(calls=int(stats[0]), neg=int(stats[1]),
                                                  cells_full=int(stats[2]), cells_band=int(stats[3]))


LAST_STATS = {}


def get_nearest_neighbors(batch_of_queries, global_index_in_matrix, start_index, seq_to_acc_list_sorted,
                          has_converged, neighbor_search_depth, _cores=1, _threads=1, _use_plain=False):
    """nearest_neighbor_graph.py:110-198 (queries = list entries [start_index, start_index+len(batch)))."""
    L = seq_to_acc_list_sorted
    conv = np.fromiter((1 if s in has_converged else 0 for s, _ in L), dtype=np.uint8, count=len(L))
    if conv.size == 0:
        conv = np.zeros(1, np.uint8)
    best, eq, et, ed, st = _run(lib().nn_oracle_1set, L, conv, neighbor_search_depth, start_index,
                                len(batch_of_queries), _cores, _threads, _use_plain)
    LAST_STATS.clear(); LAST_STATS.update(st)
    out = {}
    for i in range(start_index, start_index + len(batch_of_queries)):
        out[L[i][1]] = {}
    for q, t, d in zip(eq.tolist(), et.tolist(), ed.tolist()):
        out[L[q][1]][L[t][1]] = d
    return out


def get_nearest_neighbors_2set(batch, start_index, seq_to_acc_list_sorted, target_accessions,
                               neighbor_search_depth, _cores=1, _threads=1, _use_plain=False):
    """nearest_neighbor_graph.py:341-424."""
    L = seq_to_acc_list_sorted
    tgt = np.fromiter((1 if a in target_accessions else 0 for _, a in L), dtype=np.uint8, count=len(L))
    if tgt.size == 0:
        tgt = np.zeros(1, np.uint8)
    best, eq, et, ed, st = _run(lib().nn_oracle_2set, L, tgt, neighbor_search_depth, start_index,
                                len(batch), _cores, _threads, _use_plain)
    LAST_STATS.clear(); LAST_STATS.update(st)
    out = {}
    for i in range(start_index, start_index + len(batch)):
        if not tgt[i]:
            out[L[i][1]] = {}
    for q, t, d in zip(eq.tolist(), et.tolist(), ed.tolist()):
        out[L[q][1]][L[t][1]] = d
    return out


def get_exact_nearest_neighbor_graph(seq_to_acc_list_sorted, has_converged, params, _threads=None):
    """nearest_neighbor_graph.py:19-82; the Pool of ``params.nr_cores`` workers becomes threads."""
    cores = int(params.nr_cores)
    return get_nearest_neighbors(seq_to_acc_list_sorted, 0, 0, seq_to_acc_list_sorted, has_converged,
                                 params.neighbor_search_depth, _cores=cores,
                                 _threads=cores if _threads is None else _threads)


def get_exact_nearest_neighbor_graph_2set(seq_to_acc_list_sorted_all, target_accessions, params, _threads=None):
    """nearest_neighbor_graph.py:300-334."""
    cores = int(params.nr_cores)
    return get_nearest_neighbors_2set(seq_to_acc_list_sorted_all, 0, seq_to_acc_list_sorted_all,
                                      target_accessions, params.neighbor_search_depth, _cores=cores,
                                      _threads=cores if _threads is None else _threads)


def compute_nearest_neighbor_graph(S, has_converged, params):
    """nearest_neighbor_graph.py:237-296 (without the prints)."""
    seq_to_acc = {seq: acc for (acc, seq) in S.items()}                       # :243
    seq_to_acc_list_sorted = sorted(seq_to_acc.items(), key=lambda x: len(x[0]))  # :245-246
    graph = get_exact_nearest_neighbor_graph(seq_to_acc_list_sorted, has_converged, params)
    s1 = set(S[a] for a in graph)
    isolated = set(seq_to_acc) - s1                                          # :267-272
    return graph, isolated


def compute_2set_nearest_neighbor_graph(X, C, params):
    """nearest_neighbor_graph.py:201-234 (without the prints)."""
    q = [(seq, acc) for (acc, seq) in X.items()]
    t = [(seq, acc) for (acc, seq) in C.items()]
    sorted_all = sorted(q + t, key=lambda x: len(x[0]))                      # :208
    return get_exact_nearest_neighbor_graph_2set(sorted_all, set(C.keys()), params)


def closed_form_1set(sorted_list, has_converged, depth=2 ** 32):
    """SURVEY.md Appendix A.3 brute force (small inputs only): an independent third opinion."""
    L = sorted_list
    n = len(L)
    out = {}
    for i in range(n):
        out[L[i][1]] = {}
        if L[i][0] in has_converged:
            continue
        cand = []
        for j in range(1, n):
            if j > depth:
                break
            for t in (i - j, i + j):
                if 0 <= t < n:
                    cand.append((t, ed_plain(L[i][0], L[t][0])))
        cand = [(t, d) for t, d in cand if d > 0 or len(L[i][0]) == 0]
        if not cand:
            continue
        dmin = min(d for _, d in cand)
        if dmin <= len(L[i][0]):
            for t, d in cand:
                if d == dmin:
                    out[L[i][1]][L[t][1]] = d
    return out


def closed_form_2set(sorted_all, target_accessions):
    """Appendix A.3, 2-set, default depth."""
    L = sorted_all
    n = len(L)
    out = {}
    for i in range(n):
        if L[i][1] in target_accessions:
            continue
        out[L[i][1]] = {}
        cand = []
        for j in range(1, n):
            for t in (i - j, i + j):
                if 0 <= t < n and L[t][1] in target_accessions:
                    cand.append((t, ed_plain(L[i][0], L[t][0])))
        if not cand:
            continue
        dmin = min(d for _, d in cand)
        if dmin <= len(L[i][0]):
            for t, d in cand:
                if d == dmin:
                    out[L[i][1]][L[t][1]] = d
    return out
