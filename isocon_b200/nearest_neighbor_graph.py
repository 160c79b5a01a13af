"""Drop-in replacement for IsoCon's ``modules/nearest_neighbor_graph.py`` on B200.

Same function names, argument meaning, return values and error behaviour as the reference
module (``/root/reference/modules/nearest_neighbor_graph.py``); the arithmetic runs in
``libisocon_nn.so`` (hand-written sm_100a kernels) instead of one ``edlib.align`` call per
pair inside a ``multiprocessing.Pool``:

=============================================  =========================================
this module                                    reference
=============================================  =========================================
``edlib_ed``                                   nearest_neighbor_graph.py:104-107
``get_nearest_neighbors``                      :110-198
``get_nearest_neighbors_2set``                 :341-424
``get_exact_nearest_neighbor_graph``           :19-82
``get_exact_nearest_neighbor_graph_2set``      :300-334
``compute_nearest_neighbor_graph``             :237-296
``compute_2set_nearest_neighbor_graph``        :201-234
``get_nearest_neighbors[_2set]_helper``        :15-17, :337-339
=============================================  =========================================

Differences, all outside the results: ``params.nr_cores`` is ignored (the GPU replaces the
pool); the "processing i" progress lines are not printed; sequences must be upper-case
``ACGT`` (``ValueError`` otherwise -- the reads are 2-bit packed on the device) and the list
must be sorted by length, which every caller in IsoCon guarantees (:246, :208).

With ``torch.distributed`` initialised (one process per GPU, NCCL) every rank calls these
functions with the same arguments; the row tiles of the pair matrix are split across the
ranks and every rank returns the complete graph (``isocon_b200.sharding``).

Install over the reference with ``isocon_b200.install()``.
"""
from __future__ import print_function

import os

import numpy as np

from . import _binding
from . import sharding

_PAIR_CTX = {}


def _ctx():
    return _binding.get_context()


def _pair_ctx():
    dev = _binding.default_device()
    if dev not in _PAIR_CTX:
        _PAIR_CTX[dev] = _binding.NNContext(dev)
    return _PAIR_CTX[dev]


def edlib_ed(x, y, mode="NW", task="distance", k=1):
    """nearest_neighbor_graph.py:104-107: global edit distance, -1 when it exceeds k (k < 0: no bound)."""
    if mode != "NW" or task != "distance":
        raise NotImplementedError("the device path provides mode='NW', task='distance' only")
    ctx = _pair_ctx()
    if len(x) <= len(y):
        ctx.set_reads([x, y]); a, b = 0, 1
    else:
        ctx.set_reads([y, x]); a, b = 1, 0
    return int(ctx.ed_pairs([a], [b], [int(k)])[0])


def get_nearest_neighbors_helper(arguments):
    args, kwargs = arguments
    return get_nearest_neighbors(*args, **kwargs)


def get_nearest_neighbors_2set_helper(arguments):
    args, kwargs = arguments
    return get_nearest_neighbors_2set(*args, **kwargs)


def _order_edges(eq, et, ed):
    """Scan order of the reference: per query by offset j = |t - q|, down (t < q) before up."""
    if eq.size == 0:
        return eq, et, ed
    q64, t64 = eq.astype(np.int64), et.astype(np.int64)     # list indices < 2**30: the three fields do not overlap
    key = (q64 << 33) | (np.abs(t64 - q64) << 1) | (t64 > q64)     # one sort key: (q, |t - q|, up)
    order = np.argsort(key, kind="stable")
    key = key[order]
    keep = np.ones(key.size, dtype=bool)
    keep[1:] = key[1:] != key[:-1]                                # an edge may be reported twice
    order = order[keep]
    return eq[order], et[order], ed[order]


def _build_graph(L, mode, is_query, is_target, depth, key_range):
    """Run the device graph and rebuild the reference's dict-of-dicts (key and insertion order)."""
    ctx = _ctx()
    ctx.set_reads([s for s, _ in L])
    best, eq, et, ed = sharding.device_graph(ctx, mode, depth, is_query, is_target)
    eq, et, ed = _order_edges(eq, et, ed)
    accs = [a for _, a in L]
    if mode == 1:
        out = {a: {} for a in accs[key_range.start:key_range.stop]}
    else:
        out = {accs[i]: {} for i in key_range if not is_target[i]}
    for q, t, d in zip(eq.tolist(), et.tolist(), ed.tolist()):
        out[accs[q]][accs[t]] = d
    return out


def get_nearest_neighbors(batch_of_queries, global_index_in_matrix, start_index, seq_to_acc_list_sorted,
                          has_converged, neighbor_search_depth):
    """nearest_neighbor_graph.py:110-198.  Queries are the list entries
    ``[start_index, start_index + len(batch_of_queries))``; entries whose sequence is in
    ``has_converged`` get an empty dict and are still neighbours of the others."""
    L = seq_to_acc_list_sorted
    n = len(L)
    lo, hi = start_index, start_index + len(batch_of_queries)
    is_query = np.zeros(n, dtype=np.uint8)
    if has_converged:
        is_query[lo:hi] = [0 if L[i][0] in has_converged else 1 for i in range(lo, hi)]
    else:
        is_query[lo:hi] = 1
    return _build_graph(L, 1, is_query, None, neighbor_search_depth, range(lo, hi))


def get_nearest_neighbors_2set(batch, start_index, seq_to_acc_list_sorted, target_accessions, neighbor_search_depth):
    """nearest_neighbor_graph.py:341-424.  Entries whose accession is in ``target_accessions``
    are the candidates; every other entry of the batch range is a query."""
    L = seq_to_acc_list_sorted
    n = len(L)
    lo, hi = start_index, start_index + len(batch)
    is_target = np.fromiter((1 if a in target_accessions else 0 for _, a in L), dtype=np.uint8, count=n)
    is_query = np.zeros(n, dtype=np.uint8)
    is_query[lo:hi] = 1 - is_target[lo:hi]
    return _build_graph(L, 2, is_query, is_target, neighbor_search_depth, range(lo, hi))


def get_exact_nearest_neighbor_graph(seq_to_acc_list_sorted, has_converged, params):
    """nearest_neighbor_graph.py:19-82.  The Pool of ``params.nr_cores`` workers is replaced by
    the GPU(s); chunking never changes the result."""
    return get_nearest_neighbors(seq_to_acc_list_sorted, 0, 0, seq_to_acc_list_sorted, has_converged,
                                 params.neighbor_search_depth)


def get_exact_nearest_neighbor_graph_2set(seq_to_acc_list_sorted_all, target_accessions, params):
    """nearest_neighbor_graph.py:300-334."""
    return get_nearest_neighbors_2set(seq_to_acc_list_sorted_all, 0, seq_to_acc_list_sorted_all, target_accessions,
                                      params.neighbor_search_depth)


def _verbose_summary(graph, params):
    """The three summary lines of :226-229 / :288-291 (ZeroDivisionError with no edges, as there)."""
    if params.verbose:
        edges = sum(len(nbrs) for nbrs in graph.values())
        tot_ed = sum(d for nbrs in graph.values() for d in nbrs.values())
        print("Number of edges:", edges)
        print("Total edit distance:", tot_ed)
        print("Avg ed (ed/edges):", tot_ed / float(edges))


def compute_2set_nearest_neighbor_graph(X, C, params):
    """nearest_neighbor_graph.py:201-234: reads X and candidates C (dicts acc -> seq, no dedup),
    merged and stably sorted by length, reads before candidates at equal length."""
    merged = [(seq, acc) for acc, seq in X.items()]
    merged.extend((seq, acc) for acc, seq in C.items())
    merged.sort(key=lambda entry: len(entry[0]))
    graph = get_exact_nearest_neighbor_graph_2set(merged, set(C.keys()), params)
    _verbose_summary(graph, params)
    return graph


def compute_nearest_neighbor_graph(S, has_converged, params):
    """nearest_neighbor_graph.py:237-296: S holds unique strings; returns (graph, isolated)."""
    by_seq = {}
    for acc, seq in S.items():
        by_seq[seq] = acc                                     # last accession wins, like :243
    ordered = sorted(by_seq.items(), key=lambda entry: len(entry[0]))
    graph = get_exact_nearest_neighbor_graph(ordered, has_converged, params)
    if len(graph) == len(by_seq):
        isolated = set()                                      # every entry of the list is a key (:120, :267-272)
    else:
        isolated = set(by_seq) - set(S[acc] for acc in graph)
    print("isolated:", len(isolated))
    _verbose_summary(graph, params)
    return graph, isolated


def read_fasta(fasta_file):
    """nearest_neighbor_graph.py:84-101 (kept because the reference module exports it)."""
    acc, temp = None, []
    for line in fasta_file:
        if line[0] == '>':
            if acc is not None:
                yield acc, "".join(temp)
            acc, temp = line[1:].strip(), []
        else:
            temp.append(line.strip())
    if acc:
        yield acc, "".join(temp)
