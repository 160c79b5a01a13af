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
pool); the "processing i" progress lines are not printed; the list must be sorted by length,
which every caller in IsoCon guarantees (:246, :208).  Reads stay resident on the device
between calls: a call uploads only the sequences the device has not seen (see ``_binding.use_list``).

With ``torch.distributed`` initialised (one process per GPU, NCCL) every rank calls these
functions with the same arguments; the row tiles of the pair matrix are split across the
ranks and every rank returns the complete graph (``isocon_b200.sharding``).

Install over the reference with ``isocon_b200.install()``.
"""
from __future__ import print_function

import threading

import numpy as np

from . import _binding
from . import _hostops
from . import sharding


def _ctx():
    """ONE context per device for the graph builders, ``edlib_ed`` and the pair-list module: they share the
    resident read store, so a pair list over reads of the last graph uploads nothing."""
    return _binding.get_context()


def edlib_ed(x, y, mode="NW", task="distance", k=1):
    """nearest_neighbor_graph.py:104-107: global edit distance, -1 when it exceeds k (k < 0: no bound)."""
    if mode != "NW" or task != "distance":
        raise NotImplementedError("the device path provides mode='NW', task='distance' only")
    ctx = _ctx()
    if len(x) <= len(y):
        ctx.use_list([x, y]); a, b = 0, 1
    else:
        ctx.use_list([y, x]); a, b = 1, 0
    return int(ctx.ed_pairs([a], [b], [int(k)])[0])


def get_nearest_neighbors_helper(arguments):
    args, kwargs = arguments
    return get_nearest_neighbors(*args, **kwargs)


def get_nearest_neighbors_2set_helper(arguments):
    args, kwargs = arguments
    return get_nearest_neighbors_2set(*args, **kwargs)


def _order_edges(eq, et, ed):
    """Scan order of the reference: per query by offset j = |t - q|, down (t < q) before up.  (numpy statement of
    what ``_hostops.build_graph`` does while it fills the dicts; used by the sharding tests.)"""
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


def _sorted_by_length(seqs, accs):
    """Stable sort by length (:246, :208) of two parallel lists; returns (seqs, accs, lens) in list order."""
    lens = _hostops.lengths(seqs)
    n = len(seqs)
    if n > 1 and (lens[1:] < lens[:-1]).any():
        # (numpy's stable sort is a radix sort for 16-bit keys: 8x faster than the merge sort of int64)
        order = np.argsort(lens.astype(np.uint16) if int(lens.max()) < 65536 else lens, kind="stable")
        seqs, accs, lens = _hostops.permute(seqs, order), _hostops.permute(accs, order), lens[order]
    return seqs, accs, lens


def _build_graph(seqs, accs, lens, mode, is_query, is_target, depth, lo, hi):
    """Run the device graph over the list (parallel lists of sequences and accessions, sorted by length) and
    rebuild the reference's dict-of-dicts: a key for every entry of [lo, hi) that is not a target, in list order,
    neighbours inserted in scan order."""
    ctx = _ctx()
    ctx.use_list(seqs, lens)
    skip = is_target if mode == 2 else None
    if hi - lo < 2000:
        best, eq, et, ed = sharding.device_graph(ctx, mode, depth, is_query, is_target)
        return _hostops.build_graph(accs, lo, hi, skip, eq, et, ed)
    # the empty result dicts need nothing from the device: a helper thread makes them while the kernels run.  Making
    # dicts holds the GIL, the library calls release it: the helper is let go the moment this thread enters the
    # library for the device work (NNContext.on_run), so its work falls inside that call and not in front of it.
    box = []
    go = threading.Event()

    def prepare():
        go.wait()
        try:
            box.append(_hostops.prepare_graph(accs, lo, hi, skip))
        except BaseException as e:          # re-raised on the calling thread
            box.append(e)

    helper = threading.Thread(target=prepare)
    helper.start()
    ctx.on_run = go.set
    try:
        best, eq, et, ed = sharding.device_graph(ctx, mode, depth, is_query, is_target)
    finally:
        ctx.on_run = None
        go.set()                            # (the device work may have failed before it began)
        helper.join()
    if isinstance(box[0], BaseException):
        raise box[0]
    return _hostops.fill_graph(box[0], accs, lo, hi, eq, et, ed)


def _unzip(L):
    if not L:
        return [], []
    seqs, accs = zip(*L)
    return list(seqs), list(accs)


def _queries_1set(seqs, has_converged, lo, hi):
    is_query = np.zeros(len(seqs), dtype=np.uint8)
    if has_converged:
        is_query[lo:hi] = 1 - _hostops.contains(has_converged, seqs[lo:hi])
    else:
        is_query[lo:hi] = 1
    return is_query


def get_nearest_neighbors(batch_of_queries, global_index_in_matrix, start_index, seq_to_acc_list_sorted,
                          has_converged, neighbor_search_depth):
    """nearest_neighbor_graph.py:110-198.  Queries are the list entries
    ``[start_index, start_index + len(batch_of_queries))``; entries whose sequence is in
    ``has_converged`` get an empty dict and are still neighbours of the others."""
    seqs, accs = _unzip(seq_to_acc_list_sorted)
    lo, hi = start_index, start_index + len(batch_of_queries)
    return _build_graph(seqs, accs, None, 1, _queries_1set(seqs, has_converged, lo, hi), None,
                        neighbor_search_depth, lo, hi)


def _masks_2set(accs, target_accessions, lo, hi):
    n = len(accs)
    is_target = _hostops.contains(target_accessions, accs)
    is_query = np.zeros(n, dtype=np.uint8)
    is_query[lo:hi] = 1 - is_target[lo:hi]
    return is_query, is_target


def get_nearest_neighbors_2set(batch, start_index, seq_to_acc_list_sorted, target_accessions, neighbor_search_depth):
    """nearest_neighbor_graph.py:341-424.  Entries whose accession is in ``target_accessions``
    are the candidates; every other entry of the batch range is a query."""
    seqs, accs = _unzip(seq_to_acc_list_sorted)
    lo, hi = start_index, start_index + len(batch)
    is_query, is_target = _masks_2set(accs, target_accessions, lo, hi)
    return _build_graph(seqs, accs, None, 2, is_query, is_target, neighbor_search_depth, lo, hi)


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
    seqs = list(X.values()); seqs.extend(C.values())
    accs = list(X.keys()); accs.extend(C.keys())
    seqs, accs, lens = _sorted_by_length(seqs, accs)
    n = len(seqs)
    is_query, is_target = _masks_2set(accs, C, 0, n)
    graph = _build_graph(seqs, accs, lens, 2, is_query, is_target, params.neighbor_search_depth, 0, n)
    _verbose_summary(graph, params)
    return graph


def compute_nearest_neighbor_graph(S, has_converged, params):
    """nearest_neighbor_graph.py:237-296: S holds unique strings; returns (graph, isolated)."""
    by_seq = dict(zip(S.values(), S.keys()))                  # seq -> acc, last accession wins, like :243
    seqs, accs, lens = _sorted_by_length(list(by_seq.keys()), list(by_seq.values()))
    n = len(seqs)
    graph = _build_graph(seqs, accs, lens, 1, _queries_1set(seqs, has_converged, 0, n), None,
                         params.neighbor_search_depth, 0, n)
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
