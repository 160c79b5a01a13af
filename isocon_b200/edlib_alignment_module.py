"""Drop-in replacement for IsoCon's ``modules/edlib_alignment_module.py`` on B200.

The reference aligns an explicit list of pairs with one unbounded ``edlib.align(x, y, "NW")``
call each, inside a ``multiprocessing.Pool`` (``/root/reference/modules/edlib_alignment_module.py``);
it is called right after every graph build (``isocon_get_candidates.py:38,301``,
``isocon_statistical_test.py:289``).  Here all pairs of a call go to the device in one batch
(``isocon_nn_ed_pairs``: the same banded bit-vector arithmetic as the graph, threshold doubling
until the distance fits).

======================================================  ===================================
this module                                             reference
======================================================  ===================================
``edlib_align_sequences(matches, nr_cores)``            edlib_alignment_module.py:10-49
``edlib_align_sequences_keeping_accession(matches, ..)``  :51-99
``edlib_alignment_helper`` / ``edlib_alignment``        :103-128
``edlib_traceback``                                     :130-135 (paths are out of scope: raises)
======================================================  ===================================

``nr_cores`` is ignored.  Results (keys, key order, values) are those of the reference.
"""
import numpy as np

from . import _binding

def _ctx():
    """The graph builders' context: the pair lists IsoCon asks for right after a graph build
    (isocon_get_candidates.py:38,301; isocon_statistical_test.py:289) are over reads that are already resident."""
    return _binding.get_context()


def _distances(pairs):
    """Unbounded NW distance of every (x, y) string pair, batched on the device."""
    if not pairs:
        return []
    index = {}
    for x, y in pairs:
        if x not in index:
            index[x] = len(index)
        if y not in index:
            index[y] = len(index)
    seqs = sorted(index, key=len)                 # the library wants the list sorted by length
    pos = {s: i for i, s in enumerate(seqs)}
    a = np.fromiter((pos[x] for x, _ in pairs), dtype=np.int32, count=len(pairs))
    b = np.fromiter((pos[y] for _, y in pairs), dtype=np.int32, count=len(pairs))
    ctx = _ctx()
    ctx.use_list(seqs)                            # uploads only sequences the device has not seen
    return ctx.ed_pairs(a, b, None).tolist()


def edlib_align_sequences(matches, nr_cores=1):
    """edlib_alignment_module.py:10-49: ``matches`` maps a sequence to an iterable of sequences;
    returns ``{s1: {s2: edit distance}}``."""
    todo, seen = [], {}
    for s1 in matches:
        for s2 in matches[s1]:
            if s2 in seen.setdefault(s1, set()):
                continue                           # :14-16 (a repeated partner is aligned once)
            seen[s1].add(s2)
            todo.append((s1, s2))
    exact_edit_distances = {}
    for (s1, s2), ed in zip(todo, _distances(todo)):
        assert ed >= 0                             # :113
        exact_edit_distances.setdefault(s1, {})[s2] = ed
    return exact_edit_distances


def edlib_align_sequences_keeping_accession(matches, nr_cores=1):
    """edlib_alignment_module.py:51-99: ``matches[s1_acc][s2_acc] = (s1, s2)``; returns the same
    structure with ``(s1, s2, edit distance)``."""
    todo = [(s1_acc, s2_acc) for s1_acc in matches for s2_acc in matches[s1_acc]]
    eds = _distances([matches[a][b] for a, b in todo])
    exact_matches = {}
    for (s1_acc, s2_acc), ed in zip(todo, eds):
        assert ed >= 0
        s1, s2 = matches[s1_acc][s2_acc]
        exact_matches.setdefault(s1_acc, {})[s2_acc] = (s1, s2, ed)
    return exact_matches


def edlib_alignment(x, y, i, j, x_acc="", y_acc=""):
    """edlib_alignment_module.py:107-128 (one pair)."""
    ed = _distances([(x, y)])[0]
    assert ed >= 0
    if x_acc == y_acc == "":
        return (x, y, ed)
    return (x_acc, y_acc, (x, y, ed))


def edlib_alignment_helper(arguments):
    args, kwargs = arguments
    return edlib_alignment(*args, **kwargs)


def edlib_traceback(x, y, mode="NW", task="path", k=1):
    raise NotImplementedError("alignment paths (task='path') are outside the device path (SURVEY.md §8f-2/3)")
