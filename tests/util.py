"""Shared helpers of the test-suite (fixtures of tests/golden/, comparisons)."""
import gzip
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


class Params(object):
    """Like modules/isocon_parameters.py:10-19 with the fields the path reads."""

    def __init__(self, nr_cores=1, neighbor_search_depth=2 ** 32, verbose=False, develop_logfile=None):
        self.nr_cores = nr_cores
        self.neighbor_search_depth = neighbor_search_depth
        self.verbose = verbose
        self.develop_logfile = develop_logfile


def as_lists(G):
    return [[a, [[b, d] for b, d in v.items()]] for a, v in G.items()]


def assert_same_graph(got, want, what=""):
    """dict equality AND identical key / insertion order."""
    if isinstance(want, dict):
        want = as_lists(want)
    got = as_lists(got)
    if got == want:
        return
    gd, wd = dict((a, v) for a, v in got), dict((a, v) for a, v in want)
    assert list(gd) == list(wd), "%s: different query keys / key order" % what
    for a in wd:
        assert gd[a] == wd[a], "%s: query %s: got %r want %r" % (what, a, gd[a][:8], wd[a][:8])
    raise AssertionError(what)


def _is_foreign_case(case):
    return case["name"].startswith(("K9", "K15", "K16", "alpha"))


def known_answers(foreign=False):
    """Known-answer cases of the unmodified reference driver (oracle/make_golden.py, make_golden_r2.py).
    foreign=True: the cases with symbols beyond upper-case ACGT; False: all others."""
    cases = []
    for name in ("known_answers.json", "known_answers_r2.json"):
        with open(os.path.join(GOLD, name)) as fh:
            cases += json.load(fh)["cases"]
    return [c for c in cases if _is_foreign_case(c) == foreign]


def correction_rounds():
    with gzip.open(os.path.join(GOLD, "rounds_r2.json.gz"), "rt") as fh:
        return json.load(fh)["rounds"]


_C1 = None


def c1_expected():
    global _C1
    if _C1 is None:
        with gzip.open(os.path.join(GOLD, "c1_expected.json.gz"), "rt") as fh:
            _C1 = json.load(fh)["fixtures"]
    return _C1


def load_reads(n):
    """The reads of test/data/simulated_pacbio_reads_n_<n>.fa as {acc: seq} in file order."""
    z = np.load(os.path.join(GOLD, "c1_n%d.npz" % n))
    codes = np.empty(z["packed"].size * 4, dtype=np.uint8)
    for i in range(4):
        codes[i::4] = (z["packed"] >> (2 * i)) & 3
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes().decode()
    out, pos = {}, 0
    for acc, ln in zip(z["acc"].tolist(), z["lens"].tolist()):
        out[acc] = text[pos:pos + ln]
        pos += ln
    return out


def two_set_split(S, every=17):
    accs = list(S.keys())
    C = {"cand_%d" % i: S[a] for i, a in enumerate(accs[::every])}
    return dict(S), C


def case_kwargs(case):
    p = case.get("params", {})
    return dict(nr_cores=p.get("nr_cores", 1), neighbor_search_depth=p.get("neighbor_search_depth", 2 ** 32))


def run_case(nn, case):
    """Run one known-answer case through a module with the reference's API (oracle or product)."""
    P = Params(**case_kwargs(case))
    if case["kind"] == "1set":
        G, iso = nn.compute_nearest_neighbor_graph(case["S"], set(case.get("has_converged", [])), P)
        assert iso == set()
        return G
    return nn.compute_2set_nearest_neighbor_graph(case["X"], case["C"], P)
